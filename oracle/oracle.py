"""CPU oracle for the CRISPResso2 per-read hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (crispresso2_b200/) never
does; it fails loudly when its CUDA library is missing.

Parity status: PINNED (see c2_oracle.c header and tests/test_oracle.py).

What is restated here (reference file:line, all under /root/reference):
  read_matrix / make_matrix      CRISPResso2/CRISPResso2Align.pyx:33-99
  global_align                   CRISPResso2/CRISPResso2Align.pyx:101-434   (C: c2o_global_align)
  find_indels_substitutions      CRISPResso2/CRISPRessoCOREResources.pyx:68-187 (C: c2o_find_indels)
  reverse_complement             CRISPResso2/CRISPRessoShared.py:399-403
  new_variant                    CRISPResso2/CRISPRessoCORE.py:627-798  (get_new_variant_object, incl. the
                                 prime-editing scaffold branch :789-796 and the legacy insertion switch)
  pe_scaffold_search             CRISPResso2/plots/data_prep.py:3827-3860 (get_pe_scaffold_search)
  process_reads                  CRISPResso2/CRISPRessoCORE.py:1956-2000 (serial branch of process_fastq)
  count_vectors                  CRISPResso2/CRISPRessoCORE.py:3964-4181 (vectors, counters, the size Counters and the
                                 --coding_seq frameshift / splicing block), pinned against the reference's own
                                 CorePlotContext arguments by tests/test_cli_dropin.py
  ref1_vectors                   CRISPResso2/CRISPRessoCORE.py:4195-4272 (HDR re-projection)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "c2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    if os.path.isdir("/root/reference/CRISPResso2"):
        subprocess.call(["make", "-s", "-C", _HERE, "ref"])
    return so


class _Edits(C.Structure):
    _fields_ = [
        ("ref_positions", C.POINTER(C.c_int32)), ("n_ref_positions", C.c_int32),
        ("all_sub_pos", C.POINTER(C.c_int32)), ("n_all_sub", C.c_int32), ("all_sub_val", C.POINTER(C.c_uint8)),
        ("sub_pos", C.POINTER(C.c_int32)), ("n_sub", C.c_int32), ("sub_val", C.POINTER(C.c_uint8)),
        ("all_del_pos", C.POINTER(C.c_int32)), ("n_all_del_pos", C.c_int32),
        ("all_del_coord", C.POINTER(C.c_int32)), ("n_all_del_coord", C.c_int32),
        ("del_pos", C.POINTER(C.c_int32)), ("n_del_pos", C.c_int32),
        ("del_coord", C.POINTER(C.c_int32)), ("n_del_coord", C.c_int32),
        ("del_sizes", C.POINTER(C.c_int32)), ("n_del_sizes", C.c_int32),
        ("all_ins_pos", C.POINTER(C.c_int32)), ("n_all_ins_pos", C.c_int32),
        ("all_ins_left", C.POINTER(C.c_int32)), ("n_all_ins_left", C.c_int32),
        ("ins_pos", C.POINTER(C.c_int32)), ("n_ins_pos", C.c_int32),
        ("ins_coord", C.POINTER(C.c_int32)), ("n_ins_coord", C.c_int32),
        ("ins_sizes", C.POINTER(C.c_int32)), ("n_ins_sizes", C.c_int32),
        ("substitution_n", C.c_int64), ("deletion_n", C.c_int64), ("insertion_n", C.c_int64),
    ]


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.c2o_global_align.restype = C.c_int
        L.c2o_global_align.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.c2o_find_indels.restype = C.c_int
        L.c2o_find_indels.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.POINTER(_Edits), C.c_int64]
        L.c2o_batch_align_classify.restype = C.c_int64
        L.c2o_batch_align_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p,
                                               C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                               C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- matrices

def read_matrix(path):
    """NCBI-format substitution matrix -> int64 table indexed [ord(row), ord(col)]."""
    with open(path) as fh:
        rows = [ln.rstrip("\n") for ln in fh]
    k = 0
    while not rows[k].strip() or rows[k].lstrip()[0] == "#":
        k += 1
    cols = [ord(t) for t in rows[k].split()]
    n = max(cols) + 1
    tab = np.zeros((n, n), dtype=np.int64)
    r = 0
    for ln in rows[k + 1:]:
        if not ln:
            continue
        toks = ln.split()
        vals = [int(t) for t in toks[1:]]
        for c, v in zip(cols, vals):
            tab[cols[r], c] = v
        r += 1
    return tab


def make_matrix(match_score=5, mismatch_score=-4, n_mismatch_score=-2, n_match_score=-1):
    n = ord("T") + 1
    tab = np.zeros((n, n), dtype=np.int64)
    acgt = [ord(c) for c in "ATCG"]
    for a in acgt:
        for b in acgt:
            tab[a, b] = match_score if a == b else mismatch_score
        tab[a, ord("N")] = n_mismatch_score
        tab[ord("N"), a] = n_mismatch_score
    tab[ord("N"), ord("N")] = n_match_score
    return tab


_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "_": "_", "-": "-"}


def reverse_complement(seq):
    return "".join(_COMP[c] for c in reversed(seq.upper()))


# --------------------------------------------------------------------------- kernels

class OracleUndefined(Exception):
    """Input lies in the zone where the reference itself is undefined."""


def global_align_raw(read, ref, matrix, gap_incentive, gap_open=-1, gap_extend=-1):
    """-> (aligned_read, aligned_ref, n_match, aln_len)"""
    L = lib()
    rb, fb = read.encode(), ref.encode()
    matrix = np.ascontiguousarray(matrix, dtype=np.int64)
    gi = np.ascontiguousarray(gap_incentive, dtype=np.int64)
    if len(gi) != len(fb) + 1:
        raise ValueError("gap_incentive length mismatch")
    cap = len(rb) + len(fb) + 1
    oj = C.create_string_buffer(cap)
    oi = C.create_string_buffer(cap)
    n = C.c_int32(0)
    m = C.c_int32(0)
    rc = L.c2o_global_align(rb, len(rb), fb, len(fb), matrix.ctypes.data, matrix.shape[1], gi.ctypes.data,
                            int(gap_open), int(gap_extend), oj, oi, C.byref(n), C.byref(m))
    if rc != 0:
        raise OracleUndefined("c2o_global_align rc=%d" % rc)
    return oj.raw[:n.value].decode(), oi.raw[:n.value].decode(), m.value, n.value


def score_from_counts(n_match, aln_len):
    """Align.pyx:433-434: round(100*matchCount/float(align_counter), 3)."""
    return round(100 * n_match / float(aln_len), 3)


def global_align(read, ref, matrix, gap_incentive, gap_open=-1, gap_extend=-1):
    s1, s2, m, n = global_align_raw(read, ref, matrix, gap_incentive, gap_open, gap_extend)
    return s1, s2, score_from_counts(m, n)


def _ilist(ptr, n):
    return [int(ptr[k]) for k in range(n)]


def _pairs(ptr, n):
    return [(int(ptr[2 * k]), int(ptr[2 * k + 1])) for k in range(n)]


def find_indels_substitutions(read_al, ref_al, include_idx):
    """-> dict with the 18 fields of the reference payload (COREResources.pyx:165-186)."""
    L = lib()
    n = len(ref_al)
    inc = [int(v) for v in include_idx]
    mask_len = max([n + 2] + [v + 1 for v in inc])
    mask = np.zeros(mask_len, dtype=np.uint8)
    for v in inc:
        if v >= 0:
            mask[v] = 1
    cap = 2 * n + 8
    bufs = {}
    e = _Edits()
    for name, ctype in _Edits._fields_:
        if ctype is C.POINTER(C.c_int32):
            bufs[name] = (C.c_int32 * (2 * cap))()
            setattr(e, name, bufs[name])
        elif ctype is C.POINTER(C.c_uint8):
            bufs[name] = (C.c_uint8 * (2 * cap))()
            setattr(e, name, bufs[name])
    rc = L.c2o_find_indels(read_al.encode(), ref_al.encode(), n, mask.ctypes.data, mask_len, C.byref(e), 2 * cap)
    if rc != 0:
        raise OracleUndefined("c2o_find_indels rc=%d" % rc)
    return {
        "all_insertion_positions": _ilist(e.all_ins_pos, e.n_all_ins_pos),
        "all_insertion_left_positions": _ilist(e.all_ins_left, e.n_all_ins_left),
        "insertion_positions": _ilist(e.ins_pos, e.n_ins_pos),
        "insertion_coordinates": _pairs(e.ins_coord, e.n_ins_coord),
        "insertion_sizes": _ilist(e.ins_sizes, e.n_ins_sizes),
        "insertion_n": int(e.insertion_n),
        "all_deletion_positions": _ilist(e.all_del_pos, e.n_all_del_pos),
        "all_deletion_coordinates": _pairs(e.all_del_coord, e.n_all_del_coord),
        "deletion_positions": _ilist(e.del_pos, e.n_del_pos),
        "deletion_coordinates": _pairs(e.del_coord, e.n_del_coord),
        "deletion_sizes": _ilist(e.del_sizes, e.n_del_sizes),
        "deletion_n": int(e.deletion_n),
        "all_substitution_positions": _ilist(e.all_sub_pos, e.n_all_sub),
        "substitution_positions": _ilist(e.sub_pos, e.n_sub),
        "all_substitution_values": [chr(e.all_sub_val[k]) for k in range(e.n_all_sub)],
        "substitution_values": [chr(e.sub_val[k]) for k in range(e.n_sub)],
        "substitution_n": int(e.substitution_n),
        "ref_positions": _ilist(e.ref_positions, e.n_ref_positions),
    }


def find_indels_substitutions_legacy(read_al, ref_al, include_idx):
    """Restatement of COREResources.pyx:190-315 (`--use_legacy_insertion_quantification`).  Differences from the current function: a window insertion needs only ONE flank in the window
    (:284); deletion coordinates come from alignment columns with two end rules -- a run starting in column 0 or 1 is reported
    from reference position 0 (:252-254), a run reaching the last column ends at the last reference index, not one past it
    (:255-257); sizes are column counts; deletion_n / insertion_n are numpy sums (0.0 for an empty list)."""
    import re
    ref_positions, all_sub_pos, sub_pos, all_sub_val, sub_val = [], [], [], [], []
    inc = set(int(v) for v in include_idx)
    idx = 0
    for k, c in enumerate(ref_al):
        if c in "ATCGN":
            ref_positions.append(idx)
            if ref_al[k] != read_al[k] and read_al[k] != "-" and read_al[k] != "N":
                all_sub_pos.append(idx)
                all_sub_val.append(read_al[k])
                if idx in inc:
                    sub_pos.append(idx)
                    sub_val.append(read_al[k])
            idx += 1
        else:
            ref_positions.append(-1 if idx == 0 else -idx)
    all_del_pos, del_pos, del_coords, all_del_coords, del_sizes = [], [], [], [], []
    all_ins_pos, all_ins_left, ins_pos, ins_coords, ins_sizes = [], [], [], [], []
    for mt in re.finditer("-+", read_al):
        st, en = mt.span()
        ref_st = ref_positions[st] if st - 1 > 0 else 0
        ref_en = ref_positions[en] if en < len(ref_positions) else idx - 1
        all_del_pos.extend(range(ref_st, ref_en))
        all_del_coords.append((ref_st, ref_en))
        if inc.intersection(range(ref_st, ref_en)):
            del_pos.extend(range(ref_st, ref_en))
            del_coords.append((ref_st, ref_en))
            del_sizes.append(en - st)
    for mt in re.finditer("-+", ref_al):
        st, en = mt.span()
        if st == 0 or en == len(ref_al):
            continue
        ref_st, ref_en = ref_positions[st - 1], ref_positions[en]
        all_ins_left.append(ref_st)
        all_ins_pos.extend([ref_st, ref_en])
        if ref_st in inc or ref_en in inc:
            ins_coords.append((ref_st, ref_en))
            ins_pos.extend([ref_st, ref_en])
            ins_sizes.append(en - st)
    return {"all_insertion_positions": all_ins_pos, "all_insertion_left_positions": all_ins_left,
            "insertion_positions": ins_pos, "insertion_coordinates": ins_coords, "insertion_sizes": ins_sizes,
            "insertion_n": np.sum(ins_sizes), "all_deletion_positions": all_del_pos, "deletion_positions": del_pos,
            "deletion_coordinates": del_coords, "all_deletion_coordinates": all_del_coords, "deletion_sizes": del_sizes,
            "deletion_n": np.sum(del_sizes), "all_substitution_positions": all_sub_pos, "substitution_positions": sub_pos,
            "all_substitution_values": np.array(all_sub_val), "substitution_values": np.array(sub_val),
            "substitution_n": len(sub_pos), "ref_positions": ref_positions}


# --------------------------------------------------------------------------- per-read logic

class Params:
    """The `args` fields the hot path reads, with the reference defaults."""

    def __init__(self, **kw):
        self.aln_seed_count = 5
        self.aln_seed_min = 2
        self.needleman_wunsch_gap_open = -20
        self.needleman_wunsch_gap_extend = -2
        self.ignore_substitutions = False
        self.ignore_insertions = False
        self.ignore_deletions = False
        self.assign_ambiguous_alignments_to_first_reference = False
        self.expand_ambiguous_alignments = False
        self.discard_indel_reads = False
        self.expected_hdr_amplicon_seq = ""
        self.use_legacy_insertion_quantification = False
        for k, v in kw.items():
            setattr(self, k, v)


def _strand_choice(params, seq, ref):
    """Seed test of CRISPRessoCORE.py:656-687 -> 'fw' | 'rc' | 'both'."""
    nf = nr = 0
    for k in range(min(params.aln_seed_count, len(ref["fw_seeds"]))):
        if ref["fw_seeds"][k] in seq:
            nf += 1
        if ref["rc_seeds"][k] in seq:
            nr += 1
    if nf > params.aln_seed_min and nr == 0:
        return "fw"
    if nf == 0 and nr > params.aln_seed_min:
        return "rc"
    return "both"


def pe_scaffold_search(params, refs):
    """(scaffold start in the prime-edited amplicon, shortest telling scaffold prefix) -- plots/data_prep.py:3827-3860, called
    by process_fastq at CRISPRessoCORE.py:1814-1816; (0, None) without a scaffold sequence."""
    scaf = getattr(params, "prime_editing_pegRNA_scaffold_seq", "") or ""
    ext = getattr(params, "prime_editing_pegRNA_extension_seq", "") or ""
    if not scaf or not ext:
        return (0, None)
    pe = refs["Prime-edited"]["sequence"]
    sdna = reverse_complement(scaf.upper().replace("U", "T"))
    edna = reverse_complement(ext.upper().replace("U", "T"))
    start = pe.index(edna) + len(edna)
    n = getattr(params, "prime_editing_pegRNA_scaffold_min_match_length", 1)
    probe = edna + sdna[0:n]
    while probe in pe:
        if n > len(sdna):
            raise ValueError("scaffold found in the unedited reference sequence")
        n += 1
        probe = edna + sdna[0:n]
    return (start, sdna[0:n])


def new_variant(params, seq, refs, ref_names, matrix, pe_scaffold_dna_info=(0, None)):
    """Payload for one unique read (CRISPRessoCORE.py:627-798)."""
    go, ge = params.needleman_wunsch_gap_open, params.needleman_wunsch_gap_extend
    scores, details = [], []
    best = -1
    winners = []                       # (name, s1, s2, strand)
    for name in ref_names:
        ref = refs[name]
        mode = _strand_choice(params, seq, ref)
        strand = "+"
        if mode == "fw":
            s1, s2, sc = global_align(seq, ref["sequence"], matrix, ref["gap_incentive"], go, ge)
        elif mode == "rc":
            s1, s2, sc = global_align(reverse_complement(seq), ref["sequence"], matrix, ref["gap_incentive"], go, ge)
            strand = "-"
        else:
            s1, s2, sc = global_align(seq, ref["sequence"], matrix, ref["gap_incentive"], go, ge)
            r1, r2, rsc = global_align(reverse_complement(seq), ref["sequence"], matrix, ref["gap_incentive"], go, ge)
            if rsc > sc:
                s1, s2, sc, strand = r1, r2, rsc, "-"
        details.append((name, s1, s2, sc))
        scores.append(sc)
        if sc > best and sc > ref["min_aln_score"]:
            best = sc
            winners = [(name, s1, s2, strand)]
        elif sc == best:
            winners.append((name, s1, s2, strand))

    out = {"count": 1, "aln_scores": scores, "ref_aln_details": details, "best_match_score": best}
    if best <= 0:
        return out
    out["aln_ref_names"] = [w[0] for w in winners]
    labels = []
    for name, s1, s2, strand in winners:
        if getattr(params, "use_legacy_insertion_quantification", False):          # CRISPRessoCORE.py:721-724
            p = find_indels_substitutions_legacy(s1, s2, refs[name]["include_idxs"])
        else:
            p = find_indels_substitutions(s1, s2, refs[name]["include_idxs"])
        p["ref_name"] = name
        p["aln_scores"] = scores
        head_bad = s1[0] == "-" or s2[0] == "-" or s1[0] != s2[0]
        tail_bad = s1[-1] == "-" or s2[-1] == "-" or s1[-1] != s2[-1]
        p["irregular_ends"] = bool(head_bad or tail_bad)
        p["insertions_outside_window"] = int(len(p["all_insertion_positions"]) / 2 - len(p["insertion_positions"]) / 2)
        p["deletions_outside_window"] = len(p["all_deletion_coordinates"]) - len(p["deletion_coordinates"])
        p["substitutions_outside_window"] = len(p["all_substitution_positions"]) - len(p["substitution_positions"])
        p["total_mods"] = int(len(p["all_insertion_positions"]) / 2 + len(p["all_deletion_positions"])
                              + len(p["all_substitution_positions"]))
        p["mods_in_window"] = p["substitution_n"] + p["deletion_n"] + p["insertion_n"]
        p["mods_outside_window"] = p["total_mods"] - p["mods_in_window"]
        modified = False
        if not params.ignore_deletions and p["deletion_n"] > 0:
            modified = True
        elif not params.ignore_insertions and p["insertion_n"] > 0:
            modified = True
        elif not params.ignore_substitutions and p["substitution_n"] > 0:
            modified = True
        p["classification"] = "MODIFIED" if modified else "UNMODIFIED"
        labels.append(name + "_" + p["classification"])
        p["aln_seq"], p["aln_ref"], p["aln_strand"] = s1, s2, strand
        out["variant_" + name] = p
        out["best_match_name"] = name
    out["class_name"] = "&".join(labels)
    if len(winners) > 1:
        if params.assign_ambiguous_alignments_to_first_reference:
            out["class_name"] = labels[0]
            out["aln_ref_names"] = [winners[0][0]]
        elif not params.expand_ambiguous_alignments:
            out["class_name"] = "AMBIGUOUS"
    if getattr(params, "prime_editing_pegRNA_scaffold_seq", "") and "Prime-edited" in [w[0] for w in winners]:     # :789-796
        import copy
        loc, probe = pe_scaffold_dna_info
        pe = out["variant_Prime-edited"]
        at = pe["ref_positions"].index(loc - 1) + 1
        if pe["aln_seq"][at:at + len(probe)] == probe:
            out["aln_ref_names"] = ["Scaffold-incorporated"]
            out["class_name"] = "Scaffold-incorporated"
            moved = copy.deepcopy(pe)
            moved["ref_name"] = "Scaffold-incorporated"
            out["variant_Scaffold-incorporated"] = moved
    return out


def process_reads(read_iter, refs, ref_names, params, matrix):
    """Serial branch of process_fastq (CRISPRessoCORE.py:1825-1849, 1956-2000).

    read_iter yields read strings.  -> (variantCache, aln_stats, not_aligned)
    """
    cache = {}
    for s in read_iter:
        cache[s] = cache.get(s, 0) + 1
    st = dict.fromkeys(["N_TOT_READS", "N_CACHED_ALN", "N_CACHED_NOTALN", "N_COMPUTED_ALN", "N_COMPUTED_NOTALN",
                        "N_GLOBAL_SUBS", "N_SUBS_OUTSIDE_WINDOW", "N_MODS_IN_WINDOW", "N_MODS_OUTSIDE_WINDOW",
                        "N_READS_IRREGULAR_ENDS", "READ_LENGTH"], 0)
    lost = {}
    info = pe_scaffold_search(params, refs)
    for s in list(cache.keys()):
        c = cache[s]
        st["N_TOT_READS"] += c
        v = new_variant(params, s, refs, ref_names, matrix, info)
        v["count"] = c
        if v["best_match_score"] <= 0:
            st["N_COMPUTED_NOTALN"] += 1
            st["N_CACHED_NOTALN"] += c - 1
            lost[s] = v
            continue
        cache[s] = v
        st["N_COMPUTED_ALN"] += 1
        st["N_CACHED_ALN"] += c - 1
        p = v["variant_" + v["best_match_name"]]
        if st["READ_LENGTH"] == 0:
            st["READ_LENGTH"] = len(p["aln_seq"])
        st["N_GLOBAL_SUBS"] += (p["substitution_n"] + p["substitutions_outside_window"]) * c
        st["N_SUBS_OUTSIDE_WINDOW"] += p["substitutions_outside_window"] * c
        st["N_MODS_IN_WINDOW"] += p["mods_in_window"] * c
        st["N_MODS_OUTSIDE_WINDOW"] += p["mods_outside_window"] * c
        if p["irregular_ends"]:
            st["N_READS_IRREGULAR_ENDS"] += c
    for s in lost:
        del cache[s]
    return cache, st, lost


VECTOR_NAMES = [
    "all_insertion_count", "all_insertion_left_count", "all_deletion_count", "all_substitution_count",
    "insertion_count", "deletion_count", "substitution_count",
    "all_substitution_base_A", "all_substitution_base_C", "all_substitution_base_G", "all_substitution_base_T",
    "all_substitution_base_N",
    "all_base_count_A", "all_base_count_C", "all_base_count_G", "all_base_count_T", "all_base_count_N",
    "all_base_count_-",
    "insertion_length", "deletion_length",
    "insertion_count_noncoding", "deletion_count_noncoding", "substitution_count_noncoding",
]
SCALAR_NAMES = ["counts_total", "counts_modified", "counts_unmodified", "counts_discarded", "counts_insertion",
                "counts_deletion", "counts_substitution", "counts_only_insertion", "counts_only_deletion",
                "counts_only_substitution", "counts_insertion_and_deletion", "counts_insertion_and_substitution",
                "counts_deletion_and_substitution", "counts_insertion_and_deletion_and_substitution",
                "counts_modified_frameshift", "counts_modified_non_frameshift", "counts_non_modified_non_frameshift",
                "counts_splicing_sites_modified"]


def count_vectors(cache, refs, ref_names, params, extras=None):
    """Quantification loop (CRISPRessoCORE.py:3964-4181): per-position vectors and scalar counters, including the
    --coding_seq frameshift / splicing logic (:4083-4180).  Mutates cache[...]['count'] through the
    reverse-complement merge exactly like the reference.
    -> (vectors[ref][name] float64 arrays, scalars[ref][name] ints, class_counts, N_TOTAL)
    If `extras` is a dict it receives extras[ref] = the reference's Counters: inserted_n, deleted_n, substituted_n,
    effective_len (:4020-4043), hists_inframe, hists_frameshift (:3903-3906, :4134-4177)."""
    from collections import Counter
    vec = {r: {n: np.zeros(len(refs[r]["sequence"])) for n in VECTOR_NAMES} for r in ref_names}
    sca = {r: dict.fromkeys(SCALAR_NAMES, 0) for r in ref_names}
    ext = {r: {"inserted_n": Counter(), "deleted_n": Counter(), "substituted_n": Counter(), "effective_len": Counter(),
               "hists_inframe": Counter({0: 0}), "hists_frameshift": Counter({0: 0})} for r in ref_names}
    if extras is not None:
        extras.update(ext)
    classes = {}
    total = 0
    for s in cache:
        c = cache[s]["count"]
        if c == 0:
            continue
        rc = reverse_complement(s)
        if rc in cache and cache[rc]["count"] > 0:
            c += cache[rc]["count"]
            cache[rc]["count"] = 0
            cache[s]["count"] = c
        total += c
        v = cache[s]
        classes[v["class_name"]] = classes.get(v["class_name"], 0) + c
        if v["class_name"] == "AMBIGUOUS":
            continue
        for r in v["aln_ref_names"]:
            p = v["variant_" + r]
            V, S = vec[r], sca[r]
            if params.discard_indel_reads and (p["deletion_n"] > 0 or p["insertion_n"] > 0):
                S["counts_discarded"] += c
                continue
            S["counts_total"] += c
            S["counts_modified" if p["classification"] == "MODIFIED" else "counts_unmodified"] += c
            has_i = has_d = has_s = False
            V["all_insertion_count"][p["all_insertion_positions"]] += c        # repeated index counted once
            V["all_insertion_left_count"][p["all_insertion_left_positions"]] += c
            X = ext[r]
            eff_len = len(refs[r]["sequence"])
            if not params.ignore_insertions:
                X["inserted_n"][p["insertion_n"]] += c
                eff_len += p["insertion_n"]
                V["insertion_count"][p["insertion_positions"]] += c
                if p["insertion_n"] > 0:
                    S["counts_insertion"] += c
                    has_i = True
            V["all_deletion_count"][p["all_deletion_positions"]] += c
            if not params.ignore_deletions:
                X["deleted_n"][p["deletion_n"]] += c
                eff_len -= p["deletion_n"]
                V["deletion_count"][p["deletion_positions"]] += c
                if p["deletion_n"] > 0:
                    S["counts_deletion"] += c
                    has_d = True
            X["effective_len"][eff_len] += c
            V["all_substitution_count"][p["all_substitution_positions"]] += c
            if not params.ignore_substitutions:
                X["substituted_n"][p["substitution_n"]] += c
                V["substitution_count"][p["substitution_positions"]] += c
                if p["substitution_n"] > 0:
                    S["counts_substitution"] += c
                    has_s = True
                for pos, b in zip(p["all_substitution_positions"], p["all_substitution_values"]):
                    if b in "ATCGN":
                        V["all_substitution_base_" + b][pos] += c
            key = {(1, 1, 1): "counts_insertion_and_deletion_and_substitution",
                   (1, 1, 0): "counts_insertion_and_deletion", (0, 1, 1): "counts_deletion_and_substitution",
                   (0, 1, 0): "counts_only_deletion", (1, 0, 1): "counts_insertion_and_substitution",
                   (1, 0, 0): "counts_only_insertion", (0, 0, 1): "counts_only_substitution"}.get(
                       (int(has_i), int(has_d), int(has_s)))
            if key:
                S[key] += c
            for ch, rp in zip(p["aln_seq"], p["ref_positions"]):
                if rp >= 0:
                    V["all_base_count_" + ch][rp] += c
            # :4083-4180.  (The `elif tot_exon_len_mod != 0` arm at :4173 is unreachable: the `if` above it already
            # holds `or tot_exon_len_mod != 0`.)
            tem = sum(refs[r].get("exon_len_mods", []) or [])
            if has_i or has_d or has_s or tem != 0:
                coding = refs[r].get("contains_coding_seq", False)
                exons = set(refs[r].get("exon_positions", []))
                splice = set(refs[r].get("splicing_positions", []))
                len_mods = []
                exons_modified = spliced = False
                for (a, b), sz in zip(p["insertion_coordinates"], p["insertion_sizes"]):
                    V["insertion_length"][a] += sz * c
                    V["insertion_length"][b] += sz * c
                    if coding and exons.intersection((a, b)):
                        exons_modified = True
                        len_mods.append(sz)
                for (a, b), sz in zip(p["deletion_coordinates"], p["deletion_sizes"]):
                    V["deletion_length"][list(range(a, b))] += sz * c
                if coding:
                    hit = exons.intersection(p["deletion_positions"])
                    if hit:
                        exons_modified = True
                        len_mods.append(-len(hit))
                    if exons.intersection(p["substitution_positions"]):
                        exons_modified = True
                    if (splice.intersection(p["deletion_positions"]) or splice.intersection(p["insertion_positions"])
                            or splice.intersection(p["substitution_positions"])):
                        spliced = True
                    if spliced:
                        S["counts_splicing_sites_modified"] += c
                    if tem != 0:
                        eff = sum(len_mods) + tem
                        if eff % 3 == 0:
                            S["counts_modified_non_frameshift"] += c
                            X["hists_inframe"][eff] += c
                        else:
                            S["counts_modified_frameshift"] += c
                            X["hists_frameshift"][eff] += c
                    elif exons_modified:
                        if not len_mods:
                            S["counts_modified_non_frameshift"] += c
                            X["hists_inframe"][0] += c
                        else:
                            eff = sum(len_mods)
                            if eff % 3 == 0:
                                S["counts_modified_non_frameshift"] += c
                                X["hists_inframe"][eff] += c
                            else:
                                S["counts_modified_frameshift"] += c
                                X["hists_frameshift"][eff] += c
                    else:
                        S["counts_non_modified_non_frameshift"] += c
                        V["insertion_count_noncoding"][p["insertion_positions"]] += c
                        V["deletion_count_noncoding"][p["deletion_positions"]] += c
                        V["substitution_count_noncoding"][p["substitution_positions"]] += c
                        X["hists_inframe"][0] += c
    return vec, sca, classes, total


def ref1_vectors(cache, refs, ref_names, params):
    """HDR / prime-editing re-projection (CRISPRessoCORE.py:4195-4272); call AFTER count_vectors (it uses the counts left
    by the reverse-complement merge).  -> {ref: {name: float64 array over reference-0 positions}} for ref != ref_names[0]."""
    r0 = ref_names[0]
    L0 = len(refs[r0]["sequence"])
    names = ["ref1_all_insertion_count", "ref1_all_insertion_left_count", "ref1_all_deletion_count",
             "ref1_all_substitution_count"] + ["ref1_all_base_count_" + b for b in "ACGTN-"]
    out = {r: {n: np.zeros(L0) for n in names} for r in ref_names}
    for s, v in cache.items():
        c = v["count"]
        if c == 0:
            continue
        if len(v["aln_ref_names"]) == 1 and v["aln_ref_names"][0] == r0:
            continue
        if v["class_name"] == "AMBIGUOUS":
            continue
        _, s1, s2, _ = v["ref_aln_details"][0]
        if getattr(params, "use_legacy_insertion_quantification", False):          # CRISPRessoCORE.py:4244-4247
            p = find_indels_substitutions_legacy(s1, s2, refs[r0]["include_idxs"])
        else:
            p = find_indels_substitutions(s1, s2, refs[r0]["include_idxs"])
        for r in v["aln_ref_names"]:
            if r == r0:
                continue
            V = out[r]
            V["ref1_all_insertion_count"][p["all_insertion_positions"]] += c
            V["ref1_all_insertion_left_count"][p["all_insertion_left_positions"]] += c
            V["ref1_all_deletion_count"][p["all_deletion_positions"]] += c
            V["ref1_all_substitution_count"][p["all_substitution_positions"]] += c
            for ch, rp in zip(s1, p["ref_positions"]):
                if rp >= 0:
                    V["ref1_all_base_count_" + ch][rp] += c
    for r in ref_names:
        out[r]["ref1_all_indelsub_count"] = (out[r]["ref1_all_insertion_count"] + out[r]["ref1_all_deletion_count"]
                                             + out[r]["ref1_all_substitution_count"])
    return out


# --------------------------------------------------------------------------- compiled reference

def ref_modules():
    """The reference's own Cython modules compiled under oracle/_ref (None if absent)."""
    import importlib
    import sys
    root = os.path.join(_HERE, "_ref")
    if not os.path.isdir(os.path.join(root, "CRISPResso2")):
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        A = importlib.import_module("CRISPResso2.CRISPResso2Align")
        R = importlib.import_module("CRISPResso2.CRISPRessoCOREResources")
    except ImportError:
        return None
    return A, R
