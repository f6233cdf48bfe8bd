"""Host-side mirror of the reference's per-read driver for the GPU path.

Same call signatures as the reference functions they stand in for:
  process_fastq(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory)
      -> (aln_stats, not_aligned_variants)                        CRISPRessoCORE.py:1735-2000
  get_new_variant_object(args, fastq_seq, refs, ref_names, aln_matrix, pe_scaffold_dna_info)
      -> variant dict                                             CRISPRessoCORE.py:627-798
and, for the count arrays the quantification loop builds from variantCache (CRISPRessoCORE.py:3964-4115),
  quantify(variantCache) -> CountBlock        (the device block accumulated by the process_fastq call)

Everything decided per read (strand, alignment, best reference, classification, counts) is decided by the CUDA
kernel; this module parses the FASTQ, de-duplicates reads exactly like the reference (:1825-1849), derives the
reverse-complement merge weights of :3971-3975, launches one batch and re-labels the outputs as the reference's
dict / ResultsSlotsDict shapes.  process_fastq bypasses CRISPRessoMultiProcessing (n_processes is ignored) and
follows the SERIAL branch's statistics (:1956-1981).
"""
import os

import numpy as np

from . import _lib
from .align import read_matrix
from . import fastq
from .engine import Engine, EngineError, pack_reads
from .resources import payload_from_lists

_COMP = str.maketrans("ACGTN_-", "TGCAN_-")
_engines = {}
_blocks = {}


def reverse_complement(seq):
    """CRISPRessoShared.py:399-403 (KeyError on symbols outside ACGTN_- like the reference)."""
    up = seq.upper()
    bad = set(up) - set("ACGTN_-")
    if bad:
        raise KeyError(sorted(bad)[0])
    return up.translate(_COMP)[::-1]


def _flags(args):
    f = 0
    if getattr(args, "ignore_substitutions", False):
        f |= _lib.F_IGNORE_SUBSTITUTIONS
    if getattr(args, "ignore_insertions", False):
        f |= _lib.F_IGNORE_INSERTIONS
    if getattr(args, "ignore_deletions", False):
        f |= _lib.F_IGNORE_DELETIONS
    if getattr(args, "expand_ambiguous_alignments", False):
        f |= _lib.F_EXPAND_AMBIGUOUS
    if getattr(args, "assign_ambiguous_alignments_to_first_reference", False):
        f |= _lib.F_ASSIGN_FIRST
    if getattr(args, "discard_indel_reads", False):
        f |= _lib.F_DISCARD_INDEL_READS
    if getattr(args, "expected_hdr_amplicon_seq", "") or getattr(args, "prime_editing_pegRNA_extension_seq", ""):
        f |= _lib.F_HDR_REF1
    return f


def _unsupported(args):
    if getattr(args, "use_legacy_insertion_quantification", False):
        raise NotImplementedError("use_legacy_insertion_quantification is not built on the GPU path")
    if getattr(args, "prime_editing_pegRNA_scaffold_seq", ""):
        raise NotImplementedError("prime-editing scaffold search (CRISPRessoCORE.py:789-796) is not built on the GPU path")


def get_engine(device=0, lib_path=None):
    key = (device, lib_path)
    if key not in _engines:
        _engines[key] = Engine(device, lib_path)
    return _engines[key]


def configure_engine(engine, args, refs, ref_names, aln_matrix, edit_cap=24):
    engine.configure(refs, ref_names, aln_matrix, args.needleman_wunsch_gap_open, args.needleman_wunsch_gap_extend,
                     args.aln_seed_count, args.aln_seed_min, _flags(args), "ACGTN", edit_cap)
    return engine


def merge_weights(uniques, counts):
    """Weights the quantification loop ends up using after its reverse-complement merge (CRISPRessoCORE.py:3971-3975):
    walking the cache in first-seen order, a read absorbs the count of its reverse complement (which drops to 0);
    a palindromic read absorbs itself.  Valid because a read and its reverse complement always share their aligned
    status (same score set), so both are in the cache or neither is."""
    pos = {s: k for k, s in enumerate(uniques)}
    w = list(counts)
    for k, s in enumerate(uniques):
        if w[k] == 0:
            continue
        try:
            rc = reverse_complement(s)
        except KeyError:
            continue
        j = pos.get(rc)
        if j is not None and w[j] > 0:
            tot = w[k] + w[j]
            w[j] = 0
            w[k] = tot
    return w


class _BatchLists:
    """Plain-Python views of a batch's output arrays, converted with one .tolist() per block of reads: the per-read
    work then touches tuples instead of numpy scalars (about 5x less host time per unique read)."""
    BLOCK = 4096

    def __init__(self, res):
        self.res = res
        self.W, self.nref = res.W, res.alns.shape[1]
        self.lo = self.hi = 0
        self.text = None

    def block(self, i):
        if not (self.lo <= i < self.hi):
            self.lo = (i // self.BLOCK) * self.BLOCK
            self.hi = min(self.lo + self.BLOCK, len(self.res.recs))
            r = self.res
            self.recs = r.recs[self.lo:self.hi].tolist()     # (winner_mask, best_score_milli, best_ref, n_winners, ambiguous, status)
            self.alns = r.alns[self.lo:self.hi].tolist()     # [read][ref] -> ALN_DTYPE field order
            self.edits = r.edits[self.lo:self.hi].tolist() if r.edits is not None else None
            # one decode per block; bytes left of an alignment are undefined, hence latin-1 (never fails, 1 char per byte)
            self.text = r.strings[self.lo:self.hi].tobytes().decode("latin-1") if r.strings is not None else None
        return i - self.lo

    def pair(self, i, r, n):
        base = (((i - self.lo) * self.nref + r) * 2) * self.W
        return self.text[base + self.W - n: base + self.W], self.text[base + 2 * self.W - n: base + 2 * self.W]


# ALN_DTYPE field positions (crispresso2_b200/_lib.py)
(_A_NMATCH, _A_ALNLEN, _A_SCORE, _A_STRAND, _A_STATUS, _A_NEDITS, _A_INS_N, _A_DEL_N, _A_SUB_N, _A_NINS_ALL, _A_NINS_WIN,
 _A_NDEL_ALL, _A_NDEL_WIN, _A_NDELPOS_ALL, _A_NSUB_ALL, _A_IRR, _A_MOD) = range(17)


def _variant_from(res, i, seq, ref_names, refs):
    """dict of get_new_variant_object (CRISPRessoCORE.py:709-798) for read i of a batch result."""
    L = getattr(res, "_lists", None)
    if L is None:
        L = res._lists = _BatchLists(res)
    k = L.block(i)
    winner_mask, best_milli, _, _, ambiguous, _ = L.recs[k]
    nref = len(ref_names)
    alns = L.alns[k]
    scores = [alns[r][_A_SCORE] / 1000.0 for r in range(nref)]
    details = []
    for r in range(nref):
        s1, s2 = L.pair(i, r, alns[r][_A_ALNLEN])
        details.append((ref_names[r], s1, s2, scores[r]))
    v = {"count": 1, "aln_scores": scores, "ref_aln_details": details}
    if best_milli <= 0:
        v["best_match_score"] = -1
        return v
    winners = [r for r in range(nref) if (winner_mask >> r) & 1]
    v["aln_ref_names"] = [ref_names[r] for r in winners]
    v["best_match_score"] = best_milli / 1000.0
    labels = []
    for r in winners:
        a = alns[r]
        name = ref_names[r]
        s1, s2 = details[r][1], details[r][2]
        n = a[_A_NEDITS]
        if a[_A_STATUS] & _lib.ST_EDIT_OVERFLOW or L.edits is None or n > len(L.edits[k][r]):
            raise OverflowError("edit list overflow")
        p = payload_from_lists(a[_A_INS_N], a[_A_DEL_N], a[_A_SUB_N], L.edits[k][r][:n], s2)
        p.ref_name = name
        p.aln_scores = scores
        p.irregular_ends = bool(a[_A_IRR])
        n_ins_all, n_ins_win = a[_A_NINS_ALL], a[_A_NINS_WIN]
        p.insertions_outside_window = n_ins_all - n_ins_win
        p.deletions_outside_window = a[_A_NDEL_ALL] - a[_A_NDEL_WIN]
        p.substitutions_outside_window = a[_A_NSUB_ALL] - a[_A_SUB_N]
        p.total_mods = n_ins_all + a[_A_NDELPOS_ALL] + a[_A_NSUB_ALL]
        p.mods_in_window = a[_A_SUB_N] + a[_A_DEL_N] + a[_A_INS_N]
        p.mods_outside_window = p.total_mods - p.mods_in_window
        p.classification = "MODIFIED" if a[_A_MOD] else "UNMODIFIED"
        labels.append(name + "_" + p.classification)
        p.aln_seq, p.aln_ref = s1, s2
        p.aln_strand = "-" if a[_A_STRAND] else "+"
        v["variant_" + name] = p
        v["best_match_name"] = name
    v["class_name"] = "&".join(labels)
    if len(winners) > 1:
        if ambiguous:
            v["class_name"] = "AMBIGUOUS"
        elif len(labels) > 1 and not (res_flags(res) & _lib.F_EXPAND_AMBIGUOUS):
            v["class_name"] = labels[0]
            v["aln_ref_names"] = [ref_names[winners[0]]]
    return v


def res_flags(res):
    return getattr(res, "flags", 0)


def align_uniques(engine, uniques, counts, ref_names, refs, flags, weights=None, packed=None):
    """One GPU batch over unique reads -> (BatchResult, merge weights).  `packed` = (bytes, offsets) of `uniques`
    when the caller already holds them in the engine's layout."""
    weights = merge_weights(uniques, counts) if weights is None else weights
    buf, off = packed if packed is not None else pack_reads(uniques)
    res = engine.align_packed(buf, off, count=np.asarray(counts, dtype=np.int32), qweight=np.asarray(weights, dtype=np.int32))
    res.flags = flags
    st = res.recs["status"]
    hard = st & ~np.uint32(_lib.ST_EDIT_OVERFLOW)
    if hard.any():
        k = int(np.nonzero(hard)[0][0])
        raise EngineError("read %d (%r...) outside the engine's contract: status %d" % (k, uniques[k][:30], int(st[k])))
    return res, weights


def process_fastq(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory,
                  engine=None, aln_matrix=None):
    _unsupported(args)
    if aln_matrix is None:
        loc = args.needleman_wunsch_aln_matrix_loc
        if not os.path.isabs(loc) and not os.path.exists(loc):
            raise FileNotFoundError("needleman_wunsch_aln_matrix_loc %r not found (pass an absolute path)" % loc)
        aln_matrix = read_matrix(loc)
    engine = engine or get_engine()
    # FASTQ read + dedup of CRISPRessoCORE.py:1820-1849, done natively (c2b_fastq_dedup: threads, exact); the packed
    # unique sequences feed the batch call directly
    dd = fastq.dedup_file(fastq_filename, lib_path=engine.lib_path)
    packed = None
    if not variantCache:
        variantCache.update(zip(dd.uniques, dd.counts.tolist()))
        packed = (dd.buf, dd.off)
    else:                                                   # caller pre-seeded the cache: same += semantics, same key order
        for seq, c in zip(dd.uniques, dd.counts.tolist()):
            variantCache[seq] = variantCache.get(seq, 0) + c
    configure_engine(engine, args, refs, ref_names, aln_matrix)
    engine.counts_reset()
    uniques = list(variantCache.keys())
    counts = [variantCache[s] for s in uniques]
    flags = _flags(args)
    res, weights = align_uniques(engine, uniques, counts, ref_names, refs, flags, packed=packed)
    over = np.nonzero(res.recs["status"] & _lib.ST_EDIT_OVERFLOW)[0]
    fix = {}
    if len(over):
        # their counts are already in the block (only the edit LIST was truncated): re-run them with a cap that cannot
        # overflow and zero weights, just to fetch the complete lists
        cap0 = engine.edit_cap
        engine.set_edit_cap(max(engine.ref_lens) + _lib.MAX_READ_LEN)
        sub = [uniques[k] for k in over]
        zero = np.zeros(len(sub), dtype=np.int32)
        r2 = engine.align(sub, count=zero, qweight=zero)
        r2.flags = flags
        engine.set_edit_cap(cap0)
        fix = {int(k): (r2, j) for j, k in enumerate(over)}

    st = dict.fromkeys(["N_TOT_READS", "N_CACHED_ALN", "N_CACHED_NOTALN", "N_COMPUTED_ALN", "N_COMPUTED_NOTALN",
                        "N_GLOBAL_SUBS", "N_SUBS_OUTSIDE_WINDOW", "N_MODS_IN_WINDOW", "N_MODS_OUTSIDE_WINDOW",
                        "N_READS_IRREGULAR_ENDS", "READ_LENGTH"], 0)
    not_aligned = {}
    class_extra = {}
    for k, seq in enumerate(uniques):                       # CRISPRessoCORE.py:1956-1981
        c = counts[k]
        st["N_TOT_READS"] += c
        rr, kk = fix.get(k, (res, k))
        v = _variant_from(rr, kk, seq, ref_names, refs)
        v["count"] = c
        if v["best_match_score"] <= 0:
            st["N_COMPUTED_NOTALN"] += 1
            st["N_CACHED_NOTALN"] += c - 1
            not_aligned[seq] = v
            continue
        variantCache[seq] = v
        if "&" in v["class_name"] and weights[k] > 0:       # joined labels (--expand_ambiguous_alignments): host-side class count
            class_extra[v["class_name"]] = class_extra.get(v["class_name"], 0) + int(weights[k])
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
    for seq in not_aligned:
        del variantCache[seq]
    block = engine.counts()
    dev = block.aln_stats_partial()                         # the kernel's own sums must agree with the records
    for key, val in dev.items():
        if val != st[key]:
            raise EngineError("device aln_stats disagree with per-read records for %s: %d != %d" % (key, val, st[key]))
    block.class_extra = class_extra
    _blocks[id(variantCache)] = block
    return st, not_aligned


def process_fastq_sharded(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory,
                          engine=None, aln_matrix=None, group=None):
    """process_fastq for one process per GPU (torch.distributed initialised): every rank reads and de-duplicates the FASTQ,
    aligns its contiguous slice of the unique reads (the reference's own sharding rule, CRISPRessoCORE.py:1172-1195), then
    the count blocks are summed with ONE all-reduce (NCCL on the device block; gloo on the CPU test path) and the per-read
    variants and statistics are gathered, so that every rank returns exactly what the single-process call returns --
    same variantCache (keys in first-seen order), same aln_stats, same count block behind quantify()."""
    import torch.distributed as dist
    from . import dist as cdist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return process_fastq(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory,
                             engine=engine, aln_matrix=aln_matrix)
    _unsupported(args)
    if aln_matrix is None:
        aln_matrix = read_matrix(args.needleman_wunsch_aln_matrix_loc)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    engine = engine or get_engine()
    dd = fastq.dedup_file(fastq_filename, lib_path=engine.lib_path)
    for seq, c in zip(dd.uniques, dd.counts.tolist()):
        variantCache[seq] = variantCache.get(seq, 0) + c
    configure_engine(engine, args, refs, ref_names, aln_matrix)
    engine.counts_reset()
    uniques = list(variantCache.keys())
    counts = [variantCache[s] for s in uniques]
    flags = _flags(args)
    weights = merge_weights(uniques, counts)                # needs the global unique table: before sharding
    lo, hi = cdist.shard_bounds(len(uniques), rank, world)
    mine = uniques[lo:hi]
    local = []                                              # (variant or None) per unique of this shard
    if mine:
        res, _ = align_uniques(engine, mine, counts[lo:hi], ref_names, refs, flags, weights=weights[lo:hi])
        over = np.nonzero(res.recs["status"] & _lib.ST_EDIT_OVERFLOW)[0]
        fix = {}
        if len(over):
            cap0 = engine.edit_cap
            engine.set_edit_cap(max(engine.ref_lens) + _lib.MAX_READ_LEN)
            zero = np.zeros(len(over), dtype=np.int32)
            r2 = engine.align([mine[k] for k in over], count=zero, qweight=zero)
            r2.flags = flags
            engine.set_edit_cap(cap0)
            fix = {int(k): (r2, j) for j, k in enumerate(over)}
        for k, seq in enumerate(mine):
            rr, kk = fix.get(k, (res, k))
            v = _variant_from(rr, kk, seq, ref_names, refs)
            v["count"] = counts[lo + k]
            local.append(v)
    merged = cdist.allreduce_counts(engine, group)          # the path's one collective on device data
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    st = dict.fromkeys(["N_TOT_READS", "N_CACHED_ALN", "N_CACHED_NOTALN", "N_COMPUTED_ALN", "N_COMPUTED_NOTALN",
                        "N_GLOBAL_SUBS", "N_SUBS_OUTSIDE_WINDOW", "N_MODS_IN_WINDOW", "N_MODS_OUTSIDE_WINDOW",
                        "N_READS_IRREGULAR_ENDS", "READ_LENGTH"], 0)
    not_aligned, class_extra = {}, {}
    k = 0
    for part in parts:                                      # rank order = unique order: the serial loop of :1956-1981
        for v in part:
            seq, c = uniques[k], counts[k]
            st["N_TOT_READS"] += c
            if v["best_match_score"] <= 0:
                st["N_COMPUTED_NOTALN"] += 1
                st["N_CACHED_NOTALN"] += c - 1
                not_aligned[seq] = v
            else:
                variantCache[seq] = v
                if "&" in v["class_name"] and weights[k] > 0:
                    class_extra[v["class_name"]] = class_extra.get(v["class_name"], 0) + int(weights[k])
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
            k += 1
    assert k == len(uniques)
    for seq in not_aligned:
        del variantCache[seq]
    block = engine.counts(raw=merged)
    for key, val in block.aln_stats_partial().items():
        if val != st[key]:
            raise EngineError("device aln_stats disagree with per-read records for %s: %d != %d" % (key, val, st[key]))
    block.class_extra = class_extra
    _blocks[id(variantCache)] = block
    return st, not_aligned


def quantify(variantCache):
    """Count block accumulated on the device by the process_fastq call that filled this variantCache."""
    return _blocks[id(variantCache)]


def get_new_variant_object(args, fastq_seq, refs, ref_names, aln_matrix, pe_scaffold_dna_info=None, engine=None):
    _unsupported(args)
    engine = engine or get_engine()
    configure_engine(engine, args, refs, ref_names, aln_matrix, edit_cap=max(len(refs[r]["sequence"]) for r in ref_names)
                     + len(fastq_seq) + 1)
    res, _ = align_uniques(engine, [fastq_seq], [1], ref_names, refs, _flags(args), weights=[0])
    return _variant_from(res, 0, fastq_seq, ref_names, refs)
