"""Paired-end merge mode (`--crispresso_merge`): the reference's process_paired_fastq (CRISPRessoCORE.py:1245-1733) with its
Needleman-Wunsch calls served from ONE GPU batch.

get_new_variant_object_from_paired (:987-1169) aligns mate 1 and the reverse-complemented mate 2 against every amplicon, on
the forward strand and / or as reverse complements (2-4 `CRISPResso2Align.global_align` calls per amplicon, :1035-1053), then
merges the two alignments column by column with the base qualities (get_consensus_alignment_from_pairs, :829-985).  The dynamic
programming is what costs; everything after it is the reference's own bookkeeping, which this module leaves to the reference:

  1. both FASTQ files are parsed and de-duplicated natively (fastq.dedup_file);
  2. every distinct mate sequence and its reverse complement is aligned forward-only against all amplicons in one batch
     (C2B_F_NO_STRAND_SEARCH: the strand logic stays with the caller, exactly as global_align has none);
  3. `CRISPResso2Align.global_align` is re-bound, for the duration of the call, to a lookup into that batch (a sequence the
     batch does not hold -- there should be none -- is aligned by a single GPU call, never on the CPU);
  4. get_consensus_alignment_from_pairs (:829-985), the per-column merge of the two mates' alignments, is re-bound to its native
     restatement (c2b_consensus_from_pairs; the reference's own unit test for it and a differential fuzz pin it);
  5. the reference's own process_paired_fastq runs unchanged on top.

Same arguments, same return value, same variantCache as the reference; parity is by construction as long as global_align's
results are the reference's (tests/test_reference_unit_tests.py, tests/test_gpu_parity.py).
"""
import numpy as np

from . import _lib, align, fastq
from .engine import pack_reads

_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANTGCAN")


def get_consensus_alignment_from_pairs(aln_seq_r1, aln_ref_r1, score_r1, qual_r1, aln_seq_r2, aln_ref_r2, score_r2, qual_r2, lib_path=None):
    """CRISPRessoCORE.get_consensus_alignment_from_pairs (:829-985), natively (c2b_consensus_from_pairs).
    -> (final_aln, final_qual, final_ref, homology score, caching_is_ok)"""
    import ctypes as C
    L = _lib.load(lib_path)
    parts = (aln_seq_r1, aln_ref_r1, qual_r1, aln_seq_r2, aln_ref_r2, qual_r2)
    if not all(p.isascii() for p in parts):
        raise ValueError("get_consensus_alignment_from_pairs: non-ASCII sequence or quality string")
    b = [p.encode("ascii") for p in parts]
    cap = len(b[1]) + len(b[4]) + 1
    out = [C.create_string_buffer(cap) for _ in range(3)]
    n, nq, m, ok = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    rc = L.c2b_consensus_from_pairs(b[0], len(b[0]), b[1], len(b[1]), float(score_r1), b[2], len(b[2]),
                                    b[3], len(b[3]), b[4], len(b[4]), float(score_r2), b[5], len(b[5]),
                                    out[0], out[1], out[2], cap, C.byref(n), C.byref(nq), C.byref(m), C.byref(ok))
    if rc in (_lib.E_LIMIT, _lib.E_STATE):
        raise IndexError("string index out of range")
    if rc != 0:
        raise RuntimeError("c2b_consensus_from_pairs failed (%d)" % rc)
    cols = n.value
    return (out[0].raw[:cols].decode("ascii"), out[1].raw[:nq.value].decode("ascii"), out[2].raw[:cols].decode("ascii"),
            round(float(100 * m.value / float(cols)), 3), bool(ok.value))


class AlignmentMemo:
    """(read, amplicon) -> (aligned read, aligned amplicon, score), answered from one forward-only batch."""

    def __init__(self, engine, seqs, refs, ref_names, aln_matrix, gap_open, gap_extend):
        self.engine, self.matrix, self.go, self.ge = engine, aln_matrix, int(gap_open), int(gap_extend)
        self.ref_index = {refs[r]["sequence"]: k for k, r in enumerate(ref_names)}
        self.gi = [refs[r]["gap_incentive"] for r in ref_names]
        self.index = {s: k for k, s in enumerate(seqs)}
        self.hits = self.misses = 0
        engine.configure(refs, ref_names, aln_matrix, gap_open, gap_extend, 0, 0, _lib.F_NO_STRAND_SEARCH, "ACGTN", 0)
        buf, off = pack_reads([s.encode() for s in seqs])
        zero = np.zeros(len(seqs), dtype=np.int32)
        self.res = engine.align_packed(buf, off, count=zero, qweight=zero, edits=False, compact=True) if seqs else None
        if self.res is not None:
            st = self.res.recs["status"] & ~np.uint32(_lib.ST_EDIT_OVERFLOW)
            self.ok = st == 0
        self.cache = {}

    def global_align(self, pystr_seqj, pystr_seqi, matrix=None, gap_incentive=None, gap_open=-1, gap_extend=-1):
        """Signature of CRISPResso2Align.global_align (Align.pyx:101)."""
        k = self.index.get(pystr_seqj)
        r = self.ref_index.get(pystr_seqi)
        if (k is None or r is None or not self.ok[k] or int(gap_open) != self.go or int(gap_extend) != self.ge
                or matrix is not self.matrix or not (gap_incentive is self.gi[r] or np.array_equal(gap_incentive, self.gi[r]))):
            self.misses += 1
            return align.global_align(pystr_seqj, pystr_seqi, matrix, gap_incentive, gap_open, gap_extend, engine=None)
        self.hits += 1
        key = (k, r)
        out = self.cache.get(key)
        if out is None:
            s1, s2 = self.res.pair(k, r)
            out = self.cache[key] = (s1, s2, self.res.score(k, r))
        return out


def mate_sequences(fastq1_filename, fastq2_filename, lib_path=None):
    """Every sequence process_paired_fastq can hand to global_align: mate 1 as read (:1549) and reverse-complemented (:1041,
    :1049); mate 2 reverse-complemented -- upper-cased by reverse_complement -- (:1552) and that reverse-complemented again."""
    out, seen = [], set()

    def add(s):
        if s and s not in seen and len(s) <= _lib.MAX_READ_LEN and not (set(s) - set("ACGTN")):
            seen.add(s)
            out.append(s)

    for path, second in ((fastq1_filename, False), (fastq2_filename, True)):
        dd = fastq.dedup_file(path, lib_path=lib_path)
        for s in dd.uniques:
            raw = s.encode("utf-8", errors="surrogateescape")
            rc = raw.translate(_COMP)[::-1].decode("latin-1")
            if second:
                add(rc)
                add(s.upper())
            else:
                add(s)
                add(rc)
    return out


def process_paired_fastq(original, CRISPResso2Align, engine, fastq1_filename, fastq2_filename, variantCache, ref_names, refs, args,
                         files_to_remove, output_directory, fastq_write_out_file=None, aln_matrix=None):
    """`original`: the reference's process_paired_fastq; `CRISPResso2Align`: the module whose global_align it calls."""
    seqs = mate_sequences(fastq1_filename, fastq2_filename, lib_path=engine.lib_path)
    memo = AlignmentMemo(engine, seqs, refs, ref_names, aln_matrix, args.needleman_wunsch_gap_open, args.needleman_wunsch_gap_extend)
    saved_ga, saved_rm = CRISPResso2Align.global_align, CRISPResso2Align.read_matrix
    CRISPResso2Align.global_align = memo.global_align
    CRISPResso2Align.read_matrix = lambda path: aln_matrix          # the matrix object the memo was built with (:1277)
    import functools
    import sys
    host = sys.modules[original.__module__]                         # the module whose globals the reference's loop resolves
    saved_cons = getattr(host, "get_consensus_alignment_from_pairs", None)
    if saved_cons is not None:
        host.get_consensus_alignment_from_pairs = functools.partial(get_consensus_alignment_from_pairs, lib_path=engine.lib_path)
    try:
        n_proc, args.n_processes = args.n_processes, "1"            # the serial branch: the lookups live in this process
        try:
            if fastq_write_out_file is None:
                out = original(fastq1_filename, fastq2_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory)
            else:
                out = original(fastq1_filename, fastq2_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory,
                               fastq_write_out_file)
        finally:
            args.n_processes = n_proc
    finally:
        CRISPResso2Align.global_align, CRISPResso2Align.read_matrix = saved_ga, saved_rm
        if saved_cons is not None:
            host.get_consensus_alignment_from_pairs = saved_cons
    process_paired_fastq.last_memo = memo
    return out
