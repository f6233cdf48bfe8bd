"""CPU-only parity tests of the KERNEL LOGIC: crispresso2_b200/csrc/c2b_core.cuh compiled with g++ against the
fiber warp emulator (tests/emu/), driven through the same C ABI and the same Python host code as the GPU build.
These prove the algorithm (wavefront DP, tagged tie-breaks, traceback window, row-space classification, count
block) against the reference goldens before any GPU time is spent; tests/test_gpu_parity.py repeats them on the
real sm_100a build."""
import numpy as np
import pytest

import golden_util as G
import parity_util as PU
from crispresso2_b200 import _lib, synth
from crispresso2_b200.engine import Engine, pack_reads
from oracle import oracle as O


@pytest.fixture(scope="module")
def emu():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    return Engine(lib_path=build_emu.build())


@pytest.mark.parametrize("case", ["fanc_cas9", "fanc_params", "synth_hdr", "fanc_pe_scaffold"])
def test_golden_whole_path(emu, case, tmp_path):
    PU.check_golden_case(emu, case, tmp_path)


def test_golden_synth_single_subset(emu, tmp_path):
    PU.check_golden_case(emu, "synth_single", tmp_path, max_reads=300)


def test_align_vectors(emu):
    import gzip, json, os
    from crispresso2_b200 import align, resources
    with gzip.open(os.path.join(G.GOLD, "align_vectors.json.gz"), "rt") as fh:
        cases = json.load(fh)
    m = O.make_matrix()
    for c in cases[::3]:
        got = align.global_align(c["read"], c["ref"], m, np.array(c["gi"], dtype=np.int64), c["go"], c["ge"], engine=emu)
        assert got == (c["s1"], c["s2"], c["score"]), c
        a, ed = emu.classify_pair(c["s1"], c["s2"], c["inc"])
        p = resources.payload_from_device(a, ed, c["s1"], c["s2"])
        assert not G.payload_equal(c["payload"], p), (c, p.__dict__)


@pytest.mark.parametrize("flags", [{}, {"ignore_substitutions": True}, {"discard_indel_reads": True},
                                   {"ignore_deletions": True, "ignore_insertions": True}])
def test_seeded_batch_against_oracle(emu, flags):
    rng = np.random.default_rng(11)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    reads = synth.synth_reads(rng, amp, 150, 250, sub_rate=0.01, rc_frac=0.08, n_rate=0.002, cut=ref["cut_point"])
    reads = [r.tobytes().decode() for r in reads] + ["".join(rng.choice(list("ACGT"), 250)) for _ in range(3)]
    PU.check_against_oracle(emu, {"Reference": ref}, ["Reference"], O.Params(**flags), reads, O.make_matrix())


def test_ragged_and_long_inputs_against_oracle(emu):
    """Read lengths 30..400, a 300-bp amplicon (two row blocks), a wide quantification window."""
    rng = np.random.default_rng(5)
    amp = synth.random_amplicon(rng, 300)
    ref = synth.amplicon_setup(amp, guide_start=140, window_size=20)
    reads = []
    for _ in range(60):
        L = int(rng.integers(30, 401))
        s = synth.synth_reads(rng, amp, 1, 300, sub_rate=0.02, cut=ref["cut_point"])[0].tobytes().decode()
        s = (s + "".join(rng.choice(list("ACGT"), 120)))[:L]
        reads.append(s)
    PU.check_against_oracle(emu, {"Reference": ref}, ["Reference"], O.Params(), reads, O.make_matrix())


def test_three_amplicons_and_ambiguity_flags(emu):
    rng = np.random.default_rng(3)
    amp = synth.random_amplicon(rng, 120)
    hdr = amp[:60] + "TGA" + amp[63:68] + "ACGTAC" + amp[68:]
    snp = list(amp)
    for p in (20, 50, 70, 100):
        snp[p] = "A" if snp[p] != "A" else "C"
    snp = "".join(snp)
    refs = {"WT": synth.amplicon_setup(amp, guide_start=45), "HDR": synth.amplicon_setup(hdr, guide_start=45),
            "SNP": synth.amplicon_setup(snp, guide_start=45), "WT2": synth.amplicon_setup(amp, guide_start=45)}
    names = ["WT", "HDR", "SNP", "WT2"]
    reads = []
    for a in (amp, hdr, snp):
        reads += [r.tobytes().decode() for r in synth.synth_reads(rng, a, 40, 120, sub_rate=0.01, rc_frac=0.1, cut=62)]
    for kw in ({}, {"expand_ambiguous_alignments": True}, {"assign_ambiguous_alignments_to_first_reference": True},
               {"expected_hdr_amplicon_seq": hdr}, {"expected_hdr_amplicon_seq": hdr, "expand_ambiguous_alignments": True}):
        PU.check_against_oracle(emu, refs, names, O.Params(**kw), reads, O.make_matrix())


def test_bad_symbols_and_empty(emu):
    rng = np.random.default_rng(1)
    amp = synth.random_amplicon(rng, 120)
    ref = synth.amplicon_setup(amp, guide_start=50)
    emu.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2)
    assert len(emu.align([]).recs) == 0
    res = emu.align([amp, amp[:50] + "x" + amp[51:]])
    assert res.recs["status"][0] == 0 and res.recs["status"][1] == _lib.ST_BAD_CHAR


def test_packed_pair_path_equals_32bit_path(emu):
    rng = np.random.default_rng(77)
    amp = synth.random_amplicon(rng, 180)
    ref = synth.amplicon_setup(amp, guide_start=80)
    base = synth.synth_reads(rng, amp, 120, 180, sub_rate=0.03, rc_frac=0.1, n_rate=0.004, cut=ref["cut_point"])
    reads = [r.tobytes().decode() for r in base]
    for k in range(0, len(reads), 5):
        reads[k] = reads[k][: 60 + (k % 100)]
    out = []
    for flags in (0, _lib.F_NO_PAIRING):
        emu.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, flags, "ACGTN", 40)
        emu.counts_reset()
        res = emu.align(reads)
        out.append((res, emu.counts_raw(), emu.path_counts()))
    (a, ca, pa), (b, cb, pb) = out
    assert pa[0] > 10 and pb[0] == 0
    assert (a.recs == b.recs).all() and (a.alns == b.alns).all() and (ca == cb).all()
    for i in range(len(reads)):
        assert a.pair(i) == b.pair(i)
    (ea, fa), (eb, fb) = PU.edits_canonical(a), PU.edits_canonical(b)
    assert fa.all() and fb.all() and (ea == eb).all()


@pytest.mark.parametrize("I", [250, 131])
def test_ring_banded_path_equals_full_matrix(emu, I):
    PU.check_ring_equals_full(emu, n=96, I=I, seed=40 + I, oracle_subset=48)


def test_leftover_list_with_single_reads_and_odd_tail(emu):
    PU.check_leftover_singles(emu)
    PU.check_leftover_singles(emu, n=9, seed=3)


def test_pairs_longer_than_512_columns(emu):
    PU.check_long_pairs(emu, n=16)


def test_legacy_insertion_quantification(emu):
    PU.check_legacy(emu)


def test_narrow_first_tier_equals_the_wide_ring(emu):
    PU.check_narrow_equals_wide(emu, n=160, oracle_subset=64)
    PU.check_narrow_equals_wide(emu, n=96, I=131, seed=7)


def test_pooled_ref_id(emu):
    PU.check_pooled(emu, n_amplicons=4, reads_per=24)
    PU.check_pooled(emu, n_amplicons=40, reads_per=3, seed=5)          # more references than C2B_MAX_REFS: Pooled only


def test_chunked_pipeline_equals_one_chunk(emu, monkeypatch):
    """c2b_align_batch splits a batch into pipelined chunks (uneven first/last); results must not depend on it."""
    rng = np.random.default_rng(8)
    amp = synth.random_amplicon(rng, 100)
    ref = synth.amplicon_setup(amp, guide_start=40)
    reads = [r.tobytes().decode() for r in synth.synth_reads(rng, amp, 61, 100, sub_rate=0.02, cut=ref["cut_point"])]
    reads[5] = reads[5][:70]
    out = []
    # (chunk size, bounce): the pinned bounce buffers of callers with pageable arrays, forced on and off
    for chunk, bounce, compact in ((None, "0", False), ("7", "0", False), ("7", "1", False), ("9", "1", True), (None, "1", True)):
        if chunk:
            monkeypatch.setenv("C2B_CHUNK", chunk)
        else:
            monkeypatch.delenv("C2B_CHUNK", raising=False)
        monkeypatch.setenv("C2B_FORCE_BOUNCE", bounce)
        emu.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", 16)
        emu.counts_reset()
        buf, off = pack_reads(reads)
        res = emu.align_packed(buf, off, compact=compact)
        out.append((res, emu.counts_raw()))
    monkeypatch.delenv("C2B_CHUNK", raising=False)
    monkeypatch.delenv("C2B_FORCE_BOUNCE", raising=False)
    a, ca = out[0]
    for b, cb in out[1:]:
        assert (a.recs == b.recs).all() and (a.alns == b.alns).all() and (ca == cb).all()
        (ea, fa), (eb, fb) = PU.edits_canonical(a), PU.edits_canonical(b)      # entries past n_edits are undefined
        assert (fa == fb).all() and (ea[fa] == eb[fb]).all()
        for i in range(len(reads)):
            assert a.pair(i) == b.pair(i)


def test_chunked_pipeline_with_per_read_amplicon(emu, monkeypatch):
    """Pooled batches (compact [read][0] outputs) through several pipelined chunks."""
    monkeypatch.setenv("C2B_CHUNK", "10")
    try:
        PU.check_pooled(emu, n_amplicons=5, reads_per=9, seed=33)
    finally:
        monkeypatch.delenv("C2B_CHUNK", raising=False)


def test_long_amplicon_three_row_blocks(emu):
    """A 610-bp amplicon (three 256-row blocks on the 32-bit path) with 150..400-bp reads."""
    rng = np.random.default_rng(12)
    amp = synth.random_amplicon(rng, 610)
    ref = synth.amplicon_setup(amp, guide_start=300, window_size=5)
    reads = []
    for k in range(10):
        L = int(rng.integers(150, 401))
        s0 = int(rng.integers(0, 610 - L))
        s = synth.synth_reads(rng, amp, 1, 610, sub_rate=0.02, cut=ref["cut_point"])[0].tobytes().decode()
        reads.append(s[s0:s0 + L])
    reads.append(amp[100:500])
    PU.check_against_oracle(emu, {"Reference": ref}, ["Reference"], O.Params(), reads, O.make_matrix())


def test_banded_slab_falls_back_to_full_slab(emu):
    PU.check_band_fallback(emu, n=24)


def test_coding_seq_frameshift_splicing_and_size_histograms(emu):
    PU.check_coding_seq(emu, n_reads=60)


def test_small_odd_batches_keep_the_warp_convergent(emu):
    """Tiny, ragged batches (odd read counts, mixed amplicons inside a work group, mixed lengths): the emulator aborts on
    any collective that not every lane reaches, which on the GPU would be undefined behaviour."""
    for seed in range(6):
        PU.check_pooled(emu, n_amplicons=2 + seed, reads_per=1 + seed % 4, seed=100 + seed)
    rng = np.random.default_rng(2)
    amp = synth.random_amplicon(rng, 150)
    ref = synth.amplicon_setup(amp, guide_start=60)
    for n in (1, 2, 3, 5, 7, 9, 15, 17):
        reads = [r.tobytes().decode() for r in synth.synth_reads(rng, amp, n, 150, sub_rate=0.02, rc_frac=0.2, cut=ref["cut_point"])]
        for k in range(0, n, 3):
            reads[k] = reads[k][: 40 + 13 * k]
        PU.check_against_oracle(emu, {"Reference": ref}, ["Reference"], O.Params(), reads, O.make_matrix())


def test_random_configurations_against_oracle(emu):
    for seed in range(20):
        PU.check_random_config(emu, seed)


def test_batch_gate_matches_and_catches_corruption(emu):
    """oracle/batch_gate.py (the >=100k-read gate bench.py and the GPU tests run) on a small emulator batch: a clean batch
    passes on every field and on the count block; a batch with one flipped output byte / one changed scalar is caught."""
    from oracle import batch_gate as BG
    rng = np.random.default_rng(3)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    refs, names = {"Reference": ref}, ["Reference"]
    reads = synth.synth_reads(rng, amp, 96, 250, sub_rate=0.01, rc_frac=0.1, n_rate=0.002, cut=ref["cut_point"])
    emu.configure(refs, names, O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", 64)
    emu.counts_reset()
    off = np.arange(len(reads) + 1, dtype=np.int64) * 250
    res = emu.align_packed(reads.reshape(-1), off)
    summ, quant = BG.run(reads.reshape(-1), off, refs, names, O.Params(), O.make_matrix(), res.recs, res.alns, res.strings,
                         res.edits, res.W, n_workers=2)
    assert summ["n"] == 96 and summ["n_bad"] == 0, summ
    assert BG.compare_block(emu.counts(), quant[None], names) == []
    k = int(np.nonzero(res.recs["best_score_milli"] > 0)[0][5])
    s2 = res.strings.copy()
    s2[k, 0, 0, res.W - 3] = ord("A") if s2[k, 0, 0, res.W - 3] != ord("A") else ord("C")
    summ2, _ = BG.run(reads.reshape(-1), off, refs, names, O.Params(), O.make_matrix(), res.recs, res.alns, s2, res.edits,
                      res.W, n_workers=2)
    assert summ2["n_bad"] == 1
    a2 = res.alns.copy()
    a2[k, 0]["substitution_n"] += 1
    summ3, _ = BG.run(reads.reshape(-1), off, refs, names, O.Params(), O.make_matrix(), res.recs, a2, res.strings, res.edits,
                      res.W, n_workers=2)
    assert summ3["n_bad"] == 1


def test_seed_tests_disagreeing_across_references(emu):
    assert PU.check_seed_disagreement(emu, n=768) > 20
