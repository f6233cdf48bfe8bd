"""Parity tests proper: the sm_100a build of the engine, called through the C ABI, against the reference goldens,
the oracle on seeded inputs, and size-independent properties at the benchmark's full size."""
import gzip
import json
import os

import numpy as np
import pytest

import golden_util as G
import parity_util as PU
from crispresso2_b200 import _lib, align, resources, synth
from crispresso2_b200.engine import Engine, pack_reads
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return Engine(0)


@pytest.mark.parametrize("case", G.CASES)
def test_golden_whole_path(eng, case, tmp_path):
    PU.check_golden_case(eng, case, tmp_path)


def test_align_vectors_and_classify_kats(eng):
    with gzip.open(os.path.join(G.GOLD, "align_vectors.json.gz"), "rt") as fh:
        cases = json.load(fh)
    m = O.make_matrix()
    for c in cases:
        got = align.global_align(c["read"], c["ref"], m, np.array(c["gi"], dtype=np.int64), c["go"], c["ge"], engine=eng)
        assert got == (c["s1"], c["s2"], c["score"]), c
        a, ed = eng.classify_pair(c["s1"], c["s2"], c["inc"])
        assert not G.payload_equal(c["payload"], resources.payload_from_device(a, ed, c["s1"], c["s2"]))


def test_reference_kats_through_the_dropin_names(eng):
    """The reference's own known-answer tests (tests/unit_tests/test_CRISPResso2Align.py) read the same here."""
    m = O.make_matrix()
    Z = lambda n: np.zeros(n, dtype=np.int64)
    assert align.global_align("ATTA", "ATTA", matrix=m, gap_incentive=Z(5), engine=eng) == ("ATTA", "ATTA", 100.0)
    assert align.global_align("ANNG", "ATCG", matrix=m, gap_incentive=Z(5), engine=eng) == ("A-NNG", "ATC-G", 40.0)
    assert align.global_align("AAAA", "TTTT", matrix=m, gap_incentive=Z(5), engine=eng) == ("---AAAA", "TTTT---", 0.0)
    assert align.global_align("A", "A", matrix=m, gap_incentive=Z(2), engine=eng) == ("A", "A", 100.0)
    gi = Z(6); gi[2] = 1
    assert align.global_align("ATTA", "ATTTA", matrix=m, gap_incentive=gi, engine=eng)[:2] == ("AT-TA", "ATTTA")
    assert align.global_align("ATTA", "ATTTA", matrix=m, gap_incentive=Z(5), engine=eng) == 0      # length mismatch


@pytest.mark.parametrize("flags", [{}, {"ignore_substitutions": True}, {"discard_indel_reads": True},
                                   {"ignore_deletions": True, "ignore_insertions": True}])
def test_seeded_batch_against_oracle(eng, flags):
    rng = np.random.default_rng(11)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    reads = synth.synth_reads(rng, amp, 1200, 250, sub_rate=0.01, rc_frac=0.08, n_rate=0.002, cut=ref["cut_point"])
    reads = [r.tobytes().decode() for r in reads] + ["".join(rng.choice(list("ACGT"), 250)) for _ in range(10)]
    PU.check_against_oracle(eng, {"Reference": ref}, ["Reference"], O.Params(**flags), reads, O.make_matrix())


def test_ragged_and_long_inputs_against_oracle(eng):
    """Read lengths 30..400, a 300-bp amplicon (two row blocks), a wide quantification window."""
    rng = np.random.default_rng(5)
    amp = synth.random_amplicon(rng, 300)
    ref = synth.amplicon_setup(amp, guide_start=140, window_size=20)
    reads = []
    for _ in range(300):
        L = int(rng.integers(30, 401))
        s = synth.synth_reads(rng, amp, 1, 300, sub_rate=0.02, cut=ref["cut_point"])[0].tobytes().decode()
        s = (s + "".join(rng.choice(list("ACGT"), 120)))[:L]
        reads.append(s)
    PU.check_against_oracle(eng, {"Reference": ref}, ["Reference"], O.Params(), reads, O.make_matrix())


def test_three_amplicons_and_ambiguity_flags(eng):
    rng = np.random.default_rng(3)
    amp = synth.random_amplicon(rng, 200)
    hdr = amp[:100] + "TGA" + amp[103:108] + "ACGTAC" + amp[108:]
    snp = list(amp)
    for p in (20, 60, 110, 150, 180):
        snp[p] = "A" if snp[p] != "A" else "C"
    snp = "".join(snp)
    refs = {"WT": synth.amplicon_setup(amp, guide_start=85), "HDR": synth.amplicon_setup(hdr, guide_start=85),
            "SNP": synth.amplicon_setup(snp, guide_start=85), "WT2": synth.amplicon_setup(amp, guide_start=85)}
    names = ["WT", "HDR", "SNP", "WT2"]       # WT2 == WT: every WT read is ambiguous
    reads = []
    for a in (amp, hdr, snp):
        reads += [r.tobytes().decode() for r in synth.synth_reads(rng, a, 150, 200, sub_rate=0.01, rc_frac=0.1, cut=102)]
    for kw in ({}, {"expand_ambiguous_alignments": True}, {"assign_ambiguous_alignments_to_first_reference": True},
               {"expected_hdr_amplicon_seq": hdr}, {"expected_hdr_amplicon_seq": hdr, "expand_ambiguous_alignments": True},
               {"expected_hdr_amplicon_seq": hdr, "assign_ambiguous_alignments_to_first_reference": True}):
        PU.check_against_oracle(eng, refs, names, O.Params(**kw), reads, O.make_matrix())


def test_full_size_properties(eng):
    """1M x 250 bp (the benchmark workload): properties that need no oracle run."""
    rng = np.random.default_rng(42)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    n = 1 << 20
    reads = synth.synth_reads_fast(rng, amp, n, 250, cut=ref["cut_point"])
    eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", 8)
    eng.counts_reset()
    off = np.arange(n + 1, dtype=np.int64) * 250
    res = eng.align_packed(reads.reshape(-1), off)
    a = res.alns[:, 0]
    assert (res.recs["status"] & ~np.uint32(_lib.ST_EDIT_OVERFLOW)).max() == 0
    # (1) stripping gaps from the aligned read gives back the read; from the aligned reference, the amplicon
    S = res.strings[:, 0]
    W = res.W
    cols = np.arange(W)[None, :] >= (W - a["aln_len"].astype(np.int64))[:, None]
    rd = S[:, 0, :]
    rf = S[:, 1, :]
    keep_r = cols & (rd != ord("-"))
    assert (keep_r.sum(1) == 250).all()
    assert (rd[keep_r].reshape(n, 250) == reads).all()
    keep_f = cols & (rf != ord("-"))
    assert (keep_f.sum(1) == 250).all()
    assert (rf[keep_f].reshape(n, 250) == np.frombuffer(amp.encode(), dtype=np.uint8)[None, :]).all()
    # (2) match count and score recomputed from the strings
    m = (cols & (rd == rf)).sum(1)
    assert (m == a["n_match"]).all()
    # (3) no column holds two gaps; an insertion column never touches a deletion column
    assert not (cols & (rd == ord("-")) & (rf == ord("-"))).any()
    ins, dele = cols & (rf == ord("-")), cols & (rd == ord("-"))
    assert not (ins[:, 1:] & dele[:, :-1]).any() and not (ins[:, :-1] & dele[:, 1:]).any()
    # (4) count block: every position is covered exactly once per counted read
    blk = eng.counts()
    V = blk.vectors("Reference")
    tot = blk.scalar("Reference", "TOTAL")
    aligned = res.recs["best_score_milli"] > 0
    overflow = (res.recs["status"] & _lib.ST_EDIT_OVERFLOW) != 0          # edit LIST truncated; counts complete
    assert 0 < overflow.sum() < 1000
    assert tot == int(aligned.sum())
    cover = sum(V["all_base_count_" + b] for b in "ACGTN-")
    assert (cover == tot).all()
    assert (V["all_deletion_count"] == V["all_base_count_-"]).all()
    assert blk.scalar("Reference", "MODIFIED") + blk.scalar("Reference", "UNMODIFIED") == tot
    assert blk.scalar("Reference", "MODIFIED") == int(a["modified"][aligned].sum())
    # size Counters and class_counts recomputed from the per-read records (every read has weight 1 here)
    H = blk.size_histograms("Reference")
    for key, col in (("inserted_n", a["insertion_n"]), ("deleted_n", a["deletion_n"]), ("substituted_n", a["substitution_n"]),
                     ("effective_len", 250 + a["insertion_n"].astype(np.int64) - a["deletion_n"].astype(np.int64))):
        want = np.bincount(np.asarray(col[aligned], dtype=np.int64))
        assert dict(H[key]) == {int(k): int(v) for k, v in enumerate(want) if v}, key
    assert blk.class_counts() == {"Reference_MODIFIED": int(a["modified"][aligned].sum()),
                                  "Reference_UNMODIFIED": int((a["modified"][aligned] == 0).sum())}
    # re-running the overflowed reads with a cap that cannot overflow (and zero weights) yields the complete lists
    idx = np.nonzero(overflow)[0]
    eng.set_edit_cap(512)
    zero = np.zeros(len(idx), dtype=np.int32)
    res3 = eng.align_packed(reads[idx].reshape(-1), np.arange(len(idx) + 1, dtype=np.int64) * 250, count=zero, qweight=zero)
    eng.set_edit_cap(8)
    assert (res3.recs["status"] == 0).all() and (res3.alns[:, 0]["n_edits"] > 8).all()
    assert eng.counts().scalar("Reference", "TOTAL") == int(aligned.sum())
    # (4b) the first 131 072 reads against the oracle, every field of every read + the count block of that sub-batch
    # (SURVEY.md 8d asks for >= 100k reads per config; oracle/batch_gate.py spreads them over the host cores)
    from oracle import batch_gate as BG
    G = 1 << 17
    eng.set_edit_cap(64)
    eng.counts_reset()
    resg = eng.align_packed(reads[:G].reshape(-1), off[:G + 1])
    summ, quant = BG.run(reads[:G].reshape(-1), off[:G + 1], {"Reference": ref}, ["Reference"], O.Params(), O.make_matrix(),
                         resg.recs, resg.alns, resg.strings, resg.edits, resg.W)
    assert summ["n"] == G and summ["n_bad"] == 0, summ
    assert BG.compare_block(eng.counts(), quant[None], ["Reference"]) == []
    for f in _lib.ALN_DTYPE.names:                                      # and the 1M-read launch gave the same records
        if f not in ("n_edits", "status"):
            assert (resg.alns[:, 0][f] == a[:G][f]).all(), f
    eng.set_edit_cap(8)
    # (5) batch-composition independence: a shuffled sub-batch reproduces its records bit for bit
    pick = rng.permutation(n)[:50000]
    eng.counts_reset()
    res2 = eng.align_packed(reads[pick].reshape(-1), np.arange(len(pick) + 1, dtype=np.int64) * 250)
    assert (res2.alns[:, 0] == res.alns[pick, 0]).all()
    valid = cols[pick][:, None, None, :]                               # bytes left of W - aln_len are undefined
    assert ((res2.strings == res.strings[pick]) | ~valid).all()


def test_empty_batch_and_bad_symbols(eng):
    rng = np.random.default_rng(1)
    amp = synth.random_amplicon(rng, 120)
    ref = synth.amplicon_setup(amp, guide_start=50)
    eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2)
    res = eng.align([])
    assert len(res.recs) == 0
    res = eng.align([amp, amp[:50] + "x" + amp[51:], amp.lower()[:60]])
    assert res.recs["status"][0] == 0 and res.recs["status"][1] == _lib.ST_BAD_CHAR and res.recs["status"][2] == _lib.ST_BAD_CHAR


def test_packed_pair_path_equals_32bit_path(eng):
    """The two-reads-per-warp 16-bit path and the one-read 32-bit path must agree bit for bit (records, strings,
    edit lists, count block) -- mixed lengths so that both pairs and singles occur in the default run."""
    rng = np.random.default_rng(77)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    base = synth.synth_reads(rng, amp, 40000, 250, sub_rate=0.02, rc_frac=0.05, n_rate=0.002, cut=ref["cut_point"])
    reads = [r.tobytes().decode() for r in base]
    for k in range(0, len(reads), 7):                      # every 7th read shortened: breaks some pairs
        reads[k] = reads[k][: 100 + (k % 140)]
    buf, off = pack_reads(reads)
    out = []
    for flags in (0, _lib.F_NO_PAIRING):
        eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, flags, "ACGTN", 24)
        eng.counts_reset()
        res = eng.align_packed(buf, off)
        out.append((res, eng.counts_raw(), eng.path_counts()))
    (a, ca, pa), (b, cb, pb) = out
    assert pa[0] > 15000 and pb[0] == 0 and pb[1] == 20000          # the pairing order puts equal lengths side by side
    assert (a.recs == b.recs).all() and (a.alns == b.alns).all() and (ca == cb).all()
    W = a.W
    cols = np.arange(W)[None, :] >= (W - a.alns[:, 0]["aln_len"].astype(np.int64))[:, None]
    assert ((a.strings[:, 0] == b.strings[:, 0]) | ~cols[:, None, :]).all()
    (ea, fa), (eb, fb) = PU.edits_canonical(a), PU.edits_canonical(b)
    assert (fa == fb).all() and (ea[fa] == eb[fb]).all()


@pytest.mark.parametrize("I", [250, 256, 131, 64])
def test_ring_banded_path_equals_full_matrix(eng, I):
    PU.check_ring_equals_full(eng, n=24000, I=I, seed=40 + I, oracle_subset=400)


@pytest.mark.parametrize("case", G.CASES)
def test_allele_tables_equal_the_reference(eng, case, tmp_path):
    """df_alleles, Alleles_frequency_table text and the tables around the cut (crispresso2_b200/alleles.py) against what the
    unmodified reference made of the same FASTQ (tests/golden)."""
    assert PU.check_alleles(eng, case, tmp_path) > 100


def test_leftover_list_with_single_reads_and_odd_tail(eng):
    PU.check_leftover_singles(eng, n=4001, seed=5)
    PU.check_leftover_singles(eng, n=43, seed=77)


def test_pairs_longer_than_512_columns(eng):
    PU.check_long_pairs(eng, n=400)


def test_legacy_insertion_quantification(eng):
    PU.check_legacy(eng, n=3000)


@pytest.mark.parametrize("I", [250, 256, 131])
def test_narrow_first_tier_equals_the_wide_ring(eng, I):
    PU.check_narrow_equals_wide(eng, n=32000, I=I, seed=50 + I, oracle_subset=480)


def test_pooled_ref_id(eng):
    """BASELINE configs[3] shape: many amplicons, each read aligned to its own one (ref_id), one launch."""
    PU.check_pooled(eng, n_amplicons=24, reads_per=120, amp_len=(180, 280))
    PU.check_pooled(eng, n_amplicons=96, reads_per=24, amp_len=(180, 280), seed=8)   # configs[3]: 96 amplicons in one configuration


def test_long_amplicon_three_row_blocks(eng):
    rng = np.random.default_rng(12)
    amp = synth.random_amplicon(rng, 610)
    ref = synth.amplicon_setup(amp, guide_start=300, window_size=5)
    reads = []
    for k in range(200):
        L = int(rng.integers(150, 401))
        s0 = int(rng.integers(0, 610 - L))
        s = synth.synth_reads(rng, amp, 1, 610, sub_rate=0.02, cut=ref["cut_point"])[0].tobytes().decode()
        reads.append(s[s0:s0 + L])
    reads.append(amp[100:500])
    PU.check_against_oracle(eng, {"Reference": ref}, ["Reference"], O.Params(), reads, O.make_matrix())


def test_banded_slab_falls_back_to_full_slab(eng):
    PU.check_band_fallback(eng, n=3000)


def test_coding_seq_frameshift_splicing_and_size_histograms(eng):
    PU.check_coding_seq(eng, n_reads=400)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mixed", [False, True])
def test_split_kernels_equal_general_kernel_and_chunking(eng, monkeypatch, mixed):
    """The two-kernel form (ALIGN + CLASSIFY + general kernel over the left-overs, chunks of 128 Ki reads) against the general
    kernel alone (C2B_NO_SPLIT=1) with a different chunking: same records, strings, edit lists and count block."""
    rng = np.random.default_rng(12)
    amp = synth.random_amplicon(rng, 250)
    ref = synth.amplicon_setup(amp)
    n = 150_001
    reads = synth.synth_reads_fast(rng, amp, n, 250, sub_rate=0.01, cut=ref["cut_point"])
    if mixed:
        lens = rng.integers(200, 251, size=n)
        lens[rng.random(n) < 0.7] = 250
    else:
        lens = np.full(n, 250)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    buf = reads[np.arange(250)[None, :] < lens[:, None]]
    cnt = rng.integers(1, 4, size=n).astype(np.int32)
    out = []
    for mode in ("split", "general"):
        if mode == "general":
            monkeypatch.setenv("C2B_NO_SPLIT", "1")
            monkeypatch.setenv("C2B_CHUNK", "40000")
        eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", 8)
        eng.counts_reset()
        res = eng.align_packed(buf, off, count=cnt, qweight=cnt)
        out.append((res, eng.counts_raw()))
    monkeypatch.delenv("C2B_NO_SPLIT", raising=False)
    monkeypatch.delenv("C2B_CHUNK", raising=False)
    (a, ca), (b, cb) = out
    assert (a.recs == b.recs).all() and (a.alns == b.alns).all() and (ca == cb).all()
    assert (a.recs["best_score_milli"] > 0).mean() > 0.9
    W = a.W
    valid = (np.arange(W)[None, :] >= (W - a.alns[:, 0]["aln_len"].astype(np.int64))[:, None])[:, None, None, :]
    assert ((a.strings == b.strings) | ~valid).all()
    (ea, fa), (eb, fb) = PU.edits_canonical(a), PU.edits_canonical(b)
    assert (fa == fb).all() and (ea[fa] == eb[fb]).all()


def test_random_configurations_against_oracle(eng):
    for seed in range(100, 160):
        PU.check_random_config(eng, seed)


def test_seed_tests_disagreeing_across_references(eng):
    assert PU.check_seed_disagreement(eng, n=8192) > 200
