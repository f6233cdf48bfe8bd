"""Host side of the drop-in (CPU, warp-emulator engine): lazy variantCache entries, the compact op-stream outputs and
their host-side expansion, the native reverse-complement merge, screening of reads outside the engine's contract."""
import copy
import json
import os
import pickle
import sys

import numpy as np
import pytest

import parity_util as PU
from crispresso2_b200 import _lib, core, lazy, synth
from crispresso2_b200.engine import Engine
from oracle import oracle as O


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    return Engine(lib_path=build_emu.build())


def _fastq(tmp_path, reads, name="r.fastq"):
    fq = tmp_path / name
    with open(fq, "w") as fh:
        for k, s in enumerate(reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    return str(fq)


def _setup(seed=5, n=60, L=200):
    rng = np.random.default_rng(seed)
    amp = synth.random_amplicon(rng, L)
    ref = synth.amplicon_setup(amp, guide_start=L // 2 - 10)
    reads = synth.synth_reads(rng, amp, n, L, sub_rate=0.02, rc_frac=0.2, n_rate=0.003, cut=ref["cut_point"])
    reads = [r.tobytes().decode() for r in reads]
    return amp, ref, reads + reads[:15]                       # duplicates: counts > 1


def test_lazy_variants_are_the_reference_dicts(emu, tmp_path):
    amp, ref, reads = _setup()
    refs, names = {"Reference": ref}, ["Reference"]
    cache = {}
    st, lost = core.process_fastq(_fastq(tmp_path, reads), cache, names, refs, PU.args_from(vars(O.Params())), [], str(tmp_path),
                                  engine=emu, aln_matrix=O.make_matrix())
    cache_o, st_o, lost_o = O.process_reads(reads, refs, names, O.Params(), O.make_matrix())
    assert st == st_o and list(cache) == list(cache_o) and set(lost) == set(lost_o)
    vals = list(cache.values())
    assert all(isinstance(v, lazy.LazyVariant) and isinstance(v, dict) for v in vals)
    assert all(v._k >= 0 for v in vals)                      # nothing was materialised by process_fastq itself
    v0 = vals[0]
    assert v0["count"] == cache_o[list(cache_o)[0]]["count"] and v0._k >= 0      # 'count' alone does not materialise
    v0["count"] = 0                                          # the reference's rc-merge writes counts before reading anything else
    assert v0._k >= 0
    assert v0["class_name"] in ("Reference_MODIFIED", "Reference_UNMODIFIED") and v0._k < 0
    assert v0["count"] == 0                                  # ... and the written count survives materialisation
    v0["count"] = cache_o[list(cache_o)[0]]["count"]
    for s, want in cache_o.items():
        got = cache[s]
        for k in ("count", "aln_ref_names", "aln_scores", "best_match_score", "class_name", "best_match_name"):
            assert got[k] == want[k], (s, k)
        assert [tuple(x) for x in got["ref_aln_details"]] == [tuple(x) for x in want["ref_aln_details"]]
        p, q = got["variant_Reference"], want["variant_Reference"]
        assert isinstance(p["ref_positions"], list) and isinstance(p["all_insertion_positions"], list)
        assert p["ref_positions"] == q["ref_positions"] and str(p["all_deletion_positions"]) == str(q["all_deletion_positions"])
    # mapping protocol on a fresh (unmaterialised) entry
    cache2 = {}
    core.process_fastq(_fastq(tmp_path, reads), cache2, names, refs, PU.args_from(vars(O.Params())), [], str(tmp_path),
                       engine=emu, aln_matrix=O.make_matrix())
    it = iter(cache2.values())
    a, b, c, d, e = (next(it) for _ in range(5))
    assert "class_name" in a and a._k < 0
    assert set(b.keys()) >= {"count", "aln_scores", "ref_aln_details", "best_match_score", "aln_ref_names", "class_name"}
    assert c.get("best_match_name") == "Reference" and c.get("nope", 7) == 7
    assert len(d) == len(b) and dict(d) == d
    blob = pickle.loads(pickle.dumps(e))
    assert type(blob) is dict and blob["class_name"] == e["class_name"]
    assert copy.deepcopy(e)["aln_scores"] == e["aln_scores"]
    with pytest.raises(KeyError):
        e["no_such_key"]
    # unaligned reads come back lazily too, with the reference's four keys
    junk = "".join(np.random.default_rng(1).choice(list("ACGT"), 200))
    cache3 = {}
    st3, lost3 = core.process_fastq(_fastq(tmp_path, reads + [junk]), cache3, names, refs, PU.args_from(vars(O.Params())), [],
                                    str(tmp_path), engine=emu, aln_matrix=O.make_matrix())
    assert junk in lost3 and junk not in cache3 and st3["N_COMPUTED_NOTALN"] == 1
    assert set(lost3[junk].keys()) == {"count", "aln_scores", "ref_aln_details", "best_match_score"}
    assert lost3[junk]["best_match_score"] == -1


def test_preseeded_cache_keeps_order_and_adds_counts(emu, tmp_path):
    amp, ref, reads = _setup(seed=9, n=20)
    refs, names = {"Reference": ref}, ["Reference"]
    cache = {reads[3]: 5, reads[0]: 2}                        # caller-seeded counts (the += semantics of :1839-1844)
    st, _ = core.process_fastq(_fastq(tmp_path, reads), cache, names, refs, PU.args_from(vars(O.Params())), [], str(tmp_path),
                               engine=emu, aln_matrix=O.make_matrix())
    assert list(cache)[:2] == [reads[3], reads[0]]
    assert cache[reads[3]]["count"] == 5 + reads.count(reads[3]) and cache[reads[0]]["count"] == 2 + reads.count(reads[0])
    assert st["N_TOT_READS"] == len(reads) + 7


def test_compact_outputs_expand_to_the_same_strings(emu):
    rng = np.random.default_rng(4)
    amp = synth.random_amplicon(rng, 250)
    hdr = amp[:120] + "TGA" + amp[123:127] + "ACGTAC" + amp[127:]
    refs = {"WT": synth.amplicon_setup(amp), "HDR": synth.amplicon_setup(hdr)}
    names = ["WT", "HDR"]
    reads = np.concatenate([synth.synth_reads(rng, a, 40, 250, sub_rate=0.02, rc_frac=0.2, cut=126) for a in (amp, hdr)])
    lens = rng.integers(120, 251, size=len(reads))
    lens[::3] = 250
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    buf = reads[np.arange(250)[None, :] < lens[:, None]]
    emu.configure(refs, names, O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", 32)
    emu.counts_reset()
    a = emu.align_packed(buf, off)
    ca = emu.counts_raw()
    emu.counts_reset()
    b = emu.align_packed(buf, off, compact=True)
    assert (a.recs == b.recs).all() and (a.alns == b.alns).all() and (ca == emu.counts_raw()).all()
    assert b.strings is None and b.ops.shape == (len(reads), 2, a.W // 32)
    assert ((b.meta & 0xffff) == a.alns["aln_len"]).all() and (((b.meta >> 16) & 1) == a.alns["strand"]).all()
    exp = b.strings_block(0, len(reads))
    cols = np.arange(a.W)[None, None, None, :] >= (a.W - a.alns["aln_len"].astype(np.int64))[:, :, None, None]
    assert ((exp == a.strings) | ~cols).all()
    for i in (0, 7, 41, 79):
        for r in (0, 1):
            assert b.pair(i, r) == a.pair(i, r)
    # single-slot entry
    import ctypes as C
    i, r = 5, 1
    n = int(a.alns[i, r]["aln_len"])
    o1, o2 = C.create_string_buffer(n), C.create_string_buffer(n)
    read = buf[off[i]:off[i + 1]].tobytes()
    rc = emu.L.c2b_expand_alignment(emu.h, b.ops[i, r].ctypes.data, int(b.meta[i, r]), read, len(read), refs[names[r]]["sequence"].encode(),
                                    len(refs[names[r]]["sequence"]), o1, o2)
    assert rc == 0 and (o1.raw.decode(), o2.raw.decode()) == a.pair(i, r)


def test_native_rc_merge_equals_the_python_rule():
    rng = np.random.default_rng(8)
    base = ["".join(rng.choice(list("ACGT"), int(rng.integers(5, 40)))) for _ in range(300)]
    uniq = []
    for s in base:
        uniq.append(s)
        if rng.random() < 0.4:
            uniq.append(core.reverse_complement(s))
    uniq += ["ACGT", "AATT", "ACGN", "NNNN", "acgt", "ACGU", "AC-GT", "TTAA"]     # palindromes, N, lower case, a symbol the reference rejects
    uniq = list(dict.fromkeys(uniq))
    rng.shuffle(uniq)
    counts = rng.integers(1, 9, size=len(uniq)).astype(np.int32)
    from crispresso2_b200.engine import pack_reads
    buf, off = pack_reads(uniq)
    got = core.merge_weights_packed(buf, off, counts)
    want = core.merge_weights(uniq, counts.tolist())
    assert got.tolist() == want
    member = (rng.random(len(uniq)) < 0.7).astype(np.uint8)
    got2 = core.merge_weights_packed(buf, off, counts, member=member)
    keep = [k for k in range(len(uniq)) if member[k]]
    want2 = core.merge_weights([uniq[k] for k in keep], [int(counts[k]) for k in keep])
    assert [int(got2[k]) for k in keep] == want2


def test_out_of_contract_reads_do_not_abort_the_run(emu, tmp_path):
    amp, ref, reads = _setup(seed=3, n=30)
    refs, names = {"Reference": ref}, ["Reference"]
    odd = [reads[0].lower(), reads[1][:50] + "R" + reads[1][51:], "A" * 600, reads[2][:80] + "Y" + reads[2][81:]]
    mixed = reads[:10] + [odd[0]] + reads[10:20] + [odd[1], odd[2]] + reads[20:] + [odd[3], odd[2]]
    cache = {}
    st, lost = core.process_fastq(_fastq(tmp_path, mixed), cache, names, refs, PU.args_from(vars(O.Params())), [], str(tmp_path),
                                  engine=emu, aln_matrix=O.make_matrix())
    cache_o, st_o, lost_o = O.process_reads(reads, refs, names, O.Params(), O.make_matrix())
    assert list(cache) == list(cache_o)
    for s in odd:
        assert s in lost and lost[s]["best_match_score"] == -1 and s not in cache
    assert lost[odd[2]]["count"] == 2
    assert st["N_TOT_READS"] == len(mixed) and st["N_COMPUTED_NOTALN"] == st_o["N_COMPUTED_NOTALN"] + 4
    assert st["N_CACHED_NOTALN"] == st_o["N_CACHED_NOTALN"] + 1
    for k in ("N_COMPUTED_ALN", "N_CACHED_ALN", "N_GLOBAL_SUBS", "N_MODS_IN_WINDOW", "READ_LENGTH"):
        assert st[k] == st_o[k], k
    blk = core.quantify(cache)
    vec, sca, classes, total = O.count_vectors(cache_o, refs, names, O.Params(), {})
    assert blk.class_counts() == classes
    with pytest.raises(core.EngineError):
        core.process_fastq(_fastq(tmp_path, mixed), {}, names, refs, PU.args_from(vars(O.Params())), [], str(tmp_path),
                           engine=emu, aln_matrix=O.make_matrix(), on_out_of_contract="error")


def test_assign_first_wins_over_expand_when_both_flags_are_set(emu, tmp_path):
    """CRISPRessoCORE.py:780-785 tests assign_ambiguous_alignments_to_first_reference first; the CLI rejects the combination but
    library callers can pass it."""
    rng = np.random.default_rng(2)
    amp = synth.random_amplicon(rng, 150)
    refs = {"A": synth.amplicon_setup(amp, guide_start=60), "B": synth.amplicon_setup(amp, guide_start=60)}     # identical: every read is a tie
    reads = [r.tobytes().decode() for r in synth.synth_reads(rng, amp, 12, 150, sub_rate=0.02, cut=refs["A"]["cut_point"])]
    p = O.Params(assign_ambiguous_alignments_to_first_reference=True, expand_ambiguous_alignments=True)
    cache = {}
    core.process_fastq(_fastq(tmp_path, reads), cache, ["A", "B"], refs, PU.args_from(vars(p)), [], str(tmp_path), engine=emu,
                       aln_matrix=O.make_matrix())
    cache_o, _, _ = O.process_reads(reads, refs, ["A", "B"], p, O.make_matrix())
    for s, want in cache_o.items():
        assert cache[s]["class_name"] == want["class_name"] and cache[s]["aln_ref_names"] == want["aln_ref_names"] == ["A"]
    vec, sca, classes, total = O.count_vectors(cache_o, refs, ["A", "B"], p, {})
    assert core.quantify(cache).class_counts() == classes


def test_bulk_keys_carry_the_hashes_python_would_compute():
    """lazy.make_keys pre-computes the str hashes of the variantCache keys on plain threads (csrc/c2b_pyext.c: hash_ahead) and
    fill_cache stages an empty dict through a right-sized temporary: equal strings built any other way must hash alike, find the
    entries, and the cache must keep first-seen order; a pre-seeded cache takes the plain path."""
    from crispresso2_b200 import lazy
    rng = np.random.default_rng(3)
    n = 20000                                                # above hash_ahead's and the staging dict's thresholds
    lens = rng.integers(0, 60, size=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    buf = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(off[-1])).astype(np.uint8)
    keys = lazy.make_keys(buf, off)
    raw = buf.tobytes()
    for k in range(0, n, 37):
        fresh = raw[off[k]:off[k + 1]].decode("ascii")
        assert keys[k] == fresh and hash(keys[k]) == hash(fresh)

    class LV(dict):
        __slots__ = ("_k",)

    uniq = {}
    for k, s in enumerate(keys):
        uniq.setdefault(s, k)
    sel = np.zeros(n, dtype=np.uint8)
    sel[list(uniq.values())] = 1
    sel[::11] = 0
    counts = np.ones(n, dtype=np.int32)
    cache = {}
    made = lazy.fill_cache(cache, keys, sel, counts, LV, 1)
    want = [k for k in range(n) if sel[k]]
    assert made == len(want) == len(cache) and [v._k for v in cache.values()] == want and list(cache) == [keys[k] for k in want]
    assert all(raw[off[k]:off[k + 1]].decode("ascii") in cache for k in want[::53])
    seeded = {"seed": 1}
    assert lazy.fill_cache(seeded, keys, sel, counts, LV, 1) == len(want) and list(seeded)[0] == "seed" and len(seeded) == len(want) + 1
