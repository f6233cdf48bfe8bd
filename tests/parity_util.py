"""Parity checks shared by the CPU-emulator tests (-m "not gpu") and the GPU tests (-m gpu): the same
assertions run against either build of the engine."""
import types

import numpy as np

import golden_util as G
from crispresso2_b200 import core
from oracle import oracle as O


def args_from(params):
    a = types.SimpleNamespace(**params)
    a.use_legacy_insertion_quantification = False
    a.prime_editing_pegRNA_scaffold_seq = ""
    a.needleman_wunsch_aln_matrix_loc = "EDNAFULL"
    a.n_processes = "1"
    return a


def check_golden_case(engine, case, tmp_path, max_reads=None):
    """process_fastq on the fixture's reads must reproduce the reference's variantCache, aln_stats and count files."""
    rec = G.load(case)
    refs = G.refs_from(rec)
    reads = rec["reads"] if max_reads is None else rec["reads"][:max_reads]
    fq = tmp_path / (case + ".fastq")
    with open(fq, "w") as fh:
        for k, s in enumerate(reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    args = args_from(rec["params"])
    cache = {}
    stats, lost = core.process_fastq(str(fq), cache, rec["ref_names"], refs, args, [], str(tmp_path), engine=engine,
                                     aln_matrix=O.make_matrix())
    if max_reads is None:
        assert stats == rec["aln_stats"]
        assert list(cache.keys()) == list(rec["variants"].keys())
        assert set(lost) == set(rec["not_aligned"])
    for s, got in cache.items():
        want = rec["variants"][s]
        for k in ("aln_ref_names", "aln_scores", "best_match_score", "class_name", "best_match_name"):
            assert got[k] == want[k], (s, k, got[k], want[k])
        if max_reads is None:
            assert got["count"] == want["count"]
        assert [list(d) for d in got["ref_aln_details"]] == want["ref_aln_details"], s
        for r in want["aln_ref_names"]:
            bad = G.payload_equal(want["variant_" + r], got["variant_" + r])
            assert not bad, (s, r, bad, {b: (want["variant_" + r][b], got["variant_" + r][b]) for b in bad})
    for s, got in lost.items():
        assert got["aln_scores"] == rec["not_aligned"][s]["aln_scores"]
    if max_reads is not None:
        return
    block = core.quantify(cache)
    for r in rec["ref_names"]:
        seq = refs[r]["sequence"]
        V = block.vectors(r)
        tot = block.scalar(r, "TOTAL")
        assert G.mod_count_text(seq, V, tot) == G.file_for(rec, r, "Modification_count_vectors.txt")
        assert G.qw_count_text(seq, V, tot) == G.file_for(rec, r, "Quantification_window_modification_count_vectors.txt")
        nf = G.nuc_freq_rows(G.file_for(rec, r, "Nucleotide_frequency_table.txt"))
        for b in "ACGTN-":
            assert (nf[b] == V["all_base_count_" + b]).all(), (r, b)


def check_against_oracle(engine, refs, ref_names, params, reads, matrix):
    """Engine vs oracle on arbitrary reads: per-read variants, aln_stats, every count vector and counter."""
    import tempfile, os
    cache_o, stats_o, lost_o = O.process_reads(reads, refs, ref_names, params, matrix)
    d = tempfile.mkdtemp()
    fq = os.path.join(d, "r.fastq")
    with open(fq, "w") as fh:
        for k, s in enumerate(reads):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    args = args_from({k: getattr(params, k) for k in vars(params)})
    cache = {}
    stats, lost = core.process_fastq(fq, cache, ref_names, refs, args, [], d, engine=engine, aln_matrix=matrix)
    assert stats == stats_o
    assert list(cache.keys()) == list(cache_o.keys())
    assert set(lost) == set(lost_o)
    for s, want in cache_o.items():
        got = cache[s]
        for k in ("count", "aln_ref_names", "aln_scores", "best_match_score", "class_name", "best_match_name"):
            assert got[k] == want[k], (s, k, got[k], want[k])
        assert [tuple(x) for x in got["ref_aln_details"]] == [tuple(x) for x in want["ref_aln_details"]], s
        for r in want["aln_ref_names"]:
            bad = G.payload_equal(want["variant_" + r], got["variant_" + r])
            assert not bad, (s, r, bad)
    vec, sca, classes, total = O.count_vectors(cache_o, refs, ref_names, params)
    block = core.quantify(cache)
    for r in ref_names:
        V = block.vectors(r)
        for name in O.VECTOR_NAMES:
            assert (V[name] == vec[r][name]).all(), (r, name, np.nonzero(V[name] != vec[r][name]))
        S = block.scalars(r)
        for name in O.SCALAR_NAMES:
            assert S[name] == sca[r][name], (r, name, S[name], sca[r][name])
