#!/usr/bin/env python
"""Generates tests/golden/*.json.gz + *.txt by running the UNMODIFIED reference in this container.

Run from the repo root:  python tests/golden/gen_golden.py
Needs /root/reference (pure-python files, imported in place) and oracle/_ref (the reference's two Cython
modules compiled by oracle/Makefile).  The GPU box has neither; it only reads the committed fixtures.

How: import the reference's CRISPRessoCORE with three stubs (seaborn, the matplotlib plotting module, the
package version lookup -- recipe in SURVEY.md Appendix C), wrap its module-global `process_fastq`
(CRISPRessoCORE.py:1735, looked up at :3750) to record what went in (reads, refs, params) and what came
out (variantCache, aln_stats), let main() finish, and keep the count-vector tables it writes.
"""
import gzip
import importlib.metadata as _md
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))

_sb = types.ModuleType("seaborn")
_sb.set_context = _sb.set = _sb.set_style = lambda *a, **k: None
sys.modules["seaborn"] = _sb
_orig_version = _md.version
_md.version = lambda n: "2.3.4" if n == "CRISPResso2" else _orig_version(n)
import CRISPResso2  # noqa: E402  (oracle/_ref/CRISPResso2: empty __init__ + compiled modules)

CRISPResso2.__path__.append("/root/reference/CRISPResso2")
_fake = types.ModuleType("CRISPResso2.plots.CRISPRessoPlot")
_fake.setMatplotlibDefaults = lambda *a, **k: None
sys.modules["CRISPResso2.plots.CRISPRessoPlot"] = _fake
sys.modules["CRISPResso2.plots.upsetplot"] = types.ModuleType("CRISPResso2.plots.upsetplot")
from CRISPResso2 import CRISPRessoCORE  # noqa: E402

from crispresso2_b200 import synth  # noqa: E402

FANC = ("CGGATGTTCCAATCAGTACGCAGAGAGTCGCCGTCTCCAAGGTGAAAGCGGAAGTAGGGCCTTCGCGCACCTCATGGAATCCCTTCTGCAGCACCTGGATCGCT"
        "TTTCCGAGCTTCTGGCGGTCTCAAGCACTACCTACGTCAGCACCTGGGACCCCGCCACCGTGCGCCGGGCCTTGCAGTGGGCGCGCTACCTGCGCCACATCCAT"
        "CGGCGCTTTGGTCGG")
FANC_HDR = ("CGGCCGGATGTTCCAATCAGTACGCAGAGAGTCGCCGTCTCCAAGGTGAAAGCTGAAGTAGGGCCTTCGCGCACCTCATGGAATCCCTTCTGCAGCTTTT"
            "CCGAGCTTCTGGCGGTCTCAAGCACTACCTACGTCAGCACCTGGGACCCCGCCACCGTGCGCCGGGCCTTGCAGTGGGCGCGCTACCTGCGCCACATCCA"
            "TCGGCGCTTTGGTCGG")

PARAM_KEYS = ["aln_seed_count", "aln_seed_min", "needleman_wunsch_gap_open", "needleman_wunsch_gap_extend",
              "ignore_substitutions", "ignore_insertions", "ignore_deletions",
              "assign_ambiguous_alignments_to_first_reference", "expand_ambiguous_alignments", "discard_indel_reads",
              "prime_editing_pegRNA_extension_seq", "prime_editing_pegRNA_scaffold_seq", "prime_editing_pegRNA_scaffold_min_match_length"]


def _plain(o):
    if isinstance(o, dict):
        return {str(k): _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, np.ndarray):
        return [_plain(v) for v in o.tolist()]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if hasattr(o, "__slots__") and hasattr(o, "__dict__"):
        return _plain(o.__dict__)
    return o


def run_case(name, argv, keep_files):
    rec = {}
    orig = CRISPRessoCORE.process_fastq

    def spy(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory):
        opener = gzip.open if fastq_filename.endswith(".gz") else open
        with opener(fastq_filename, "rt") as fh:
            rec["reads"] = [ln.strip() for k, ln in enumerate(fh) if k % 4 == 1]
        rec["ref_names"] = list(ref_names)
        rec["refs"] = {r: {"sequence": refs[r]["sequence"], "gap_incentive": _plain(refs[r]["gap_incentive"]),
                           "include_idxs": _plain(refs[r]["include_idxs"]), "fw_seeds": list(refs[r]["fw_seeds"]),
                           "rc_seeds": list(refs[r]["rc_seeds"]), "min_aln_score": refs[r]["min_aln_score"]}
                       for r in ref_names}
        rec["params"] = {k: getattr(args, k) for k in PARAM_KEYS}
        out = orig(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory)
        rec["aln_stats"] = _plain(out[0])
        rec["not_aligned"] = {s: {"count": v["count"], "aln_scores": _plain(v["aln_scores"]),
                                  "best_match_score": v["best_match_score"]} for s, v in out[1].items()}
        rec["variants"] = _plain({s: v for s, v in variantCache.items()})
        return out

    CRISPRessoCORE.process_fastq = spy
    orig_ctx = CRISPRessoCORE.CorePlotContext

    def ctx_spy(*a, **kw):
        # the HDR "ref1" re-projection vectors (CRISPRessoCORE.py:4195-4272) only leave main() through this object
        keys = ["ref1_all_insertion_count_vectors", "ref1_all_insertion_left_count_vectors", "ref1_all_deletion_count_vectors",
                "ref1_all_substitution_count_vectors", "ref1_all_indelsub_count_vectors", "ref1_all_base_count_vectors"]
        rec["ref1"] = {k: _plain(kw.get(k, {})) for k in keys}
        # allele-level consumers: df_alleles as main() built it (CRISPRessoCORE.py:4298-4303) and what the
        # reference's own get_dataframe_around_cut_asymmetrical (CRISPRessoShared.py:1518-1531) makes of it
        from CRISPResso2 import CRISPRessoShared
        df = kw["df_alleles"]
        cols = ["#Reads", "Aligned_Sequence", "Reference_Sequence", "n_inserted", "n_deleted", "n_mutated", "Reference_Name",
                "Read_Status", "Aligned_Reference_Names", "Aligned_Reference_Scores", "ref_positions", "%Reads"]
        rec["alleles"] = {"tsv": df.loc[:, cols].to_csv(sep="\t", header=True, index=None), "index": [int(x) for x in df.index],
                          "dtypes": {c: str(df[c].dtype) for c in cols}, "n_total": int(kw["N_TOTAL"]), "around_cut": {}}
        for rn in kw["ref_names"]:
            cuts = [int(c) for c in kw["refs"][rn]["sgRNA_cut_points"]] or [len(kw["refs"][rn]["sequence"]) // 2]
            for cut in cuts:
                for (pl, pr) in ((20, 20), (7, 31)):
                    sub = df.loc[df["Reference_Name"] == rn]
                    if sub.shape[0] == 0:
                        continue
                    out = CRISPRessoShared.get_dataframe_around_cut_asymmetrical(sub, cut, pl, pr)
                    rec["alleles"]["around_cut"]["%s|%d|%d|%d" % (rn, cut, pl, pr)] = out.to_csv(sep="\t", header=True)
        return orig_ctx(*a, **kw)

    CRISPRessoCORE.CorePlotContext = ctx_spy
    work = tempfile.mkdtemp(prefix="c2gold_")
    old_argv, old_cwd = sys.argv, os.getcwd()
    try:
        os.chdir(work)
        sys.argv = ["CRISPResso"] + argv + ["--suppress_plots", "--suppress_report", "-o", work]
        try:
            CRISPRessoCORE.main()
        except SystemExit as e:
            if e.code not in (0, None):
                raise
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        CRISPRessoCORE.process_fastq = orig
        CRISPRessoCORE.CorePlotContext = orig_ctx
    run_dirs = [d for d in os.listdir(work) if d.startswith("CRISPResso_on_")]
    assert len(run_dirs) == 1, run_dirs
    rd = os.path.join(work, run_dirs[0])
    rec["files"] = {}
    for f in sorted(os.listdir(rd)):
        if any(f.endswith(k) for k in keep_files):
            with open(os.path.join(rd, f)) as fh:
                rec["files"][f] = fh.read()
    with gzip.open(os.path.join(OUT, name + ".json.gz"), "wt") as fh:
        json.dump(rec, fh, separators=(",", ":"))
    print(name, "reads", len(rec["reads"]), "uniques", len(rec["variants"]) + len(rec["not_aligned"]),
          "files", list(rec["files"]))
    return rec, rd, work


KEEP = ["Modification_count_vectors.txt", "Quantification_window_modification_count_vectors.txt",
        "Nucleotide_frequency_table.txt", "CRISPResso_quantification_of_editing_frequency.txt",
        "CRISPResso_mapping_statistics.txt"]


def check_vendored(rec, exp_dir, names):
    """The reference run made here must reproduce the golden files the reference repo vendors."""
    for ours, theirs in names:
        with open(os.path.join(exp_dir, theirs)) as fh:
            want = fh.read()
        assert rec["files"][ours] == want, "vendored golden %s not reproduced" % theirs
        print("  vendored golden reproduced:", theirs)


def alignment_vectors():
    """Known-answer + random vectors for global_align / find_indels_substitutions from the compiled reference."""
    from CRISPResso2 import CRISPResso2Align as A, CRISPRessoCOREResources as R
    import random
    m = A.read_matrix("/root/reference/CRISPResso2/EDNAFULL")
    rng = random.Random(20260923)
    cases = []

    def add(read, ref, gi, go, ge, inc):
        s1, s2, sc = A.global_align(read, ref, matrix=m, gap_incentive=np.array(gi, dtype=np.int64), gap_open=go,
                                    gap_extend=ge)
        p = _plain(R.find_indels_substitutions(s1, s2, inc).__dict__)
        cases.append({"read": read, "ref": ref, "gi": gi, "go": go, "ge": ge, "inc": inc, "s1": s1, "s2": s2,
                      "score": sc, "payload": p})

    for I in [12, 25, 60, 100, 180, 250]:
        for _ in range(40):
            ref = "".join(rng.choice("ACGT") for _ in range(I))
            if rng.random() < 0.1:
                k = rng.randrange(I)
                ref = ref[:k] + "N" + ref[k + 1:]
            read = []
            for c in ref:
                u = rng.random()
                if u < 0.03:
                    read.append(rng.choice("ACGTN"))
                elif u < 0.04:
                    continue
                elif u < 0.05:
                    read.append(c + rng.choice("ACGT"))
                else:
                    read.append(c)
            read = "".join(read)
            u = rng.random()
            if u < 0.3:
                k, p = rng.randrange(1, max(2, I // 4)), rng.randrange(I)
                read = read[:p] + read[p + k:]
            elif u < 0.45:
                k, p = rng.randrange(1, 12), rng.randrange(I)
                read = read[:p] + "".join(rng.choice("ACGT") for _ in range(k)) + read[p:]
            elif u < 0.5:
                read = "".join(rng.choice("ACGT") for _ in range(rng.randrange(max(4, I // 2), I + 20)))
            if len(read) < 3:
                continue
            gi = [0] * (I + 1)
            for _ in range(rng.randrange(0, 3)):
                gi[rng.randrange(I + 1)] = rng.choice([1, 1, 2, 5])
            go, ge = rng.choice([(-20, -2), (-20, -2), (-1, -1), (-5, -3), (-10, 0)])
            inc = sorted(rng.sample(range(I), rng.randrange(0, min(I, 8))))
            add(read, ref, gi, go, ge, inc)
    with gzip.open(os.path.join(OUT, "align_vectors.json.gz"), "wt") as fh:
        json.dump(cases, fh, separators=(",", ":"))
    print("align_vectors", len(cases))


def pe_scaffold_case():
    """prime editing with a scaffold sequence: 'Scaffold-incorporated' re-labelling (CRISPRessoCORE.py:789-796, :3759-3764)"""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import pe_case
    fq = os.path.join(tempfile.gettempdir(), "c2gold_pe.fastq")
    ext, scaffold = pe_case.write_fastq(fq, FANC)
    rec, rd, work = run_case("fanc_pe_scaffold", ["-r1", fq, "-a", FANC, "--prime_editing_pegRNA_spacer_seq", pe_case.SPACER,
                                                  "--prime_editing_pegRNA_extension_seq", ext, "--prime_editing_pegRNA_scaffold_seq", scaffold],
                             KEEP)
    assert any(v.get("class_name") == "Scaffold-incorporated" for v in rec["variants"].values())
    shutil.rmtree(work)


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["fanc_pe_scaffold"]:
        pe_scaffold_case()
        return
    alignment_vectors()

    fanc_fq = "/root/reference/tests/FANC.Cas9.fastq"
    rec, rd, work = run_case("fanc_cas9", ["-r1", fanc_fq, "-a", FANC, "-g", "GGAATCCCTTCTGCAGCACC"], KEEP)
    check_vendored(rec, "/root/reference/tests/expectedResults/CRISPResso_on_FANC.Cas9",
                   [("CRISPResso_quantification_of_editing_frequency.txt",
                     "CRISPResso_quantification_of_editing_frequency.txt"),
                    ("Nucleotide_frequency_table.txt", "Nucleotide_frequency_table.txt")])
    shutil.rmtree(work)

    rec, rd, work = run_case("fanc_params", [
        "-r1", fanc_fq, "-a", FANC, "-g", "GGAATCCCTTCTGCAGCACC", "-e", FANC_HDR,
        "-c", "GGGCCTTCGCGCACCTCATGGAATCCCTTCTGCAGCACCTGGATCGCTTTT", "--dump", "-qwc", "20-30_45-50", "-q", "30",
        "--default_min_aln_score", "80", "-an", "FANC", "-n", "params", "--base_editor_output",
        "-fg", "AGCCTTGCAGTGGGCGCGCTA,CCCACTGAAGGCCC", "--dsODN", "GCTAGATTTCCCAAGAAGA", "-gn", "hi", "-fgn", "dear"],
        KEEP)
    check_vendored(rec, "/root/reference/tests/expectedResults/CRISPResso_on_params",
                   [("CRISPResso_quantification_of_editing_frequency.txt",
                     "CRISPResso_quantification_of_editing_frequency.txt"),
                    ("FANC.Nucleotide_frequency_table.txt", "FANC.Nucleotide_frequency_table.txt")])
    shutil.rmtree(work)

    # synthetic single amplicon (generator spec of SURVEY 8d), incl. reverse-complemented reads and N's
    rng = np.random.default_rng(42)
    amp = synth.random_amplicon(rng, 250)
    reads = synth.synth_reads(rng, amp, 1500, 250, sub_rate=0.005, rc_frac=0.05, n_rate=0.001, cut=126)
    fq = os.path.join(tempfile.gettempdir(), "c2gold_synth1.fastq")
    synth.write_fastq(fq, reads)
    rec, rd, work = run_case("synth_single", ["-r1", fq, "-a", amp, "-g", amp[110:130]], KEEP)
    shutil.rmtree(work)

    # synthetic HDR: WT + HDR allele (3-bp substitution + 6-bp insertion near the cut), mixed reads
    hdr = amp[:120] + "TGA" + amp[123:127] + "ACGTAC" + amp[127:]
    r_wt = synth.synth_reads(rng, amp, 700, 250, cut=126)
    r_hdr = synth.synth_reads(rng, hdr, 400, 250, del_frac=0.05, ins_frac=0.02, cut=130)
    allr = np.concatenate([r_wt, r_hdr])[rng.permutation(1100)]
    fq2 = os.path.join(tempfile.gettempdir(), "c2gold_synth2.fastq")
    synth.write_fastq(fq2, allr)
    rec, rd, work = run_case("synth_hdr", ["-r1", fq2, "-a", amp, "-g", amp[110:130], "-e", hdr], KEEP)
    shutil.rmtree(work)
    pe_scaffold_case()


if __name__ == "__main__":
    main()
