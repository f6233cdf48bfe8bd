#!/usr/bin/env python
"""Runs the UNMODIFIED reference `CRISPResso` main() (imported from /root/reference through the stubs of
tests/golden/gen_golden.py) either as-is or with its module-global `process_fastq` (CRISPRessoCORE.py:1735,
resolved at :3750) re-bound to crispresso2_b200.core.process_fastq -- the one-line integration of
INTEGRATION.md section 2.  Used by tests/test_cli_dropin.py, which diffs the two output folders byte for byte.

usage: cli_dropin_runner.py <reference|b200> <engine-lib-or-'default'> <outdir> <vectors.json> -- <CRISPResso argv>

In b200 mode the count vectors / counters the reference's own quantification loop built (CRISPRessoCORE.py:3964-4303,
captured from the CorePlotContext it constructs at :4836) are also compared here with the engine's count block and
the outcome is written to <vectors.json>.

TEST INFRASTRUCTURE: needs /root/reference and oracle/_ref; never runs on the GPU box.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402


def main():
    mode, lib, outdir, vec_json = sys.argv[1:5]
    argv = sys.argv[sys.argv.index("--") + 1:]
    import gen_golden as GG                                  # installs the stubs, imports the reference CORE
    CORE = GG.CRISPRessoCORE
    report = {"mode": mode, "mismatch": [], "checked": 0}
    if "--crispresso_merge" in argv:                          # fastp is a third-party binary this image does not have: tests/fake_fastp.py
        CORE.check_fastp = lambda: None
    if mode == "b200":
        from crispresso2_b200 import core
        from crispresso2_b200.engine import Engine
        eng = Engine(lib_path=None if lib == "default" else lib)
        state = {}

        # the shipped launcher's re-binding (crispresso2_b200/launcher.py: process_fastq, filterFastqs, the table around the cut),
        # with the count block of the call kept for the comparison below
        from crispresso2_b200 import launcher
        launcher.bind(CORE, engine=eng, lib_path=None if lib == "default" else lib)
        bound = CORE.process_fastq

        def process_fastq(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory):
            out = bound(fastq_filename, variantCache, ref_names, refs, args, files_to_remove, output_directory)
            state["block"] = core.quantify(variantCache)
            state["ref_names"] = list(state["block"].ref_names)      # prime editing with a scaffold: + 'Scaffold-incorporated'
            return out

        CORE.process_fastq = process_fastq
        orig_ctx = CORE.CorePlotContext

        def ctx_spy(*a, **kw):
            if "block" not in state:                          # paired-end merge mode: the reference's loop built the counts itself
                from crispresso2_b200 import paired
                memo = paired.process_paired_fastq.last_memo
                report["memo"] = {"hits": memo.hits, "misses": memo.misses, "sequences": len(memo.index)}
                return orig_ctx(*a, **kw)
            blk = state["block"]
            bad = report["mismatch"]

            def cmp(name, got, want):
                report["checked"] += 1
                if not np.array_equal(np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)):
                    bad.append(name)

            for r in state["ref_names"]:
                V = blk.vectors(r)
                S = blk.scalars(r)
                for ours, theirs in (("all_insertion_count", "all_insertion_count_vectors"),
                                     ("all_insertion_left_count", "all_insertion_left_count_vectors"),
                                     ("all_deletion_count", "all_deletion_count_vectors"),
                                     ("all_substitution_count", "all_substitution_count_vectors"),
                                     ("insertion_count", "insertion_count_vectors"),
                                     ("deletion_count", "deletion_count_vectors"),
                                     ("substitution_count", "substitution_count_vectors")):
                    cmp(r + ":" + theirs, V[ours], kw[theirs][r])
                for ch in "ACGTN":
                    cmp(r + ":all_substitution_base_" + ch, V["all_substitution_base_" + ch],
                        kw["all_substitution_base_vectors"][r + "_" + ch])
                for ch in "ACGTN-":
                    cmp(r + ":all_base_count_" + ch, V["all_base_count_" + ch], kw["all_base_count_vectors"][r + "_" + ch])
                for ours, theirs in (("counts_total", "counts_total"), ("counts_modified", "counts_modified"),
                                     ("counts_unmodified", "counts_unmodified"), ("counts_discarded", "counts_discarded"),
                                     ("counts_insertion", "counts_insertion"), ("counts_deletion", "counts_deletion"),
                                     ("counts_substitution", "counts_substitution")):
                    cmp(r + ":" + theirs, [S[ours]], [kw[theirs][r]])
                for ours, theirs in (("insertion_count_noncoding", "insertion_count_vectors_noncoding"),
                                     ("deletion_count_noncoding", "deletion_count_vectors_noncoding"),
                                     ("substitution_count_noncoding", "substitution_count_vectors_noncoding")):
                    cmp(r + ":" + theirs, V[ours], kw[theirs][r])
                for name in ("counts_modified_frameshift", "counts_modified_non_frameshift",
                             "counts_non_modified_non_frameshift", "counts_splicing_sites_modified"):
                    cmp(r + ":" + name, [S[name]], [kw[name][r]])
                if S["counts_total"] > 0:                   # length vectors leave main() as per-position means (:4394-4404)
                    for ln, ct, theirs in (("insertion_length", "insertion_count", "insertion_length_vectors"),
                                           ("deletion_length", "deletion_count", "deletion_length_vectors")):
                        mean = np.zeros(len(V[ln]))
                        mask = V[ct] > 0
                        mean[mask] = V[ln][mask] / V[ct][mask]
                        cmp(r + ":" + theirs, mean, kw[theirs][r])
                inframe, frameshift = blk.frame_histograms(r)
                for got, theirs in ((inframe, "hists_inframe"), (frameshift, "hists_frameshift")):
                    report["checked"] += 1
                    if dict(got) != {int(k): int(v) for k, v in kw[theirs][r].items()}:
                        bad.append(r + ":" + theirs)
                # the size Counters leave main() only as the plot arrays of :4348-4377
                H = blk.size_histograms(r)
                rr = kw["refs"][r]
                for key, xs, ys in (("substituted_n", "x_bins_mut", "y_values_mut"), ("inserted_n", "x_bins_ins", "y_values_ins"),
                                    ("deleted_n", "x_bins_del", "y_values_del")):
                    cmp(r + ":" + ys, [H[key][int(x)] for x in rr[xs]], rr[ys])
                    cmp(r + ":" + xs, np.arange(max(15, max(list(H[key].keys()) or [0])) + 1), rr[xs])
                L = len(rr["sequence"])
                cmp(r + ":hdensity", [H["effective_len"][int(x) + L] for x in rr["hlengths"]], rr["hdensity"])
                if "ref1_all_deletion_count_vectors" in kw and kw["ref1_all_deletion_count_vectors"]:
                    R1 = blk.vectors_ref1(r)
                    for k in ("insertion", "insertion_left", "deletion", "substitution", "indelsub"):
                        cmp(r + ":ref1_all_%s" % k, R1["ref1_all_%s_count" % k], kw["ref1_all_%s_count_vectors" % k][r])
                    for ch in "ACGTN-":
                        cmp(r + ":ref1_base_" + ch, R1["ref1_all_base_count_" + ch], kw["ref1_all_base_count_vectors"][r + "_" + ch])
            report["checked"] += 1
            if dict(blk.class_counts()) != {k: int(v) for k, v in kw["class_counts"].items()}:
                bad.append("class_counts")
            return orig_ctx(*a, **kw)

        CORE.CorePlotContext = ctx_spy
    os.makedirs(outdir, exist_ok=True)
    os.chdir(outdir)
    sys.argv = ["CRISPResso"] + argv + ["--suppress_plots", "--suppress_report", "-o", outdir]
    try:
        CORE.main()
    except SystemExit as e:
        if e.code not in (0, None):
            raise
    with open(vec_json, "w") as fh:
        json.dump(report, fh)


if __name__ == "__main__":
    main()
