"""Paired-end merge mode on the GPU box: the UNMODIFIED reference's process_paired_fastq (baseline/_ref,
CRISPRessoCORE.py:1245-1733, serial branch) against crispresso2_b200.paired.process_paired_fastq on the same two FASTQ files --
same variantCache (keys, counts, classes, aligned strings) and aln_stats required, pairs/s of both printed as one JSON line.
usage: python tools/paired_bench.py [n_pairs]"""
import json
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from baseline import ref_shim
from crispresso2_b200 import core, paired, synth


def write_pairs(p1, p2, amp, n, seed=5):
    comp = str.maketrans("ACGT", "TGCA")
    rc = lambda s: s.translate(comp)[::-1]
    rnd = random.Random(seed)
    L, cut = 150, len(amp) // 2
    with open(p1, "w") as f1, open(p2, "w") as f2:
        for k in range(n):
            t, u = amp, rnd.random()
            if u < 0.25:
                d = rnd.randint(1, 12)
                t = t[:cut - d // 2] + t[cut - d // 2 + d:]
            elif u < 0.35:
                t = t[:cut] + "".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 5))) + t[cut:]
            elif u < 0.55:
                p = rnd.randint(20, len(t) - 20)
                t = t[:p] + rnd.choice("ACGT") + t[p + 1:]
            if rnd.random() < 0.15:
                t = rc(t)
            m1, m2 = t[:L], rc(t[-L:])
            if rnd.random() < 0.2:
                p = rnd.randint(len(t) - L + 2, L - 3)
                m1 = m1[:p] + rnd.choice("ACGT") + m1[p + 1:]
            q1 = "".join(rnd.choice("II5#") for _ in m1)
            q2 = "".join(rnd.choice("II5#") for _ in m2)
            f1.write("@p%d\n%s\n+\n%s\n" % (k, m1, q1))
            f2.write("@p%d\n%s\n+\n%s\n" % (k, m2, q2))


def run(n, device=0):
    """-> dict of the comparison (None when baseline/_ref is absent)"""
    if not ref_shim.available():
        return None
    CORE = ref_shim.load_core()
    from CRISPResso2 import CRISPRessoShared, CRISPResso2Align
    import logging
    for name in list(logging.root.manager.loggerDict):
        if name.startswith("CRISPResso"):
            logging.getLogger(name).setLevel(logging.ERROR)
    rng = np.random.default_rng(4)
    amp = synth.random_amplicon(rng, 223)
    ref = synth.amplicon_setup(amp)
    d = tempfile.mkdtemp(prefix="c2b_pair_")
    r1, r2 = os.path.join(d, "R1.fastq"), os.path.join(d, "R2.fastq")
    write_pairs(r1, r2, amp, n)
    args = CRISPRessoShared.getCRISPRessoArgParser("Core").parse_args(["-r1", r1, "-r2", r2, "-a", amp, "--crispresso_merge"])
    args.n_processes = "1"
    refs, names = {"Reference": ref}, ["Reference"]
    out_fd = os.dup(1)
    os.dup2(2, 1)                                            # the reference logs to stdout
    try:
        c_ref = {}
        t0 = time.perf_counter()
        st_ref, lost_ref = CORE.process_paired_fastq(r1, r2, c_ref, names, refs, args, [], d)
        t_ref = time.perf_counter() - t0
        eng = core.get_engine(device)
        loc = os.path.join(CORE._ROOT, args.needleman_wunsch_aln_matrix_loc)
        m = core.read_matrix(loc)
        best = None
        for rep in range(2):                                 # first pass pays the engine's start-up
            c_gpu = {}
            t0 = time.perf_counter()
            st_gpu, lost_gpu = paired.process_paired_fastq(CORE.process_paired_fastq, CRISPResso2Align, eng, r1, r2, c_gpu, names, refs,
                                                           args, [], d, aln_matrix=m)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        sys.stdout.flush()
        os.dup2(out_fd, 1)
        os.close(out_fd)
    memo = paired.process_paired_fastq.last_memo
    same = (st_ref == st_gpu and list(c_ref) == list(c_gpu) and sorted(lost_ref) == sorted(lost_gpu)
            and all(c_ref[k]["count"] == c_gpu[k]["count"] and c_ref[k]["class_name"] == c_gpu[k]["class_name"]
                    and c_ref[k]["ref_aln_details"] == c_gpu[k]["ref_aln_details"] for k in c_ref))
    import shutil
    shutil.rmtree(d, ignore_errors=True)
    return ({"pairs": n, "identical_to_reference": bool(same), "reference_pairs_per_s": n / t_ref, "reference_seconds": t_ref,
                      "b200_pairs_per_s": n / best, "b200_seconds": best, "speedup": t_ref / best,
                      "global_align_calls_served_from_the_batch": memo.hits, "single_gpu_calls": memo.misses,
                      "distinct_sequences_in_the_batch": len(memo.index), "aligned_unique": st_gpu["N_COMPUTED_ALN"],
             "call": "CRISPRessoCORE.process_paired_fastq (unmodified reference, serial branch) vs crispresso2_b200.paired.process_paired_fastq "
                     "on the same two FASTQ files of 150 bp mates; best of 2 for the GPU path"})


if __name__ == "__main__":
    out = run(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(os.environ.get("LOCAL_RANK", "0")))
    print(json.dumps(out))
