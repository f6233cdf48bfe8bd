"""Phase timing of the native FASTQ front end (C2B_FASTQ_VERBOSE) on the bench's 1 Mi-read FASTQ, plain and all-unique."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from crispresso2_b200 import fastq, synth
os.environ["C2B_FASTQ_VERBOSE"] = "1"
w = bench.Workload("single", 1 << 20, 0)
amp = w.refs["Reference"]["sequence"]
d = tempfile.mkdtemp(prefix="c2b_ing_")
for name, reads in (("bench batch", w.buf.reshape(-1, 250)),
                    ("all unique", synth.synth_reads_fast(np.random.default_rng(77), amp, 1 << 20, 250, sub_rate=0.02, cut=w.refs["Reference"]["cut_point"]))):
    fq = os.path.join(d, "r.fastq")
    synth.write_fastq_fast(fq, reads)
    for rep in range(6):
        dev = 0 if (rep >= 3 and os.environ.get("C2B_INGEST_GPU", "1") != "0") else None
        t0 = time.perf_counter()
        dd = fastq.dedup_file(fq, device=dev)
        dt = time.perf_counter() - t0
        print(("gpu  " if dev is not None else "host ") + "%s rep %d: %.3f s, %d unique of %d, %.2f M reads/s, %.2f GB/s" % (name, rep, dt, len(dd.counts), dd.n_reads, dd.n_reads / dt / 1e6, os.path.getsize(fq) / dt / 1e9), file=sys.stderr)
