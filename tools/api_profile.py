"""Where the wall clock of core.process_fastq goes (GPU box): cProfile of one call on the bench's all-unique FASTQ.
usage: python tools/api_profile.py [n_reads] [sub_rate]"""
import cProfile, os, pstats, sys, tempfile, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from crispresso2_b200 import core, synth
from crispresso2_b200.engine import Engine
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
sub = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
w = bench.Workload("single", 1024, 0)
amp = w.refs["Reference"]["sequence"]
reads = synth.synth_reads_fast(np.random.default_rng(77), amp, n, 250, sub_rate=sub, cut=w.refs["Reference"]["cut_point"])
d = tempfile.mkdtemp(prefix="c2b_prof_")
fq = os.path.join(d, "r.fastq")
synth.write_fastq_fast(fq, reads)
a = types.SimpleNamespace(**vars(w.params))
a.use_legacy_insertion_quantification = False; a.prime_editing_pegRNA_scaffold_seq = ""; a.prime_editing_pegRNA_extension_seq = ""
a.needleman_wunsch_aln_matrix_loc = "EDNAFULL"; a.n_processes = "1"
eng = Engine(0)
m = O.make_matrix()
for rep in range(2):
    cache = {}
    t0 = time.perf_counter()
    core.process_fastq(fq, cache, w.ref_names, w.refs, a, [], d, engine=eng, aln_matrix=m)
    print("pass %d: %.3f s  %s" % (rep, time.perf_counter() - t0, {k: round(v, 4) for k, v in core.last_timings.items() if isinstance(v, float)}))
cache = {}
pr = cProfile.Profile()
pr.enable()
core.process_fastq(fq, cache, w.ref_names, w.refs, a, [], d, engine=eng, aln_matrix=m)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
