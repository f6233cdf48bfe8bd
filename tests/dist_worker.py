"""Worker for tests/test_dist_gloo.py: world_size ranks (gloo, CPU) each run the emulator build of the engine on
their shard of the reads, then merge the count blocks with the one all-reduce the path has."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def main():
    import torch.distributed as dist
    import build_emu
    from crispresso2_b200 import core, dist as cdist, synth
    from crispresso2_b200.engine import Engine
    from oracle import oracle as O
    out_path = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(99)
    amp = synth.random_amplicon(rng, 120)
    ref = synth.amplicon_setup(amp, guide_start=50)
    reads = [r.tobytes().decode() for r in synth.synth_reads(rng, amp, 90, 120, sub_rate=0.01, rc_frac=0.1, cut=ref["cut_point"])]
    uniques = list(dict.fromkeys(reads))
    counts = [reads.count(u) for u in uniques]
    weights = core.merge_weights(uniques, counts)            # needs the global unique table: computed before sharding
    lo, hi = cdist.shard_bounds(len(uniques), rank, world)
    eng = Engine(lib_path=build_emu.build())
    eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2)
    eng.counts_reset()
    eng.align(uniques[lo:hi], count=np.asarray(counts[lo:hi], dtype=np.int32), qweight=np.asarray(weights[lo:hi], dtype=np.int32))
    merged = cdist.allreduce_counts(eng)
    if rank == 0:
        blk = eng.counts(raw=merged)
        V = blk.vectors("Reference")
        with open(out_path, "w") as fh:
            json.dump({"world": world, "vectors": {k: v.tolist() for k, v in V.items()}, "scalars": blk.scalars("Reference"), "classes": blk.class_counts(),
                       "sizes": {k: {str(a): b for a, b in v.items()} for k, v in blk.size_histograms("Reference").items()},
                       "amp": amp, "reads": reads}, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
