"""Worker for tests/test_dist_gloo.py (gloo, CPU, emulator build of the engine) and tests/test_gpu_dist.py (nccl, one GPU per
rank, sm_100a library, NO explicit engine: every rank must pick the GPU named by LOCAL_RANK): world_size ranks run
core.process_fastq_sharded on the same FASTQ; rank 0 writes what it returned.
usage: dist_worker2.py <fastq> <out.json> [gloo|nccl] [golden case]"""
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
    import golden_util as G
    import parity_util as PU
    from crispresso2_b200 import core
    from crispresso2_b200.engine import Engine
    from oracle import oracle as O
    fq, out_path = sys.argv[1:3]
    backend = sys.argv[3] if len(sys.argv) > 3 else "gloo"
    if backend == "nccl":
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    else:
        dist.init_process_group("gloo")
    rank = dist.get_rank()
    case = sys.argv[4] if len(sys.argv) > 4 else "synth_hdr"
    rec = G.load(case)
    refs = G.refs_from(rec)
    args = PU.args_from(rec["params"])
    if not args.prime_editing_pegRNA_extension_seq:
        args.expected_hdr_amplicon_seq = refs[rec["ref_names"][1]]["sequence"]
    eng = Engine(lib_path=build_emu.build()) if backend == "gloo" else None       # nccl: the default engine of the rank's own GPU
    cache = {}
    stats, lost = core.process_fastq_sharded(fq, cache, rec["ref_names"], refs, args, [], os.path.dirname(out_path),
                                             engine=eng, aln_matrix=O.make_matrix())
    if rank == 0:
        blk = core.quantify(cache)
        summary = {"stats": stats, "keys": list(cache.keys()), "lost": sorted(lost),
                   "classes": {s: v["class_name"] for s, v in cache.items()},
                   "payload_ok": all(not G.payload_equal(rec["variants"][s]["variant_" + r], cache[s]["variant_" + r])
                                     for s in cache for r in rec["variants"][s]["aln_ref_names"]),
                   "vec": {r: {k: v.tolist() for k, v in blk.vectors(r).items()} for r in blk.ref_names},
                   "class_counts": blk.class_counts()}
        with open(out_path, "w") as fh:
            json.dump(summary, fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
