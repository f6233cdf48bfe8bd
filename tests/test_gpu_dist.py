"""One process per GPU over NCCL (needs >= 2 GPUs; skipped on a single-GPU box): core.process_fastq_sharded with no explicit
engine -- every rank must take the GPU named by LOCAL_RANK, the count blocks meet in one all-reduce on device memory, the compact
per-read results are gathered -- must return on rank 0 what the reference's serial loop produced for the HDR fixture."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import golden_util as G  # noqa: E402

pytestmark = pytest.mark.gpu


def test_process_fastq_sharded_over_two_gpus_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rec = G.load("synth_hdr")
    fq = tmp_path / "hdr.fastq"
    with open(fq, "w") as fh:
        for k, s in enumerate(rec["reads"]):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    out = tmp_path / "sharded.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(HERE, "dist_worker2.py"), str(fq), str(out), "nccl"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.load(open(out))
    assert got["stats"] == rec["aln_stats"]
    assert got["keys"] == list(rec["variants"].keys())
    assert got["lost"] == sorted(rec["not_aligned"])
    assert got["classes"] == {s: v["class_name"] for s, v in rec["variants"].items()}
    assert got["payload_ok"]
    refs = G.refs_from(rec)
    for rname in rec["ref_names"]:
        seq = refs[rname]["sequence"]
        V = {k: np.asarray(v) for k, v in got["vec"][rname].items()}
        tot = int(sum(V["all_base_count_" + b][0] for b in "ACGTN-"))
        assert G.mod_count_text(seq, V, tot) == G.file_for(rec, rname, "Modification_count_vectors.txt")
