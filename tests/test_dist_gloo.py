"""N>1 path on CPU: world_size-2 gloo run of the sharded engine + all-reduce of the count block, against the
oracle's single-process quantification."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    from crispresso2_b200.dist import shard_bounds, shard_by_cells
    for n in (0, 1, 7, 64, 1001):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[k][1] == spans[k + 1][0] for k in range(w - 1))
    lens = np.array([50, 300, 120, 250, 250, 75, 299, 60])
    parts = [shard_by_cells(lens, r, 3) for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(8))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= 300


def test_two_ranks_gloo_merge_matches_oracle(tmp_path):
    out = tmp_path / "merged.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py"), str(out)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.load(open(out))
    assert got["world"] == 2
    from crispresso2_b200 import synth
    from oracle import oracle as O
    ref = synth.amplicon_setup(got["amp"], guide_start=50)
    refs = {"Reference": ref}
    cache, stats, lost = O.process_reads(got["reads"], refs, ["Reference"], O.Params(), O.make_matrix())
    extras = {}
    vec, sca, classes, total = O.count_vectors(cache, refs, ["Reference"], O.Params(), extras)
    assert got["classes"] == classes
    for name in ("inserted_n", "deleted_n", "substituted_n", "effective_len"):
        assert got["sizes"][name] == {str(a): b for a, b in extras["Reference"][name].items()}, name
    for name in O.VECTOR_NAMES:
        assert got["vectors"][name] == vec["Reference"][name].tolist(), name
    for name in O.SCALAR_NAMES:
        assert got["scalars"][name] == sca["Reference"][name], name


@pytest.mark.parametrize("case", ["synth_hdr", "fanc_pe_scaffold"])
def test_process_fastq_sharded_over_two_ranks_equals_single_process(tmp_path, case):
    """The drop-in process_fastq with unique reads sharded over 2 ranks (gloo; count block all-reduced, variants gathered):
    every rank must hold what the reference's serial loop produces -- checked on the HDR golden fixture and on the prime-editing
    one (the 'Scaffold-incorporated' segments are summed over ranks next to the block)."""
    import golden_util as G
    import parity_util as PU
    rec = G.load(case)
    fq = tmp_path / "hdr.fastq"
    with open(fq, "w") as fh:
        for k, s in enumerate(rec["reads"]):
            fh.write("@r%d\n%s\n+\n%s\n" % (k, s, "I" * len(s)))
    out = tmp_path / "sharded.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "dist_worker2.py"), str(fq), str(out), "gloo", case]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.load(open(out))
    assert got["stats"] == rec["aln_stats"]
    assert got["keys"] == list(rec["variants"].keys())
    assert got["lost"] == sorted(rec["not_aligned"])
    assert got["classes"] == {s: v["class_name"] for s, v in rec["variants"].items()}
    assert got["payload_ok"]
    refs = G.refs_from(rec)
    names = list(rec["ref_names"])
    if case == "fanc_pe_scaffold":
        names.append("Scaffold-incorporated")
        refs["Scaffold-incorporated"] = refs["Prime-edited"]
        assert "Scaffold-incorporated" in got["class_counts"]
    for rname in names:
        seq = refs[rname]["sequence"]
        V = {k: np.asarray(v) for k, v in got["vec"][rname].items()}
        tot = int(V["all_base_count_A"][0] + V["all_base_count_C"][0] + V["all_base_count_G"][0] + V["all_base_count_T"][0]
                  + V["all_base_count_N"][0] + V["all_base_count_-"][0])
        assert G.mod_count_text(seq, V, tot) == G.file_for(rec, rname, "Modification_count_vectors.txt")
