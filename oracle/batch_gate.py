"""Large-sample parity gate: device results of one batch against the CPU oracle, read by read -- TEST INFRASTRUCTURE ONLY.

Used by bench.py (the in-bench parity gate over >= 100k reads of the timed batch, SURVEY.md 8(d) last row) and by
tests/test_gpu_parity.py.  The product never imports this module.

The caller hands over what the engine produced for a batch (c2b_read_rec / c2b_aln_rec / aligned strings / edit lists,
exactly as they came back through the C ABI) and the batch's inputs.  `run()` writes them to a scratch directory and
starts a CLEAN python process (no CUDA context, so fork is safe) that splits the reads over a fork pool.  Every worker
  * runs the oracle's get_new_variant_object restatement (oracle.new_variant: seed test, global_align per strand and
    reference, best reference, find_indels_substitutions, classification) on each of its reads,
  * builds the product's own variant dict for the same read from the device arrays (crispresso2_b200.core._variant_from,
    the code a caller of process_fastq sees) and compares every field: scores, every (aligned read, aligned reference,
    score) of ref_aln_details, best_match_score, aln_ref_names, class_name, best_match_name and all payload slots,
  * runs the oracle's quantification loop (oracle.count_vectors, + ref1_vectors in HDR mode) over its chunk.
The chunk quantifications are additive; their sum is returned so the caller can compare it with the engine's count
block of the same reads (`compare_block`).
"""
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_G = {}


def _norm(g):
    if hasattr(g, "tolist"):
        g = g.tolist()
    if isinstance(g, tuple):
        g = list(g)
    if isinstance(g, list):
        g = [list(x) if isinstance(x, tuple) else x for x in g]
    return g


def _payload_diff(want, got):
    return [k for k, w in want.items() if _norm(got[k]) != _norm(w)]


def _worker(span):
    lo, hi = span
    from oracle import oracle as O
    from crispresso2_b200 import core, _lib
    from crispresso2_b200.engine import BatchResult
    g = _G
    refs, ref_names, params, matrix = g["refs"], g["ref_names"], g["params"], g["matrix"]
    buf, off, rid = g["buf"], g["off"], g["ref_id"]
    res = BatchResult(g["recs"][lo:hi], g["alns"][lo:hi], g["strings"][lo:hi], g["edits"][lo:hi] if g["edits"] is not None else None, g["W"])
    res.flags = g["flags"]
    if rid is not None:                                     # compact Pooled outputs: winner bit of the read's own amplicon -> bit 0
        recs = res.recs.copy()
        recs["winner_mask"] = (recs["winner_mask"] >> (rid[lo:hi].astype(np.uint32) & 31)) & 1
        res.recs = recs
    bad, n_bad, n_skipped = [], 0, 0
    cache = {}                                              # per (reference set): seq -> oracle variant with count
    text = buf[off[lo]:off[hi]].tobytes().decode("latin-1")
    base = int(off[lo])
    for i in range(lo, hi):
        seq = text[int(off[i]) - base:int(off[i + 1]) - base]
        names = ref_names if rid is None else [ref_names[int(rid[i])]]
        key = (seq, names[0]) if rid is not None else seq
        hit = cache.get(key)
        if hit is not None:
            want = hit
            want["count"] += 1
        else:
            try:
                want = O.new_variant(params, seq, refs, names, matrix)
            except O.OracleUndefined:
                n_skipped += 1
                continue
            cache[key] = want
        try:
            got = core._variant_from(res, i - lo, seq, names, refs)
        except OverflowError:
            bad.append((i, "edit_overflow")); n_bad += 1
            continue
        diff = []
        for k in ("aln_scores", "best_match_score"):
            if got[k] != want[k]:
                diff.append(k)
        if [tuple(x) for x in got["ref_aln_details"]] != [tuple(x) for x in want["ref_aln_details"]]:
            diff.append("ref_aln_details")
        if want["best_match_score"] > 0:
            for k in ("aln_ref_names", "class_name", "best_match_name"):
                if got.get(k) != want[k]:
                    diff.append(k)
            if not diff:
                for r in want["aln_ref_names"]:
                    d = _payload_diff(want["variant_" + r], got["variant_" + r])
                    diff.extend("variant_%s.%s" % (r, k) for k in d)
        if diff:
            n_bad += 1
            if len(bad) < 5:
                bad.append((i, diff))
    # quantification of this chunk (additive over chunks); Pooled: one cache per amplicon
    out = {"n": hi - lo, "n_bad": n_bad, "bad": bad, "n_skipped": n_skipped, "quant": {}}
    groups = {}
    if rid is None:
        groups[None] = {s: v for s, v in cache.items() if v["best_match_score"] > 0}
    else:
        for (s, name), v in cache.items():
            if v["best_match_score"] > 0:
                groups.setdefault(name, {})[s] = v
    for name, c in groups.items():
        names = ref_names if name is None else [name]
        extras = {}
        vec, sca, classes, total = O.count_vectors(c, refs, names, params, extras)
        q = {"vec": vec, "sca": sca, "classes": classes, "total": total, "extras": extras}
        if getattr(params, "expected_hdr_amplicon_seq", "") and name is None:
            q["ref1"] = O.ref1_vectors(c, refs, names, params)
        out["quant"][name] = q
    return out


def _merge(acc, q):
    from collections import Counter
    if acc is None:
        return q
    for r in q["vec"]:
        for n, v in q["vec"][r].items():
            acc["vec"][r][n] = acc["vec"][r][n] + v
        for n, v in q["sca"][r].items():
            acc["sca"][r][n] += v
        for n, v in q["extras"][r].items():
            acc["extras"][r][n] = Counter(acc["extras"][r][n]) + Counter(v)
            if n in ("hists_inframe", "hists_frameshift"):
                acc["extras"][r][n].setdefault(0, 0)
    for k, v in q["classes"].items():
        acc["classes"][k] = acc["classes"].get(k, 0) + v
    acc["total"] += q["total"]
    if "ref1" in q:
        for r in q["ref1"]:
            for n, v in q["ref1"][r].items():
                acc["ref1"][r][n] = acc["ref1"][r][n] + v
    return acc


def _main(workdir, n_workers):
    import multiprocessing as mp
    with open(os.path.join(workdir, "meta.pkl"), "rb") as fh:
        meta = pickle.load(fh)
    g = _G
    g.update(meta)
    for name in ("buf", "off", "recs", "alns", "strings", "edits", "ref_id"):
        p = os.path.join(workdir, name + ".npy")
        g[name] = np.load(p, mmap_mode="r") if os.path.exists(p) else None
    n = len(g["off"]) - 1
    per = max(64, -(-n // (n_workers * 4)))
    spans = [(a, min(a + per, n)) for a in range(0, n, per)]
    t0 = time.time()
    if n_workers > 1:
        with mp.get_context("fork").Pool(n_workers) as pool:
            outs = pool.map(_worker, spans)
    else:
        outs = [_worker(s) for s in spans]
    quant = {}
    res = {"n": 0, "n_bad": 0, "bad": [], "n_skipped": 0}
    for o in outs:
        res["n"] += o["n"]; res["n_bad"] += o["n_bad"]; res["n_skipped"] += o["n_skipped"]
        res["bad"].extend(o["bad"][:3])
        for name, q in o["quant"].items():
            quant[name] = _merge(quant.get(name), q)
    res["bad"] = res["bad"][:10]
    res["seconds"] = time.time() - t0
    res["workers"] = n_workers
    with open(os.path.join(workdir, "result.pkl"), "wb") as fh:
        pickle.dump({"summary": res, "quant": quant}, fh)
    print(json.dumps({k: (v if k != "bad" else [str(b) for b in v]) for k, v in res.items()}))


def run(buf, off, refs, ref_names, params, matrix, recs, alns, strings, edits, W, flags=0, ref_id=None, n_workers=None,
        keep=False):
    """-> (summary dict, quant): summary['n_bad'] == 0 means every read's device result equals the oracle's.
    quant[None] (or quant[ref_name] for Pooled batches) = the oracle's quantification of the batch."""
    n_workers = n_workers or os.cpu_count() or 1
    d = tempfile.mkdtemp(prefix="c2b_gate_")
    try:
        np.save(os.path.join(d, "buf.npy"), np.ascontiguousarray(buf, dtype=np.uint8))
        np.save(os.path.join(d, "off.npy"), np.ascontiguousarray(off, dtype=np.int64))
        np.save(os.path.join(d, "recs.npy"), np.ascontiguousarray(recs))
        np.save(os.path.join(d, "alns.npy"), np.ascontiguousarray(alns))
        np.save(os.path.join(d, "strings.npy"), np.ascontiguousarray(strings))
        if edits is not None:
            np.save(os.path.join(d, "edits.npy"), np.ascontiguousarray(edits))
        if ref_id is not None:
            np.save(os.path.join(d, "ref_id.npy"), np.ascontiguousarray(ref_id, dtype=np.int32))
        with open(os.path.join(d, "meta.pkl"), "wb") as fh:
            pickle.dump({"refs": refs, "ref_names": list(ref_names), "params": params, "matrix": np.asarray(matrix),
                         "W": int(W), "flags": int(flags)}, fh)
        env = dict(os.environ)
        env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
        env.pop("CUDA_VISIBLE_DEVICES", None)
        p = subprocess.run([sys.executable, "-m", "oracle.batch_gate", d, str(n_workers)], capture_output=True, text=True, env=env, cwd=ROOT)
        if p.returncode != 0:
            raise RuntimeError("batch_gate worker failed:\n" + p.stdout[-2000:] + p.stderr[-4000:])
        with open(os.path.join(d, "result.pkl"), "rb") as fh:
            out = pickle.load(fh)
        return out["summary"], out["quant"]
    finally:
        if not keep:
            import shutil
            shutil.rmtree(d, ignore_errors=True)


def _time_worker(span):
    lo, hi = span
    from oracle import oracle as O
    g = _G
    buf, off, rid = g["buf"], g["off"], g["ref_id"]
    t0 = time.time()
    for i in range(lo, hi):
        seq = buf[off[i]:off[i + 1]].tobytes().decode("latin-1")
        names = g["ref_names"] if rid is None else [g["ref_names"][int(rid[i])]]
        try:
            O.new_variant(g["params"], seq, g["refs"], names, g["matrix"])
        except O.OracleUndefined:
            pass
    return time.time() - t0


def time_oracle(buf, off, refs, ref_names, params, ref_id, n, n_workers=None):
    """Wall time of the oracle's per-read path over the first n reads on a fork pool (bench.py's cpu_baseline for the
    multi-amplicon configs).  Call BEFORE any CUDA context exists in this process, or accept fork-after-CUDA (workers never
    touch CUDA)."""
    import multiprocessing as mp
    from oracle import oracle as O
    n_workers = n_workers or os.cpu_count() or 1
    _G.update(buf=np.asarray(buf), off=np.asarray(off), ref_id=ref_id, refs=refs, ref_names=list(ref_names), params=params,
              matrix=O.make_matrix())
    per = max(16, -(-n // n_workers))
    spans = [(a, min(a + per, n)) for a in range(0, n, per)]
    with mp.get_context("fork").Pool(n_workers) as pool:
        pool.map(_time_worker, [(0, 1)] * n_workers)          # warm-up: imports, liboracle.so
        t0 = time.time()
        pool.map(_time_worker, spans, chunksize=1)
        dt = time.time() - t0
    return {"seconds": dt, "workers": n_workers}


def compare_block(block, q, ref_names, hdr=False, per_amplicon=False):
    """Engine count block (crispresso2_b200.counts.CountBlock) against a merged oracle quantification -> list of mismatches.
    per_amplicon: `q` is the quantification of the reads of `ref_names` alone (Pooled: one run per amplicon), so only their
    class labels are compared."""
    from oracle import oracle as O
    bad = []
    got_classes = block.class_counts()
    if per_amplicon:
        got_classes = {k: v for k, v in got_classes.items() if any(k.startswith(r + "_") for r in ref_names)}
    if got_classes != {k: v for k, v in q["classes"].items() if v}:
        bad.append(("class_counts", got_classes, q["classes"]))
    for r in ref_names:
        V = block.vectors(r)
        for name in O.VECTOR_NAMES:
            if not (V[name] == q["vec"][r][name]).all():
                bad.append((r, name))
        S = block.scalars(r)
        for name in O.SCALAR_NAMES:
            if S[name] != q["sca"][r][name]:
                bad.append((r, name, S[name], q["sca"][r][name]))
        H = block.size_histograms(r)
        for name in ("inserted_n", "deleted_n", "substituted_n", "effective_len"):
            if {k: v for k, v in H[name].items() if v} != {k: v for k, v in q["extras"][r][name].items() if v}:
                bad.append((r, name))
        inframe, frameshift = block.frame_histograms(r)
        for got, name in ((inframe, "hists_inframe"), (frameshift, "hists_frameshift")):
            if {k: v for k, v in got.items() if v} != {k: v for k, v in q["extras"][r][name].items() if v}:
                bad.append((r, name))
    if hdr and "ref1" in q:
        for r in ref_names[1:]:
            R1 = block.vectors_ref1(r)
            for name, v in q["ref1"][r].items():
                if not (R1[name] == v).all():
                    bad.append((r, name))
    return bad


if __name__ == "__main__":
    _main(sys.argv[1], int(sys.argv[2]))
