#!/usr/bin/env python
"""bench.py -- reads/sec aligned + classified (250 bp, single amplicon), BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W            the CUDA engine (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU path on the host cores

One step = one pass of the hot path over one batch of synthetic reads (configs[1]: 1M x 250 bp, one 250-bp
amplicon, every read aligned -- no dedup shortcut, so reads/s == DP problems/s):
  value : batch already resident in HBM, outputs left in HBM (device-pointer C-ABI entry)
  e2e   : the same batch through the host-pointer C-ABI call: pinned host buffers in, pinned host buffers out,
          H2D + kernel + D2H inside the timed region
Timed with CUDA events on the engine's stream, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 250
AMP_LEN = 250
ALG_BYTES_PER_READ = 250 + 2 * 256 + 32          # read + two aligned strings + record
ALG_INTOPS_PER_READ = 625_000                    # 10 ops x 250 x 250 cells


def make_workload(n_reads, seed):
    from crispresso2_b200 import synth
    rng = np.random.default_rng(seed)
    amp = synth.random_amplicon(np.random.default_rng(42), AMP_LEN)      # same amplicon on every rank
    ref = synth.amplicon_setup(amp)
    reads = synth.synth_reads_fast(rng, amp, n_reads, READ_LEN, cut=ref["cut_point"])
    return amp, ref, reads


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                t = [x.strip() for x in out.strip().split(",")]
                if len(t) >= 6:
                    self.rows.append(t)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_baseline_port(amp, ref, reads, seconds=12.0):
    """The oracle's C restatement (single thread, no Python in the loop) on a bounded sample."""
    import ctypes as C
    from oracle import oracle as O
    L = O.lib()
    m = np.ascontiguousarray(O.make_matrix())
    mask = np.zeros(len(amp) + 2, dtype=np.uint8)
    mask[ref["include_idxs"]] = 1
    gi = np.ascontiguousarray(ref["gap_incentive"], dtype=np.int64)
    n = 2000
    done, t0, chk = 0, time.time(), C.c_int64(0)
    while time.time() - t0 < seconds and done + n <= len(reads):
        blk = np.ascontiguousarray(reads[done:done + n])
        off = np.arange(n + 1, dtype=np.int64) * READ_LEN
        L.c2o_batch_align_classify(blk.ctypes.data, off.ctypes.data, n, amp.encode(), len(amp), m.ctypes.data, m.shape[1],
                                   gi.ctypes.data, -20, -2, mask.ctypes.data, C.byref(chk))
        done += n
    dt = time.time() - t0
    return {"value": done / dt, "unit": "reads/s", "cores": 1, "kind": "port",
            "sample": "%d reads of the same workload, oracle/c2_oracle.c global_align+find_indels, 1 thread, %.1f s" % (done, dt)}


def _ref_worker(job):
    amp, gi, inc, chunk = job
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from CRISPResso2 import CRISPResso2Align as A, CRISPRessoCOREResources as R
    m = A.make_matrix()
    t0 = time.time()
    for s in chunk:
        s1, s2, sc = A.global_align(s, amp, matrix=m, gap_incentive=gi, gap_open=-20, gap_extend=-2)
        R.find_indels_substitutions(s1, s2, inc)
    return len(chunk), time.time() - t0


def reference_arm(args):
    """The reference's own Cython global_align + find_indels_substitutions (compiled from /root/reference into
    oracle/_ref, unmodified) on all host cores; falls back to the oracle port when _ref is absent."""
    import multiprocessing as mp
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    amp, ref, reads = make_workload(200_000, 1234)
    have_ref = O.ref_modules() is not None
    per_step = []
    if have_ref:
        gi = np.ascontiguousarray(ref["gap_incentive"], dtype=np.int64)
        inc = [int(v) for v in ref["include_idxs"]]
        n_per_core = 600                                         # ~0.6 ms per read -> ~0.4 s per core per step
        strs = [r.tobytes().decode() for r in reads[:cores * n_per_core * (args.steps + args.warmup)]]
        with mp.get_context("fork").Pool(cores) as pool:
            pos = 0
            for step in range(args.warmup + args.steps):
                jobs = [(amp, gi, inc, strs[pos + c * n_per_core: pos + (c + 1) * n_per_core]) for c in range(cores)]
                pos += cores * n_per_core
                t0 = time.time()
                out = pool.map(_ref_worker, jobs)
                dt = time.time() - t0
                if step >= args.warmup:
                    per_step.append((sum(o[0] for o in out), dt))
        kind = "reference"
        sample = "%d reads per step (%d per core), reference Cython global_align+find_indels_substitutions via fork pool" % (
            cores * n_per_core, n_per_core)
    else:
        base = cpu_baseline_port(amp, ref, reads, seconds=10.0)
        per_step = [(base["value"] * 1.0, 1.0)]
        kind, cores, sample = "port", 1, base["sample"]
    n_done = sum(p[0] for p in per_step)
    t = sum(p[1] for p in per_step)
    val = n_done / t
    line = {"impl": "reference", "metric": "reads/sec aligned+classified (250 bp, single amplicon)", "value": val,
            "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * t / max(1, len(per_step)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "synthetic 1M x 250 bp reads, 1 amplicon (bounded sample per step)"},
            "cpu_baseline": {"value": val, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=1 << 20, help="reads per GPU per step (weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--edit-cap", type=int, default=8)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return 0

    import torch
    import torch.distributed as dist
    from crispresso2_b200 import _lib, dist as cdist
    from crispresso2_b200.engine import Engine
    from oracle import oracle as O

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.warmup < 3:
        args.warmup = 3

    n = args.reads
    amp, ref, reads = make_workload(n, 1000 + rank)
    eng = Engine(local)
    eng.configure({"Reference": ref}, ["Reference"], O.make_matrix(), -20, -2, 5, 2, 0, "ACGTN", args.edit_cap)
    W = eng.string_width(READ_LEN)
    L = eng.L
    stream = torch.cuda.ExternalStream(L.c2b_stream(eng.h), device=dev)

    # ---- device-resident buffers (value leg) -------------------------------------------------------
    d_reads = torch.from_numpy(reads.reshape(-1)).to(dev)
    d_off = (torch.arange(n + 1, dtype=torch.int64, device=dev) * READ_LEN)
    d_recs = torch.empty(n * 16, dtype=torch.uint8, device=dev)
    d_alns = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_str = torch.empty(n * 2 * W, dtype=torch.uint8, device=dev)
    d_ed = torch.empty(n * args.edit_cap * 8, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)

    def step_device():
        rc = L.c2b_align_batch_device(eng.h, d_reads.data_ptr(), d_off.data_ptr(), n, READ_LEN, None, None, None,
                                      d_recs.data_ptr(), d_alns.data_ptr(), d_str.data_ptr(), d_ed.data_ptr())
        if rc != 0:
            raise RuntimeError(L.c2b_last_error(eng.h).decode())
        if world > 1:                       # merge of the count block: the path's only exchange (NCCL, NVLink)
            cdist.allreduce_counts(eng)

    def barrier():
        if world > 1:
            dist.barrier()
        eng.sync()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        eng.counts_reset()
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.launch_count()
    kernel_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.counts_reset()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
        kernel_ms.append(eng.last_kernel_ms())         # CUDA events around the kernel on its own stream
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    pair_items, single_items = eng.path_counts()

    # ---- parity gate on this very batch: a sample of the device results against the oracle ------------
    recs = np.frombuffer(d_recs.cpu().numpy().tobytes(), dtype=_lib.REC_DTYPE)
    alns = np.frombuffer(d_alns.cpu().numpy().tobytes(), dtype=_lib.ALN_DTYPE)
    gate_ok = True
    if rank == 0:
        strs = d_str[: 64 * 2 * W].cpu().numpy().reshape(64, 2, W)
        m = O.make_matrix()
        for i in range(64):
            s1, s2, nm, nl = O.global_align_raw(reads[i].tobytes().decode(), amp, m, ref["gap_incentive"], -20, -2)
            a = alns[i]
            g1 = strs[i, 0, W - a["aln_len"]:].tobytes().decode()
            g2 = strs[i, 1, W - a["aln_len"]:].tobytes().decode()
            if (g1, g2, int(a["n_match"]), int(a["aln_len"])) != (s1, s2, nm, nl):
                gate_ok = False
    aligned_frac = float((recs["best_score_milli"] > 0).mean())

    # ---- end-to-end leg: host buffers through c2b_align_batch --------------------------------------
    def pinned(nbytes, dtype=np.uint8):
        p = L.c2b_host_alloc(nbytes)
        if not p:
            raise MemoryError("c2b_host_alloc")
        import ctypes as C
        return np.frombuffer((C.c_uint8 * nbytes).from_address(p), dtype=np.uint8).view(dtype), p

    h_reads, p1 = pinned(n * READ_LEN)
    h_reads[:] = reads.reshape(-1)
    h_off, p2 = pinned((n + 1) * 8, np.int64)
    h_off[:] = np.arange(n + 1, dtype=np.int64) * READ_LEN
    h_recs, p3 = pinned(n * 16)
    h_alns, p4 = pinned(n * 32)
    h_str, p5 = pinned(n * 2 * W)
    h_ed, p6 = pinned(n * args.edit_cap * 8)
    h2d = n * READ_LEN + (n + 1) * 8
    d2h = n * 16 + n * 32 + n * 2 * W + n * args.edit_cap * 8

    def step_e2e():
        rc = L.c2b_align_batch(eng.h, h_reads.ctypes.data, h_off.ctypes.data, n, None, None, None, h_recs.ctypes.data,
                               h_alns.ctypes.data, h_str.ctypes.data, h_ed.ctypes.data)
        if rc != 0:
            raise RuntimeError(L.c2b_last_error(eng.h).decode())
        if world > 1:
            cdist.allreduce_counts(eng)

    e2e_steps = max(2, min(args.steps, 3))
    step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        step_e2e()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if sampler:                                  # sampled across both timed regions (device-resident and end-to-end)
        sampler.stop_flag = True
        sampler.join(timeout=3)
    h_al = np.frombuffer(h_alns.tobytes(), dtype=_lib.ALN_DTYPE)
    e2e_gate = bool((h_al["n_match"] == alns["n_match"]).all())
    # the library copies back only the right-hand Wt bytes of each W-byte string slot (Wt = widest alignment, rounded to 32)
    Wt = min(W, (int(h_al["aln_len"].max()) + 31) & ~31)
    d2h = n * 16 + n * 32 + n * 2 * Wt + n * args.edit_cap * 8
    for p in (p1, p2, p3, p4, p5, p6):
        L.c2b_host_free(p)

    # ---- reduce over ranks ----------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms, float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, k_ms = [float(x) for x in t.cpu()]
    total_reads = n * world
    value = total_reads * args.steps / (dev_ms / 1000.0)
    e2e_val = total_reads * e2e_steps / (e2e_ms / 1000.0)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = n * ALG_BYTES_PER_READ / (k_ms / 1000.0) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as fh:
                    traffic = json.load(fh).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        sm_clk = None
        clocks = sampler.summary() if sampler else {}
        sm_clk = clocks.get("sm_mhz") or 1965
        int_peak = 148 * 128 * sm_clk * 1e6            # int32 lanes/clk/SM (ALU + FMA pipes, B300_MICROARCH.md) x clock
        line = {
            "metric": "reads/sec aligned+classified (250 bp, single amplicon)", "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "synthetic 1M x 250 bp reads, 1 amplicon (BASELINE.json configs[1]), every read aligned",
                       "reads_per_gpu_per_step": n, "amplicon_len": AMP_LEN, "read_len": READ_LEN, "parallelism": "read-shard x%d" % world,
                       "l2": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2" % ((h2d + d2h) / 1e9),
                       "edit_cap": args.edit_cap, "aligned_fraction": aligned_frac,
                       "packed_pair_items": pair_items, "single_items": single_items, "band_reruns": eng.band_reruns(), "ring_pairs": eng.ring_counts()[0], "ring_fallbacks": eng.ring_counts()[1], "parity_gate": bool(gate_ok and e2e_gate)},
            "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "pipeline": ("launch per chunk" if os.environ.get("C2B_STREAMED") == "0" else
                                 "one persistent launch per batch, read bytes streamed in behind it (c2b_align_batch)")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "measured" if peaks else "fallback",
                         "kernel": "c2b_align_classify_kernel", "kernel_ms": k_ms, "alg_bytes_per_read": ALG_BYTES_PER_READ,
                         "secondary_int32": {"alg_ops_per_read": ALG_INTOPS_PER_READ,
                                             "note": "ops of the reference's full-matrix algorithm; the ring-banded DP evaluates 72 of 250 columns per row",
                                             "achieved_tops": n * ALG_INTOPS_PER_READ / (k_ms / 1000.0) / 1e12,
                                             "peak_tops_at_observed_clock": int_peak / 1e12,
                                             "frac": n * ALG_INTOPS_PER_READ / (k_ms / 1000.0) / int_peak}},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_port(amp, ref, reads)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
