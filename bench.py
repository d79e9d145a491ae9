#!/usr/bin/env python
"""bench.py -- points assigned per second per Lloyd iteration (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A step is ONE Lloyd iteration over the resident sparsified matrix: K1 masked distance +
argmin (+ fp64 re-evaluation of uncertified columns), K2 per-cluster sums/counts, (N>1: one
NCCL all-reduce of the partials), K3 centre finalisation, and the read-back of the iteration
statistics the host needs for the stop rule.  Workload at N=1: BASELINE.json configs[1]
(n=1e7, p=784, K=10, 10% nnz => 78 stored entries per point, fp32); for N>1 every rank holds
a shard of that shape (weak scaling, columns sharded by rank).  Inputs are synthetic, generated
directly in sparsified form on the device and are far larger than L2,
so no explicit flush is needed between iterations.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n per GPU, p, K, stored entries per point)
    "config2": dict(n=10_000_000, p=784, K=10, m=78,
                    label="n=1e7 p=784 k=10 10% nnz (78/col) fp32 sparsified Gaussian mixture"),
    "config3": dict(n=12_500_000, p=1024, K=64, m=51,
                    label="n=1e8/8 per GPU p=1024 k=64 5% nnz (51/col) fp32 sparsified Gaussian mixture"),
    # the one run the reference publishes timings for (figs/slides_experiment3.jpg: N=9,631,605, p=784,
    # gamma=0.05, digits 0/3/9 => K=3): assignments 1.3 s, centre update 5.7 s on unstated hardware
    "published_mnist3": dict(n=9_631_605, p=784, K=3, m=39,
                             label="reference's published run shape: n=9,631,605 p=784 k=3 5% nnz (39/col), synthetic values"),
    "tiny": dict(n=200_000, p=784, K=10, m=78, label="tiny debug shape"),
}
METRIC = "points assigned/sec per Lloyd iter"
UNIT = "points/s"
GEN_CHUNK = 500_000


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# --------------------------------------------------------------------------- data
def gen_shard_device(dev, n, p, m, K, col0, seed=2024):
    """Sparsified mixture shard generated on the GPU: columns [col0, col0+n).  Returns torch
    tensors (colptr int64 [n+1], rowidx int32 [n*m], val float32 [n*m]) and (mu, start)."""
    import torch
    g0 = torch.Generator(device="cpu").manual_seed(1234)
    mu = torch.randn(p, K, generator=g0, dtype=torch.float64)
    g1 = torch.Generator(device="cpu").manual_seed(99)
    start = mu + 0.05 * torch.randn(p, K, generator=g1, dtype=torch.float64)
    mu_d = mu.to(dev, torch.float32)
    rowidx = torch.empty(n * m, dtype=torch.int32, device=dev)
    val = torch.empty(n * m, dtype=torch.float32, device=dev)
    scale = float(p) / float(m)
    for c0 in range(0, n, GEN_CHUNK):
        c = min(GEN_CHUNK, n - c0)
        g = torch.Generator(device=dev).manual_seed(seed + (col0 + c0) // GEN_CHUNK)
        keys = torch.rand(c, p, generator=g, device=dev)
        rows = keys.topk(m, dim=1, largest=False).indices
        del keys
        rows, _ = rows.sort(dim=1)
        lab = (torch.arange(col0 + c0, col0 + c0 + c, device=dev) % K)
        v = mu_d[rows, lab[:, None]] + 0.1 * torch.randn(c, m, generator=g, device=dev)
        rowidx[c0 * m:(c0 + c) * m] = rows.reshape(-1).to(torch.int32)
        val[c0 * m:(c0 + c) * m] = (v * scale).reshape(-1)
        del rows, v, lab
    colptr = torch.arange(n + 1, dtype=torch.int64, device=dev) * m
    return colptr, rowidx, val, mu.numpy(), start.numpy()


def gen_sample_cpu(n, p, m, K, seed=2024):
    """Same distribution on the host (numpy), in the reference's format (double, 64-bit idx)."""
    rng0 = np.random.default_rng(1234)
    mu = rng0.standard_normal((p, K))
    start = mu + 0.05 * np.random.default_rng(99).standard_normal((p, K))
    rng = np.random.default_rng(seed)
    ir = np.empty(n * m, dtype=np.uint64)
    x = np.empty(n * m, dtype=np.float64)
    scale = p / m
    step = 100_000
    for c0 in range(0, n, step):
        c = min(step, n - c0)
        rows = np.argpartition(rng.random((c, p), dtype=np.float32), m - 1, axis=1)[:, :m]
        rows.sort(axis=1)
        lab = (np.arange(c0, c0 + c) % K)
        v = (mu[rows, lab[:, None]] + 0.1 * rng.standard_normal((c, m))) * scale
        ir[c0 * m:(c0 + c) * m] = rows.reshape(-1)
        x[c0 * m:(c0 + c) * m] = v.astype(np.float32).reshape(-1)
    jc = (np.arange(n + 1, dtype=np.uint64) * np.uint64(m))
    return jc, ir, x, mu, start


# --------------------------------------------------------------------------- CPU arm
class CpuLloyd:
    """One Lloyd iteration with the reference's own compiled C kernel (oracle/_ref, built from
    /root/reference/private/SparseMatrixMinusCluster.c) for the distances, followed by the
    oracle port of MATLAB's min and of the centre update, on disjoint column slices across
    host threads (the reference itself is single-threaded; this is 'all the cores it can use')."""

    def __init__(self, p, K, jc, ir, x, threads):
        from oracle import cport, refmex
        self.cport, self.refmex = cport, refmex
        self.kind = "reference" if refmex.ref_available() else "port"
        self.p, self.K = p, K
        self.n = jc.shape[0] - 1
        self.threads = max(1, int(threads))
        self.slices = []
        for w in range(self.threads):
            a, b = self.n * w // self.threads, self.n * (w + 1) // self.threads
            if b <= a:
                continue
            j = jc[a:b + 1] - jc[a]
            lo, hi = int(jc[a]), int(jc[b])
            self.slices.append((b - a, np.ascontiguousarray(j), ir[lo:hi], x[lo:hi],
                                np.ascontiguousarray(j.astype(np.int64)), ir[lo:hi].view(np.int64)))
        cport.lib()
        if self.kind == "reference":
            refmex.load_ref("SparseMatrixMinusCluster")

    def _work(self, sl, cscaled, gamma, centers):
        nn, jcu, iru, xs, jci, iri = sl
        if self.kind == "reference":
            D = self.refmex.SparseMatrixMinusCluster(self.p, nn, jcu, iru, xs, cscaled)
            dmin, a = self.cport.colmin(D)
        else:
            a, dmin = self.cport.assign(self.p, nn, jci, iri, xs, cscaled)
        _, S, N, counts = self.cport.centroid_update(self.p, nn, self.K, jci, iri, xs, a, gamma, centers, True)
        return S, N, counts, float(np.sum(dmin * dmin))

    def iterate(self, centers, gamma):
        from concurrent.futures import ThreadPoolExecutor
        cscaled = centers / gamma
        with ThreadPoolExecutor(self.threads) as ex:
            parts = list(ex.map(lambda sl: self._work(sl, cscaled, gamma, centers), self.slices))
        S = sum(pt[0] for pt in parts)
        N = sum(pt[1] for pt in parts)
        counts = sum(pt[2] for pt in parts)
        new = centers.copy()
        ok = counts > 0
        new[:, ok] = gamma * S[:, ok] / (N[:, ok] + 1e-16)
        return new


def cpu_rate(cfg, budget_s, threads, probe_n=None):
    """(points/s, sample description, kind, threads) of the CPU arm on a bounded sample."""
    p, K, m = cfg["p"], cfg["K"], cfg["m"]
    n_cpu = probe_n or min(cfg["n"], 4_000_000, max(200_000, 40_000 * threads))
    jc, ir, x, mu, start = gen_sample_cpu(n_cpu, p, m, K)
    eng = CpuLloyd(p, K, jc, ir, x, threads)
    gamma = m / p
    c = eng.iterate(start, gamma)                       # warm-up pass (page faults, thread start)
    t0 = time.perf_counter()
    passes = 0
    while True:
        c = eng.iterate(c, gamma)
        passes += 1
        el = time.perf_counter() - t0
        if el >= budget_s or passes >= 5000:
            break
    rate = passes * n_cpu / el
    return rate, f"{passes} Lloyd iterations over the first {n_cpu} columns of the workload ({el:.1f} s)", eng.kind, eng.threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    p, K, m = cfg["p"], cfg["K"], cfg["m"]
    n_cpu = min(cfg["n"], 4_000_000, max(200_000, 40_000 * threads))
    jc, ir, x, mu, start = gen_sample_cpu(n_cpu, p, m, K)
    eng = CpuLloyd(p, K, jc, ir, x, threads)
    gamma = m / p
    c = start
    for _ in range(args.warmup):
        c = eng.iterate(c, gamma)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c = eng.iterate(c, gamma)
    el = time.perf_counter() - t0
    value = args.steps * n_cpu / el
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": cfg["label"], "sample_columns": n_cpu, "p": p, "k": K, "nnz_per_col": m},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": eng.threads, "kind": eng.kind,
                         "sample": f"each step = one Lloyd iteration over {n_cpu} columns of the workload "
                                   f"on {eng.threads} host threads (reference kernel is single-threaded; "
                                   f"columns sliced across threads)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows]
        for r in rows:
            f = [s.strip() for s in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sparsifiedkmeans_b200 import Context, Dataset, Lloyd
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"           # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    n, p, K, m = (args.n or cfg["n"]), cfg["p"], cfg["K"], cfg["m"]
    gamma = m / p

    ctx = Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    colptr, rowidx, val, mu, start = gen_shard_device(dev, n, p, m, K, col0=rank * n)
    torch.cuda.synchronize()

    # pinned host copy for the end-to-end leg (made before the device tensors are released)
    # columns per end-to-end step: everything on one GPU; under torchrun a bounded sample per rank, so that the
    # pinned host copies of all ranks together stay well inside the box's memory (throughput does not depend on it)
    n_e2e = min(n, args.e2e_n) if args.e2e_n else (n if world == 1 else min(n, 4_000_000))
    h_colptr = torch.empty(n_e2e + 1, dtype=torch.int64, pin_memory=True)
    h_rowidx = torch.empty(n_e2e * m, dtype=torch.int32, pin_memory=True)
    h_val = torch.empty(n_e2e * m, dtype=torch.float32, pin_memory=True)
    h_colptr.copy_(colptr[:n_e2e + 1]); h_rowidx.copy_(rowidx[:n_e2e * m]); h_val.copy_(val[:n_e2e * m])
    torch.cuda.synchronize()

    ds = Dataset.from_device_csc(p, n, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32,
                                 val.data_ptr(), SKM_F32, store="f32", ctx=ctx)
    del colptr, rowidx, val
    torch.cuda.empty_cache()
    L = Lloyd(ds, K)
    L.set_centers(start)

    def reduce_partials(Lo):
        if world > 1:
            with torch.cuda.stream(ext):
                dist.all_reduce(Lo.partials_tensor())

    def step():
        return L.step(gamma, gamma, True, reduce=reduce_partials)

    for _ in range(max(args.warmup, 3)):
        st = step()
    ctx.synchronize()
    ctx.timing_enable(True)
    ctx.timing_read()

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    fence()
    t_w0 = time.perf_counter()
    ev0.record(ext)
    for _ in range(args.steps):
        st = step()
    ev1.record(ext)
    fence()
    t_w1 = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    tim = ctx.timing_read()
    ctx.timing_enable(False)
    clocks = sampler.stop(t_w0, t_w1) if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (K1: fused distance + argmin) ----
    peak, peak_src = _peaks()
    # live device-to-device copy rate on this box, for orientation only (read + write bytes / time); the
    # roofline denominator stays MEASURED_PEAKS.json (or the recipe's fallback)
    copy_gbs = None
    try:
        a = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        b_ = torch.empty_like(a)
        b_.copy_(a)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            b_.copy_(a)
        c1.record()
        torch.cuda.synchronize()
        copy_gbs = 5 * 2 * a.numel() / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del a, b_
    except Exception:
        pass
    k1_ms, k1_groups = tim["assign"]
    k1_ms_avg = k1_ms / max(k1_groups, 1)
    alg_bytes = n * (m * 8 + 8)                         # m*(4+4) + 4 + 4 per point
    achieved = alg_bytes / (k1_ms_avg * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu --set full capture (per-point figure x points)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f)[L.kernel_name]["dram_bytes_per_point"] * n      # keyed by the kernel the plan runs
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": L.kernel_name, "ms_per_launch": k1_ms_avg,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "d2d_copy_gbs_this_run": copy_gbs,
                "step_breakdown_ms": {k: v[0] / args.steps for k, v in tim.items() if v[1]}}

    # ---- same iteration with the opt-in modes (not the headline: the synthetic mixture is at its fixed point
    # here, so no column changes cluster -- K2 reduces to the comparison pass and every bound holds; reported
    # so the modes can be told apart from the recompute-everything iteration above) ----
    def timed_mode(incr, bounded):
        try:
            L.set_update_mode(incr)
            L.set_assign_mode(bounded)
            for _ in range(3):
                step()
            fence()
            ctx.timing_enable(True); ctx.timing_read()
            ev0.record(ext)
            for _ in range(args.steps):
                sti = step()
            ev1.record(ext)
            fence()
            tm = ctx.timing_read(); ctx.timing_enable(False)
            ti = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ti, op=dist.ReduceOp.MAX)
            ms_i = float(ti.item()) / args.steps
            kind, nch = L.last_update()
            k1i = tm["assign"][0] / args.steps
            return {"ms_per_step": ms_i, "value": world * n / (ms_i * 1e-3), "unit": UNIT, "last_update": kind,
                    "columns_moved_last_step": nch, "columns_reevaluated_last_step": L.last_assign_flagged(),
                    "assign_ms": k1i, "assign_algorithmic_GBps": alg_bytes / (k1i * 1e-3) / 1e9 if k1i else None,
                    "objective": sti.objective}
        except Exception as e:                              # never let the extra legs break the contract line
            return {"error": str(e)}
        finally:
            L.set_update_mode(False)
            L.set_assign_mode(False)
    incremental = timed_mode(True, False)
    incremental["note"] = "opt-in skm_lloyd_set_update_mode(1); the synthetic mixture is at its fixed point here"
    bounded_leg = timed_mode(True, True)
    bounded_leg["note"] = "opt-in skm_lloyd_set_assign_mode(1) + set_update_mode(1), same fixed point"

    stream_gb = ds.stream_bytes / 1e9
    # ---- end to end: host buffers in, host results out, every step ----
    e2e = None
    if args.e2e_steps > 0:
        hj, hi, hv = h_colptr.numpy(), h_rowidx.numpy(), h_val.numpy()
        from sparsifiedkmeans_b200 import lloyd_step_host
        # the same host matrix with 2-byte row indices (p <= 65536): the most compact format the ABI accepts
        h_row16 = torch.empty(n_e2e * m, dtype=torch.uint16, pin_memory=True) if p <= 65536 else None
        if h_row16 is not None:
            h_row16.copy_(h_rowidx.to(torch.uint16))
        red = (lambda t: dist.all_reduce(t)) if world > 1 else None

        def measure(rows_np, label):
            def e2e_step():
                # stateless call: X in pinned host memory, streamed over PCIe in column chunks
                newc, a, _, st2 = lloyd_step_host(p, n_e2e, hj, rows_np, hv, start, gamma, gamma, True,
                                                  want_assign=True, want_dist=False, ctx=ctx, reduce=red)
                return a, newc
            e2e_step()
            fence()
            ev0.record(ext)
            for _ in range(args.e2e_steps):
                e2e_step()
            ev1.record(ext)
            fence()
            tt = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item()) / args.e2e_steps
            h2d = hj.nbytes + rows_np.nbytes + hv.nbytes + start.nbytes
            d2h = n_e2e * 4 + start.nbytes
            return {"value": world * n_e2e / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e_ms, "columns_per_step": n_e2e,
                    "what": "lloyd_step_host: X in pinned host memory (int64 colptr, %s rows, fp32 values) streamed over "
                            "PCIe in chunks + K1/K2/K3 + assignments and centres read back, every step" % label}
        e2e_i32 = measure(hi, "int32")
        if h_row16 is not None:
            e2e = measure(h_row16.numpy(), "uint16")
            e2e["int32_rows_variant"] = {k: e2e_i32[k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")}
        else:
            e2e = e2e_i32

        # Whole-job variant (reported beside the per-step number above, never instead of it): what one
        # kmeans_sparsified call does with the handle API of INTEGRATION.md level 2 -- upload X from pinned host
        # memory ONCE, build the device images, run `steps` Lloyd iterations (centres in, statistics out, every
        # iteration), read assignments, distances and centres back.  Everything is inside the timed region
        # (host wall clock around synchronous calls); throughput = columns x iterations / time.
        try:
            rows_np = h_row16.numpy() if h_row16 is not None else hi
            def whole_job():
                ds2 = Dataset.from_csc(p, n_e2e, hj, rows_np, hv, store="f32", ctx=ctx)
                L2 = Lloyd(ds2, K)
                L2.set_centers(start)
                for _ in range(args.steps):
                    L2.step(gamma, gamma, True)
                a2, d2 = L2.assignments()
                c2 = L2.get_centers()
                L2.close(); ds2.close()
                return a2, c2
            if world == 1:
                L.close(); ds.close()                          # make room: the job builds its own images
                ds = L = None
                whole_job()
                torch.cuda.synchronize()
                tj0 = time.perf_counter()
                whole_job()
                torch.cuda.synchronize()
                tj = time.perf_counter() - tj0
                e2e["whole_job_variant"] = {
                    "value": n_e2e * args.steps / tj, "unit": UNIT, "seconds": tj, "iterations": args.steps,
                    "columns": n_e2e, "h2d_bytes_once": int(hj.nbytes + rows_np.nbytes + hv.nbytes),
                    "what": "upload from pinned host memory + device image build + iterations + read-back, all timed"}
        except Exception as ex:
            e2e["whole_job_variant"] = {"error": str(ex)}

    if rank == 0:
        cpu = None
        if not args.no_cpu:
            r, sample, kind, cores = cpu_rate(cfg, args.cpu_seconds, os.cpu_count() or 1)
            cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["label"], "n_per_gpu": n, "p": p, "k": K, "nnz_per_col": m,
                       "l2": "inputs (%.1f GB streamed per iteration) far exceed the 126 MB L2; no flush needed"
                             % stream_gb,
                       "update": "per-cluster sums recomputed from all columns every iteration (reference semantics)",
                       "parallelism": f"columns sharded over {world} GPU(s), one all-reduce of per-cluster partials per iteration"
                                      if world > 1 else "single GPU"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "incremental_update": incremental, "bounded_assign": bounded_leg,
            "rechecked_last_step": st.n_rechecked, "objective": st.objective,
        }
        print(json.dumps(out))
    if L is not None:
        L.close(); ds.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="override points per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-n", type=int, default=0, help="columns per end-to-end step (default: all)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
