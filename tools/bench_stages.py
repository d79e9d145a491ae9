#!/usr/bin/env python
"""Stage benchmarks beside bench.py's headline: K4 (FWHT, FWHT+sample) over p (BASELINE.json
configs[3]), K5 (k-means++ rounds, configs[4] shape), and dataset upload.  CUDA events on the
library stream; CPU numbers are the reference's own hadamard.c / hadamard_pthreads.c
(oracle/_ref) on a bounded sample.  Prints one JSON object per stage.

    python tools/bench_stages.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(ctx, ext, fn, reps):
    import torch
    fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(reps):
        fn()
    e1.record(ext)
    ctx.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-dense", action="store_true", help="only the dense second-pass stage")
    args = ap.parse_args()
    import torch
    from sparsifiedkmeans_b200 import Context, Dataset, fwht_f32_inplace
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
    dev = torch.device("cuda:0")
    ctx = Context(0)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    pk = peak()
    budget = (1 if args.quick else 8) * (1 << 30)          # bytes of dense input per measurement

    # ---- K4: FWHT in place and FWHT + fixed-count row sample, gamma = 0.05 ----
    for p2 in (() if args.only_dense else (512, 4096, 32768)):
        n = min(1_000_000, budget // (4 * p2))
        x = torch.randn(n, p2, device=dev)                  # column-major p2 x n
        signs = torch.sign(torch.randn(p2, device=dev))
        signs[signs == 0] = 1
        torch.cuda.synchronize()
        ms = timed(ctx, ext, lambda: fwht_f32_inplace(p2, n, x.data_ptr(), signs.data_ptr(), ctx), 5)
        gb = 2 * 4 * p2 * n / 1e9
        out = {"stage": "K4 fwht_f32_inplace", "p2": p2, "n": n, "ms": ms, "GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / pk,
               "algorithmic_bytes": "read 4*p2 + write 4*p2 per column"}
        print(json.dumps(out), flush=True)
        m = max(1, int(round(0.05 * p2)))
        keys = torch.rand(min(n, 200_000), p2, device=dev)
        rows1 = keys.topk(m, dim=1, largest=False).indices.to(torch.int32)
        del keys
        rows = rows1.repeat((n + rows1.shape[0] - 1) // rows1.shape[0], 1)[:n].contiguous()
        x.normal_()
        torch.cuda.synchronize()                            # torch's stream and the library's are independent
        holder = {}

        def fused():
            if "ds" in holder:
                holder["ds"].close()
            holder["ds"] = Dataset.from_fwht_sample(p2, n, m, x.data_ptr(), signs.data_ptr(), holder.get("rows"), ctx=ctx, seed=3)
        holder["rows"] = rows.data_ptr()
        t0 = time.perf_counter()
        fused()
        ctx.synchronize()
        t_all = (time.perf_counter() - t0) * 1e3
        ctx.timing_enable(True); ctx.timing_read()
        fused(); ctx.synchronize()
        k_ms = ctx.timing_read()["fwht"][0]
        ctx.timing_enable(False)
        gb = (4 * p2 + 8 * m + 4 * m) * n / 1e9              # dense read + (row,val) write + sampled-row list read
        out = {"stage": "K4 fwht_sample (incl. building the resident images)", "p2": p2, "n": n, "m": m, "ms_total_call": t_all,
               "kernel_ms": k_ms, "kernel_GBps": gb / k_ms * 1e3 if k_ms else None, "kernel_frac_of_hbm_peak": gb / k_ms * 1e3 / pk if k_ms else None,
               "algorithmic_bytes": "4*p2 read + 12*m (row list in, (row,val) out) per column"}
        print(json.dumps(out), flush=True)
        holder["rows"] = None                                  # rows drawn on the device (Philox)
        ctx.timing_enable(True); ctx.timing_read()
        fused(); ctx.synchronize()
        k2 = ctx.timing_read()["fwht"][0]
        ctx.timing_enable(False)
        gb2 = (4 * p2 + 8 * m) * n / 1e9
        print(json.dumps({"stage": "K4 fwht_sample with on-device row sampling", "p2": p2, "n": n, "m": m, "kernel_ms": k2,
                          "kernel_GBps": gb2 / k2 * 1e3 if k2 else None}), flush=True)
        holder["ds"].close()
        del x, rows, rows1
        torch.cuda.empty_cache()
        if not args.no_cpu:
            from oracle import refmex
            ncpu = max(1, min(n, (256 << 20) // (8 * p2)))
            xc = np.random.default_rng(0).standard_normal((p2, ncpu))
            t0 = time.perf_counter(); refmex.hadamard(xc); t_ser = time.perf_counter() - t0
            nthr = max([v for v in refmex.pthreads_variants() if v <= (os.cpu_count() or 1)] or [4])
            t0 = time.perf_counter(); refmex.hadamard_pthreads(xc, nthr); t_par = time.perf_counter() - t0
            gbc = 2 * 8 * p2 * ncpu / 1e9
            print(json.dumps({"stage": "CPU reference hadamard.c / hadamard_pthreads.c (fp64)", "p2": p2, "n": ncpu,
                              "serial_GBps": gbc / t_ser, "pthreads_GBps": gbc / t_par, "NTHREADS": nthr,
                              "host_cores": os.cpu_count()}), flush=True)

    # ---- K5: k-means++ rounds on configs[4]'s shape (n per GPU = 1e7/8, p=784, m=78) ----
    sys.path.insert(0, ROOT)
    import bench
    n = 250_000 if args.quick else 1_250_000
    if args.only_dense:
        n = 1000
    colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, 784, 78, 100, col0=0, ctx=ctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds = Dataset.from_device_csc(784, n, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32,
                                 store="f32", ctx=ctx)
    ctx.synchronize()
    t_up = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"stage": "dataset build from device CSC (SELL banked + CSR)", "n": n, "nnz": n * 78, "ms": t_up,
                      "M_entries_per_s": n * 78 / t_up / 1e3}), flush=True)
    c = ds.get_column(0)
    ms = timed(ctx, ext, lambda: ds.kpp_update(c, 78 / 784, first=False), 5)
    print(json.dumps({"stage": "K5 kpp_update (one k-means++ round incl. D^2 block sums + host read-back)", "n": n, "ms": ms,
                      "points_per_s": n / ms * 1e3, "GBps": n * (78 * 8 + 16) / 1e9 / ms * 1e3,
                      "reference_cost": "round k recomputes k centre-passes (Arthur_initialization.m:39); here 1 per round"}), flush=True)
    ds.close()
    del colptr, rowidx, val
    torch.cuda.empty_cache()

    # ---- second pass over the original dense data, X resident on the device ----
    from sparsifiedkmeans_b200 import second_pass
    for (p, K) in ((784, 10), (1024, 64)):
        n = min(2_000_000, budget // (4 * p))
        g = torch.Generator(device=dev).manual_seed(3)
        mu = torch.randn(K, p, generator=g, device=dev)
        lab = torch.arange(n, device=dev) % K
        x = mu[lab] + 0.3 * torch.randn(n, p, generator=g, device=dev)         # (n, p) row-major = p x n column-major
        cen = (mu + 0.05 * torch.randn(K, p, generator=g, device=dev)).T.double().cpu().numpy()
        lab1 = (lab + 1).int().cpu().numpy()
        torch.cuda.synchronize()
        holder = {}

        def run(sums, assign):
            holder["r"] = second_pass(None, centers=cen if assign else None, assign_in=lab1 if sums else None,
                                      want_assign=assign, want_dist=assign, ctx=ctx, x_device_ptr=x.data_ptr(),
                                      shape=(p, n), x_dtype=np.float32)
        for name, sums, assign in (("dense per-cluster means (k_dense_sums)", True, False),
                                   ("dense distance + argmin (k_dense_assign)", False, True)):
            ctx.timing_enable(True); ctx.timing_read()
            run(sums, assign); run(sums, assign)
            tim = ctx.timing_read(); ctx.timing_enable(False)
            k_ms = tim["assign"][0] / 2
            out = {"stage": "second pass: " + name, "p": p, "n": n, "K": K, "kernel_ms": k_ms,
                   "kernel_GBps": 4 * p * n / 1e9 / k_ms * 1e3, "frac_of_hbm_peak": 4 * p * n / 1e9 / k_ms * 1e3 / pk}
            if assign:
                out["rechecked"] = holder["r"]["n_rechecked"]
                out["accuracy_vs_labels"] = float(np.mean(holder["r"]["assign"] == lab1))
                out["fp32_Tinstr_per_s"] = 2 * K * p * n / 1e12 / k_ms * 1e3
            print(json.dumps(out), flush=True)
        del x, mu, lab
        torch.cuda.empty_cache()

    # ---- DCT sketch: dense host matrix -> 3xTF32 product on the tensor cores -> row sample ----
    for p in (784, 1000):
        n = 200_000 if args.quick or args.only_dense else 1_000_000
        Xh = torch.randn(n, p, dtype=torch.float32).pin_memory().numpy().T        # (p, n) view of pinned memory
        d = np.sign(np.random.default_rng(0).standard_normal(p)); d[d == 0] = 1
        m = max(1, round(0.1 * p))
        t0 = time.perf_counter()
        ds = Dataset.from_dense_host_dct(Xh, d, m, seed=5, ctx=ctx)
        ctx.synchronize()
        ds.close()
        ctx.timing_enable(True); ctx.timing_read()
        t0 = time.perf_counter()
        ds = Dataset.from_dense_host_dct(Xh, d, m, seed=5, ctx=ctx)
        ctx.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        k_ms = ctx.timing_read()["fwht"][0]; ctx.timing_enable(False)
        ds.close()
        flops = 3 * 2.0 * p * p * n
        print(json.dumps({"stage": "DCT sketch: split + 3xTF32 tcgen05 product + row sample (from pinned host memory)", "p": p, "n": n,
                          "m": m, "wall_ms": wall, "device_ms_split_gemm_sample": k_ms, "tf32_TFLOPs_incl_split_and_sample": flops / k_ms / 1e9,
                          "pcie_GBps": 4.0 * p * n / 1e9 / wall * 1e3}), flush=True)
        del Xh


if __name__ == "__main__":
    main()
