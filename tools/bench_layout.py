#!/usr/bin/env python
"""One-off cost of the streamed-image layouts (dataset build and re-ordering for the dual-table
kernels) at the config-2 shape.  Wall clock around synchronous library calls.

    python tools/bench_layout.py [--n 2000000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    args = ap.parse_args()
    import torch
    import bench
    from sparsifiedkmeans_b200 import Context, Dataset
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
    dev = torch.device("cuda:0")
    ctx = Context(0)
    n, p, m, K = args.n, 784, 78, 10
    colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, p, m, K, col0=0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds = Dataset.from_device_csc(p, n, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32,
                                 store="f32", ctx=ctx)
    ctx.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3

    def timed(layout):
        ctx.synchronize()
        t = time.perf_counter()
        r = ds.layout_check(layout)
        return (time.perf_counter() - t) * 1e3, r
    check_ms, r0 = timed(-1)
    out = {"n": n, "p": p, "m": m, "dataset_build_ms (CSC copy + validate + layout 0 + CSR)": build_ms, "check_only_ms": check_ms,
           "layout0": r0}
    for layout in (1, 2, 0, 1):
        ms, r = timed(layout)
        out.setdefault("relayout_ms", []).append({"to": layout, "ms": ms - check_ms, "wavefronts_per_step": r["wavefronts"] / max(r["steps"], 1),
                                                  "bad_columns": r["bad_columns"]})
    print(json.dumps(out))
    ds.close()


if __name__ == "__main__":
    main()
