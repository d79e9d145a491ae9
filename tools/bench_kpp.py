#!/usr/bin/env python
"""BASELINE.json configs[4]: k-means++ (Arthur_initialization) on the sparsified matrix, n=1e7, p=784,
K=100, columns sharded over the GPUs of one box.

    python tools/bench_kpp.py                                                    # one GPU, whole matrix
    python -m torch.distributed.run --nproc-per-node 8 tools/bench_kpp.py        # 8 GPUs

Each of the K-1 rounds folds the masked distance to the newest centre into the running minimum on
every shard (one pass over the shard), all-gathers the local D^2 sums, and the owning rank finds the
sampled column by a prefix search and broadcasts it.  The reference recomputes the distance to ALL chosen
centres in every round (private/Arthur_initialization.m:39), K(K-1)/2 = 4950 centre-passes.
Prints one JSON line on rank 0.
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
    ap.add_argument("--n", type=int, default=10_000_000, help="total points")
    ap.add_argument("--k", type=int, default=100)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from sparsifiedkmeans_b200 import Context, Dataset
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
    from sparsifiedkmeans_b200.distributed import CudaShardEngine, shard_bounds, sharded_arthur_initialization
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    p, m, K = 784, 78, args.k
    lo, hi = shard_bounds(args.n, world, rank)
    ctx = Context(local)
    colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, hi - lo, p, m, 10, col0=lo)
    torch.cuda.synchronize()                       # the library reads these on its own stream
    ds = Dataset.from_device_csc(p, hi - lo, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32,
                                 store="f32", ctx=ctx)
    del colptr, rowidx, val
    torch.cuda.empty_cache()
    eng = CudaShardEngine(ds, 4)
    gamma = m / p
    uni = np.random.default_rng(7).random(100_000)
    sharded_arthur_initialization(eng, 3, gamma, args.n, lo, first=5, uniforms=uni)        # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, cols = sharded_arthur_initialization(eng, K, gamma, args.n, lo, first=5, uniforms=uni)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    el = time.perf_counter() - t0
    t = torch.tensor([el], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    el = float(t.item())
    if rank == 0:
        print(json.dumps({"stage": "k-means++ (Arthur_initialization), BASELINE configs[4]", "n": args.n, "p": p, "k": K,
                          "nnz_per_col": m, "n_gpus": world, "seconds": el, "ms_per_round": 1e3 * el / (K - 1),
                          "point_centre_distances_per_s": args.n * (K - 1) / el,
                          "distinct_centres": int(len(set(idx.tolist()))),
                          "reference_centre_passes": K * (K - 1) // 2, "centre_passes_here": K - 1}))
    eng.close(); ds.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
