"""Probe of the streamed (host-resident) iteration under torchrun: time with / without the
all-reduce callback, to separate PCIe contention from collective cost."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from sparsifiedkmeans_b200 import Context, lloyd_step_host
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
if world > 1:
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
n, p, K, m = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 784, 10, 78
ctx = Context(local)
colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, p, m, K, col0=rank * n)
hj = torch.empty(n + 1, dtype=torch.int64, pin_memory=True); hi = torch.empty(n * m, dtype=torch.int32, pin_memory=True); hv = torch.empty(n * m, dtype=torch.float32, pin_memory=True)
hj.copy_(colptr); hi.copy_(rowidx); hv.copy_(val); torch.cuda.synchronize()
del colptr, rowidx, val
gamma = m / p
def run(red, tag):
    for it in range(3):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lloyd_step_host(p, n, hj.numpy(), hi.numpy(), hv.numpy(), start, gamma, gamma, True, want_assign=True, ctx=ctx, reduce=red)
        t1 = time.perf_counter()
        print(f"rank {rank} {tag} iter {it}: {1e3*(t1-t0):.1f} ms  ({(hi.numel()*8)/(t1-t0)/1e9:.1f} GB/s H2D equiv)", flush=True)
run(None, "no-reduce")
if world > 1:
    run(lambda t: dist.all_reduce(t), "all-reduce")
    dist.destroy_process_group()
