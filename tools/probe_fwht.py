"""In-place fp32 FWHT at several column lengths: ms, GB/s, fraction of the measured HBM peak (kernel variant via env)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsifiedkmeans_b200 import Context, fwht_f32_inplace
from tools.bench_stages import peak, timed

dev = torch.device("cuda:0")
ctx = Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
pk = peak()
for p2 in (2048, 4096, 8192, 16384, 32768):
    n = min(1_000_000, (8 << 30) // (4 * p2))
    x = torch.randn(n, p2, device=dev)
    ref = None
    if n * p2 <= (1 << 28):
        ref = x[:64].clone()
    signs = torch.sign(torch.randn(p2, device=dev)); signs[signs == 0] = 1
    torch.cuda.synchronize()
    ms = timed(ctx, ext, lambda: fwht_f32_inplace(p2, n, x.data_ptr(), signs.data_ptr(), ctx), 5)
    gb = 2 * 4 * p2 * n / 1e9
    print(json.dumps({"variant_env": {k: os.environ.get(k) for k in ("SKM_FWHT_X", "SKM_FWHT_NO_TMA")}, "p2": p2, "n": n, "ms": ms,
                      "GBps": gb / ms * 1e3, "frac_of_hbm_peak": gb / ms * 1e3 / pk}), flush=True)
    del x
    torch.cuda.empty_cache()
