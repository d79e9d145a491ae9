"""In-place fp32 FWHT: correctness against a torch butterfly on a few columns + ms / fraction of the HBM peak."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsifiedkmeans_b200 import Context, fwht_f32_inplace
from tools.bench_stages import peak, timed

def torch_fwht(x):                      # x: (cols, p2) float64
    c, p2 = x.shape
    y = x.clone(); h = 1
    while h < p2:
        y = y.view(c, p2 // (2 * h), 2, h)
        y = torch.stack((y[:, :, 0] + y[:, :, 1], y[:, :, 0] - y[:, :, 1]), dim=2).reshape(c, p2)
        h *= 2
    return y / math.sqrt(p2)

dev = torch.device("cuda:0"); ctx = Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=dev); pk = peak()
for p2 in [int(a) for a in (sys.argv[1:] or (2048, 4096))]:
    n = min(1_000_000, (8 << 30) // (4 * p2))
    x = torch.randn(n, p2, device=dev)
    signs = torch.sign(torch.randn(p2, device=dev)); signs[signs == 0] = 1
    idx = torch.tensor([0, 1, 2, n // 2, n - 2, n - 1], device=dev)
    want = torch_fwht((x[idx].double() * signs.double()))
    torch.cuda.synchronize()
    fwht_f32_inplace(p2, n, x.data_ptr(), signs.data_ptr(), ctx); ctx.synchronize()
    err = float((x[idx].double() - want).abs().max() / want.abs().max())
    ms = timed(ctx, ext, lambda: fwht_f32_inplace(p2, n, x.data_ptr(), signs.data_ptr(), ctx), 5)
    gb = 2 * 4 * p2 * n / 1e9
    print(json.dumps({"minw": os.environ.get("SKM_FWHT_TMA_MINW"), "p2": p2, "n": n, "rel_err": err, "ms": round(ms, 4),
                      "frac_of_hbm_peak": round(gb / ms * 1e3 / pk, 4)}), flush=True)
    del x; torch.cuda.empty_cache()
