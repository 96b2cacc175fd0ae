import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
torch.manual_seed(0)
def run(B, N, M, C, alpha, structured=False):
    if structured:
        d = synthetic.make_batch(B, N, M)
        x, y = d["feat1"][..., :C].contiguous(), d["feat2"][..., :C].contiguous()
    else:
        x = torch.randn(B, N, C) * 0.3; y = torch.randn(B, M, C) * 0.3
        y[:, 3] = x[:, 5]
    coef = torch.randn(B, N, 10)
    xg, yg = x.cuda(), y.cuda()
    out = ops.softmap_fwd(xg, yg, None, alpha=alpha, prec="fp32")
    dw = coef.cuda()
    dx32, dy32 = ops.softmap_bwd(xg, yg, alpha, out, dw, prec="fp32")
    dxt, dyt = ops.softmap_bwd(xg, yg, alpha, out, dw, prec="f16")
    torch.cuda.synchronize()
    def rel(a, b): return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    def cos(a, b): return float((a * b).sum() / (a.norm() * b.norm()))
    print(f"B{B} N{N} M{M} C{C} a{alpha} struct={structured}: dX rel {rel(dxt, dx32):.2e} cos {cos(dxt, dx32):.6f} | dY rel {rel(dyt, dy32):.2e} cos {cos(dyt, dy32):.6f} | nan {bool(torch.isnan(dxt).any() or torch.isnan(dyt).any())}", flush=True)
    return xg, yg, out, dw
run(1, 128, 128, 128, 5.0)
run(1, 256, 384, 128, 5.0)
run(2, 300, 257, 64, 20.0)
run(1, 130, 500, 128, 60.0)
run(2, 1000, 1030, 128, 10.0)
xg, yg, out, dw = run(2, 4995, 4995, 128, 100.0, structured=True)
run(2, 4995, 4995, 128, 10.0, structured=True)
import time
for prec in ("fp32", "f16"):
    for _ in range(2): ops.softmap_bwd(xg, yg, 100.0, out, dw, prec=prec)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): ops.softmap_bwd(xg, yg, 100.0, out, dw, prec=prec)
    torch.cuda.synchronize(); print(prec, "bwd ms", (time.perf_counter() - t) / 5 * 1e3)
