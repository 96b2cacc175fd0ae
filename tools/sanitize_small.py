"""Small softmap / linear / kNN calls for compute-sanitizer runs: python tools/sanitize_small.py"""
import sys
import torch
sys.path.insert(0, ".")
from dv_matcher_b200 import ops
torch.manual_seed(0)
x = torch.randn(2, 300, 64).cuda(); y = torch.randn(2, 700, 64).cuda(); v = torch.randn(2, 700, 3).cuda()
for prec in ("f16", "bf16"):
    o = ops.softmap_fwd(x, y, v, alpha=20.0, topk=10, soft=True, prec=prec)
    h = ops.softmap_fwd(x, y, None, topk=1, soft=False, prec=prec)
xb = torch.randn(1, 600, 128).cuda(); yb = torch.randn(1, 5000, 128).cuda()
o2 = ops.softmap_fwd(xb, yb, None, alpha=50.0, topk=10, soft=True, prec="f16")      # primed (>= 16 tiles)
w = torch.randn(70, 30).cuda(); b = torch.randn(70).cuda()
z = ops.linear_act_fwd(torch.randn(260, 30).cuda(), w, b, "elu")
a = torch.randn(2, 3000, 3).cuda(); c = torch.randn(2, 2500, 3).cuda()
ops.knn3(a, a, 10); ops.chamfer_fwd(a, c)
torch.cuda.synchronize()
print("ok", o.argmin[0, :3].tolist(), float(z[0, 0]))
