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
# ---- round 2 kernels
from dv_matcher_b200 import deformation_graph as dg, geometry, synthetic
o3 = ops.softmap_fwd(xb, yb, None, alpha=10.0, topk=10, soft=True, prec="f16")      # dense-window instance (alpha < 40)
# a row with 40 exact ties at distance 0: list overflow -> certificate fails -> rescue scan / exact-row cluster kernel
yt = yb.clone(); yt[:, :40] = xb[:, 5:6]
o4 = ops.softmap_fwd(xb, yt, None, alpha=50.0, topk=10, soft=True, prec="f16", want_stats=True)
dw = torch.randn(1, 600, 10).cuda()
gx, gy = ops.softmap_bwd(xb, yb, 50.0, o2, dw, prec="f16")                          # tcgen05 backward
gx32, gy32 = ops.softmap_bwd(xb, yb, 50.0, o2, dw, prec="fp32")
sc = torch.randn(300, 2000).cuda()
ti = ops.topk_select(sc, 500)
feat = torch.randn(1, 700, 128).cuda()
qi = torch.randint(0, 700, (1, 50)).cuda(); nb = torch.randint(0, 700, (1, 50, 20)).cuda()
pd = ops.pair_dist_fwd(feat, qi, nb)
if isinstance(pd, tuple): pd = pd[0]
ops.pair_dist_bwd(feat, qi, nb, pd, torch.ones_like(pd))
pts = torch.randn(2, 900, 3).cuda(); i2 = torch.randint(0, 900, (2, 100, 7)).cuda()
gr = ops.gather_rows_fwd(pts, i2); ops.gather_rows_bwd(torch.ones_like(gr), i2, 900)
d = synthetic.make_batch(2, 1500, 1500)
verts = d["xyz1"].cuda()
graph = dg.build_graphs(verts, torch.zeros(2, dtype=torch.int64).cuda())
if graph is not None:
    d9 = torch.randn(2, graph.pack.nodes_xyz.shape[1], 9).cuda() * 0.05
    out = dg.deform_from_d9(verts, graph, d9, want_sr=True)
torch.cuda.synchronize()
print("ok2", o4.stats.cpu().tolist(), float(gx.abs().max()), float((gx - gx32).abs().max()))
# ---- secondary API, LG-Net attention
from dv_matcher_b200 import secondary, lgnet
fs = torch.randn(1, 420, 32).cuda(); ft = torch.randn(1, 360, 32).cuda()
secondary.forward_source_target(fs, ft, torch.randn(1, 420, 3).cuda(), torch.randn(1, 360, 3).cuda())
secondary.cross_construct(fs, ft, torch.randn(1, 360, 3).cuda(), 10)
qq = torch.randn(1, 701, 32).cuda() * 0.5
xr = lgnet.sa_attention(qq, qq.permute(0, 2, 1).contiguous(), torch.randn(1, 64, 701).cuda(), chunk=256)
torch.cuda.synchronize()
print("ok3", float(xr.abs().max()))
