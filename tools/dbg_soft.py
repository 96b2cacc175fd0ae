import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
for n, regime, alpha in ((4995, "structured", 100.0), (4995, "unstructured", 100.0), (3000, "structured", 100.0), (20000, "structured", 100.0)):
    d = synthetic.make_batch(1, n, n, regime=regime)
    x, y, v = d["feat1"].cuda(), d["feat2"].cuda(), d["xyz2"].cuda()
    o = ops.softmap_fwd(x, y, v, alpha=alpha, prec="f16", want_stats=True)
    r = ops.softmap_fwd(x, y, v, alpha=alpha, prec="fp32")
    print(n, regime, alpha, "stats", o.stats.cpu().tolist(), "row_sum max rel diff", ((o.row_sum - r.row_sum).abs() / r.row_sum).max().item())
