import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
from torch.profiler import profile, ProfilerActivity
nb = int(os.environ.get("NB", "16")); n = int(os.environ.get("NPTS", "4995"))
d = synthetic.make_batch(nb, n, n)
x = torch.cat([d["feat1"], d["feat2"]]).cuda(); y = torch.cat([d["feat2"], d["feat1"]]).cuda(); v = torch.cat([d["xyz2"], d["xyz1"]]).cuda()
for soft in (True, False):
    f = (lambda: ops.softmap_fwd(x, y, v, alpha=100.0, prec="f16")) if soft else (lambda: ops.softmap_fwd(x, y, None, topk=1, soft=False, prec="f16"))
    for _ in range(3): f()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        for _ in range(5): f()
        torch.cuda.synchronize()
    print("soft" if soft else "hard")
    print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
