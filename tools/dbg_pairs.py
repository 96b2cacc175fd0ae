import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for p in [int(v) for v in sys.argv[1:]] or range(0, 34, 2):
    d = synthetic.make_batch(2, 50000, 50000, first_pair=p)
    x = torch.cat([d["feat1"], d["feat2"]]).cuda(); y = torch.cat([d["feat2"], d["feat1"]]).cuda(); v = torch.cat([d["xyz2"], d["xyz1"]]).cuda()
    o = ops.softmap_fwd(x, y, v, alpha=100.0, prec="f16", want_stats=True)
    print(p, "stats", o.stats.cpu().tolist(), "ms", round(t(lambda: ops.softmap_fwd(x, y, v, alpha=100.0, prec="f16")), 3), flush=True)
if os.environ.get("PROF"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        for _ in range(3): ops.softmap_fwd(x, y, v, alpha=100.0, prec="f16")
        torch.cuda.synchronize()
    print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
