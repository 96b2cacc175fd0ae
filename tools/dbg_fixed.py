"""Per-cluster cost model of the sweep: ONE wave of CTA pairs (N = 74 x 256 rows, B = 1), time vs number of column tiles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
from torch.profiler import profile, ProfilerActivity
d = synthetic.make_batch(1, 51200, 51200)
N = 74 * 256
for soft in (True, False):
    for M in [int(v) for v in os.environ.get("MS", "2560,5120,10240,20480,51200").split(",")]:
        x = d["feat1"][:, :N].contiguous().cuda(); y = d["feat2"][:, :M].contiguous().cuda()
        f = (lambda: ops.softmap_fwd(x, y, None, alpha=100.0, prec="f16")) if soft else (lambda: ops.softmap_fwd(x, y, None, topk=1, soft=False, prec="f16"))
        for _ in range(3): f()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            for _ in range(5): f()
            torch.cuda.synchronize()
        t = {}
        for e in pr.key_averages():
            if "softmap_cand_tc_kernel" in e.key:
                t["prime" if "false, true" in e.key else "sweep"] = e.device_time_total / e.count
        tiles = M // 256
        print(f"soft={soft} M={M} tiles={tiles}: sweep {t.get('sweep', 0):8.1f} us  prime {t.get('prime', 0):6.1f} us  per tile {t.get('sweep', 0) / tiles:6.3f} us", flush=True)
