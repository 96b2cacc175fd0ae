"""torch.profiler view of the training step (top CUDA kernels): python tools/prof_train.py [fp32|f16]"""
import os, sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DVM_TRAIN_PREC"] = sys.argv[1] if len(sys.argv) > 1 else "f16"
from dv_matcher_b200 import synthetic
from dv_matcher_b200.deformer import Deformer
from dv_matcher_b200.losses import GraphDeformLoss_Neural
from tools.bench_train import FeatureHead
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head, deformer = FeatureHead().to(dev), Deformer(10).to(dev)
crit = GraphDeformLoss_Neural(k_deform=10, w_dist=0.02, w_map=0.005, k_dist=500, N_dist=1000, partial=False, w_deform=0.5,
                              w_img=0, w_rank=0, w_self_rec=0.5, w_cd=0.1, w_arap=0.01, save_name="bench")
crit.cache_graphs = True
crit.graph_keys = ("a", "b")
d = {k: v.to(dev) for k, v in synthetic.make_batch(2, 4995, 4995).items()}
d1, d2 = torch.cdist(d["xyz1"], d["xyz1"]), torch.cdist(d["xyz2"], d["xyz2"])
def step():
    out = crit(head(d["feat1"]), head(d["feat2"]), d1, d2, d["xyz1"], d["xyz2"], 100.0, deformer)
    out[0].backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
if os.environ.get("CPU_TABLE"):
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
    import time
    t = time.perf_counter()
    for _ in range(20): step()
    t_issue = (time.perf_counter() - t) / 20
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t) / 20
    print(f"host issue time per step {t_issue * 1e3:.2f} ms ; wall per step {t_all * 1e3:.2f} ms")
