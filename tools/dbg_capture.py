import os, sys, random, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import synthetic, training
from dv_matcher_b200.deformer import Deformer
from dv_matcher_b200.deformation_graph import build_graphs
from dv_matcher_b200.losses import GraphDeformLoss_Neural
from tools.bench_train import FeatureHead
dev = torch.device("cuda", 0)
torch.manual_seed(0)
head, deformer = FeatureHead().to(dev), Deformer(10).to(dev)
params = list(head.parameters()) + list(deformer.parameters())
def mk():
    return GraphDeformLoss_Neural(k_deform=10, w_dist=0.02, w_map=0.005, k_dist=500, N_dist=1000, partial=False, w_deform=0.5,
                                  w_img=0, w_rank=0, w_self_rec=0.5, w_cd=0.1, w_arap=0.01, save_name="bench")
batches, graphs = [], []
for q in range(2):
    d = {k: v.to(dev) for k, v in synthetic.make_batch(2, 4995, 4995, first_pair=2 * q).items()}
    d["dist1"], d["dist2"] = torch.cdist(d["xyz1"], d["xyz1"]), torch.cdist(d["xyz2"], d["xyz2"])
    batches.append(d)
    z = torch.zeros(2, dtype=torch.int64, device=dev)
    graphs.append((build_graphs(d["xyz1"], z), build_graphs(d["xyz2"], z)))
# eager reference
crit = mk(); crit.cache_graphs = True
res_e = []
for q in range(2):
    random.seed(5 + q)
    for p in params: p.grad = None
    crit.graph_keys = (("a", q), ("b", q))
    crit._graph_cache[(("a", q), tuple(batches[q]["xyz1"].shape))] = (graphs[q][0].nodes_idx.float(), graphs[q][0])
    crit._graph_cache[(("b", q), tuple(batches[q]["xyz2"].shape))] = (graphs[q][1].nodes_idx.float(), graphs[q][1])
    d = batches[q]
    out = crit(*head_out, d["dist1"], d["dist2"], d["xyz1"], d["xyz2"], 100.0, deformer) if False else crit(head(d["feat1"]), head(d["feat2"]), d["dist1"], d["dist2"], d["xyz1"], d["xyz2"], 100.0, deformer)
    out[0].backward()
    res_e.append(([float(o) for o in out], torch.cat([p.grad.reshape(-1) for p in params]).clone()))
# captured (no reference to an eager autograd graph may be alive: its AccumulateGrad nodes would pin the default stream)
del out, crit
import gc; gc.collect()
crit2 = mk()
step = training.CapturedTrainStep(crit2, lambda a, b: (head(a), head(b)), deformer, params, 100.0)
for rep in range(2):
    for q in range(2):
        random.seed(5 + q)
        out = step(batches[q], graphs[q])
        torch.cuda.synchronize()
        g = torch.cat([p.grad.reshape(-1) for p in params])
        le, ge = res_e[q]
        print("batch", q, "loss eager", [round(v, 5) for v in le], "captured", [round(float(o), 5) for o in out],
              "grad rel diff", float((g - ge).abs().max() / ge.abs().max()), flush=True)
t = time.perf_counter()
for i in range(20): step(batches[i % 2], graphs[i % 2])
torch.cuda.synchronize()
print("captured ms/step", (time.perf_counter() - t) / 20 * 1e3)
