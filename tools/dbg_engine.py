import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import pipeline, synthetic
from dv_matcher_b200.deformer import Deformer
from dv_matcher_b200.deformation_graph import build_graphs
dev = torch.device("cuda", 0)
torch.manual_seed(0)
deformer = Deformer(10).to(dev).eval()
n, nb = 50000, 2
slots = int(os.environ.get("SLOTS", "3")); depth = int(os.environ.get("DEPTH", "3")); nring = int(os.environ.get("NRING", "3"))
ring = [synthetic.make_batch(nb, n, n, first_pair=q * nb, pin=True) for q in range(nring)]
eng = pipeline.MatchDeformEngine(deformer, alpha=100.0, prec="f16", device=dev, slots=slots)
for q in range(nring):
    eng.put_graphs(q, build_graphs(torch.cat([ring[q]["xyz1"], ring[q]["xyz2"]]).to(dev), torch.arange(2 * nb) % n))
torch.cuda.synchronize()
pending = []
def estep(i):
    h = ring[i % nring]
    pending.append(eng.submit(h["feat1"], h["feat2"], h["xyz1"], h["xyz2"], graph_key=i % nring))
    if len(pending) == depth: eng.result(pending.pop(0))
for i in range(12): estep(i)
while pending: eng.result(pending.pop(0))
torch.cuda.synchronize(); t = time.perf_counter()
K = 60
for i in range(K): estep(12 + i)
while pending: eng.result(pending.pop(0))
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"slots {slots} depth {depth} nring {nring}: {nb * K / dt:.1f} pairs/s  {dt / K * 1e3:.3f} ms/step", flush=True)
