import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
n, b = int(sys.argv[1]), int(sys.argv[2])
d = synthetic.make_batch(b, n, n)
a, c = d["xyz1"].cuda(), d["xyz2"].cuda()
for _ in range(2):
    ops.knn3(a, a, 10); ops.chamfer_fwd(a, c)
torch.cuda.synchronize()
