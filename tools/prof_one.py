#!/usr/bin/env python
"""One softmap_fwd call (plus a warm-up) for ncu captures:  python tools/prof_one.py N B PREC ALPHA REGIME [hard]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic  # noqa: E402

n, b, prec, alpha, regime = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], float(sys.argv[4]), sys.argv[5]
soft = not (len(sys.argv) > 6 and sys.argv[6] == "hard")
d = synthetic.make_batch(b, n, n, regime=regime)
x = torch.cat([d["feat1"], d["feat2"]]).cuda()
y = torch.cat([d["feat2"], d["feat1"]]).cuda()
v = torch.cat([d["xyz2"], d["xyz1"]]).cuda()
for _ in range(2):
    o = ops.softmap_fwd(x, y, v if soft else None, alpha=alpha, topk=10 if soft else 1, soft=soft, prec=prec)
torch.cuda.synchronize()
print("done", o.argmin[0, :4].tolist())
