#!/usr/bin/env python
"""One tcgen05 soft-map backward (plus a warm-up) for ncu captures:  python tools/prof_bwd.py [N=4995] [B=2] [ALPHA=100]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4995
b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
alpha = float(sys.argv[3]) if len(sys.argv) > 3 else 100.0
d = synthetic.make_batch(b, n, n)
x, y = d["feat1"].cuda(), d["feat2"].cuda()
out = ops.softmap_fwd(x, y, None, alpha=alpha, prec="f16")
dw = torch.randn(b, n, 10, device="cuda")
for _ in range(2):
    gx, gy = ops.softmap_bwd(x, y, alpha, out, dw, prec="f16")
torch.cuda.synchronize()
print("done", float(gx.abs().max()))
