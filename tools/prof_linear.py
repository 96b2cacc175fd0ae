"""Two dvm_linear_act_fwd calls on one layer shape for ncu captures: python tools/prof_linear.py ROWS K N"""
import sys
import torch
sys.path.insert(0, ".")
from dv_matcher_b200 import ops
rows, K, N = (int(a) for a in sys.argv[1:4])
x = torch.randn(rows, (K + 3) // 4 * 4, device="cuda")[:, :K]
W = torch.randn(N, K, device="cuda") / K ** 0.5
b = torch.randn(N, device="cuda")
for _ in range(2):
    y = ops.linear_act_fwd(x, W, b, "elu")
torch.cuda.synchronize()
print("done", float(y[0, 0]))
