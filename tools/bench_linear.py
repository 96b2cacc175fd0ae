"""Correctness (vs fp64) and timing of dvm_linear_act_fwd on the Deformer MLP's layer shapes."""
import json
import sys
import torch
sys.path.insert(0, ".")
from dv_matcher_b200 import ops

torch.manual_seed(0)
dev = "cuda"
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dims = [262, 512, 256, 128, 9]
res = {}
for li in range(4):
    K, N = dims[li], dims[li + 1]
    pitch = (K + 3) // 4 * 4
    xb = torch.randn(rows, pitch, device=dev)
    x = xb[:, :K]
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    act = "elu" if li < 3 else "none"
    y = ops.linear_act_fwd(x, W, b, act)
    ref64 = torch.nn.functional.linear(x[:4096].double(), W.double(), b.double())
    if act == "elu":
        ref64 = torch.nn.functional.elu(ref64)
    ref32 = torch.nn.functional.linear(x[:4096], W, b)
    if act == "elu":
        ref32 = torch.nn.functional.elu(ref32)
    e_tc = ((y[:4096].double() - ref64).abs().max() / ref64.abs().max()).item()
    e_32 = ((ref32.double() - ref64).abs().max() / ref64.abs().max()).item()
    tail_ok = bool(torch.equal(y[-1:], ops.linear_act_fwd(x[-1:].contiguous(), W, b, act))) if K % 4 == 0 else None
    for _ in range(3):
        ops.linear_act_fwd(x, W, b, act)
    t0, t1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0.record()
    for _ in range(10):
        ops.linear_act_fwd(x, W, b, act)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    t0.record()
    for _ in range(10):
        r = torch.nn.functional.linear(x, W, b)
        if act == "elu":
            r = torch.nn.functional.elu(r)
    t1.record(); torch.cuda.synchronize()
    ms_t = t0.elapsed_time(t1) / 10
    res[f"L{li}"] = dict(K=K, N=N, err_tc=e_tc, err_fp32=e_32, ms=ms, ms_torch=ms_t, tf_eff=2 * rows * K * N / ms / 1e9, tail=tail_ok)
    print(json.dumps(res[f"L{li}"]), flush=True)
