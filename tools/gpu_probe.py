"""Per-GPU health probe for the multi-GPU scaling runs: each visible GPU, ALONE and in turn, runs a device copy, a bf16 GEMM and
the match+deform sweep, so that a rank that lags in `bench.py --gpus N` can be told apart from a GPU that is slower by itself.

    python tools/gpu_probe.py            -> one JSON line per GPU
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _time(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    from dv_matcher_b200 import ops, synthetic
    d = synthetic.make_batch(2, 50000, 50000)
    for g in range(torch.cuda.device_count()):
        torch.cuda.set_device(g)
        dev = torch.device("cuda", g)
        src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        dst = torch.empty_like(src)
        copy_ms = _time(lambda: dst.copy_(src), 10)
        a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
        mm_ms = _time(lambda: a @ b, 20)
        x = torch.cat([d["feat1"], d["feat2"]]).to(dev)
        y = torch.cat([d["feat2"], d["feat1"]]).to(dev)
        v = torch.cat([d["xyz2"], d["xyz1"]]).to(dev)
        sm_ms = _time(lambda: ops.softmap_fwd(x, y, v, alpha=100.0, prec="f16"), 10)
        print(json.dumps(dict(gpu=g, name=torch.cuda.get_device_name(g), copy_GBps=round(2 * (1 << 30) / copy_ms / 1e6, 1),
                              bf16_gemm_TFLOPs=round(2 * 8192 ** 3 / mm_ms / 1e9, 1), softmap_4x50k_ms=round(sm_ms, 3))), flush=True)
        del src, dst, a, b, x, y, v
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
