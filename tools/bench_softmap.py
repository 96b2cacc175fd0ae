#!/usr/bin/env python
"""Micro-benchmark of dvm_softmap_fwd: whole-call time and the candidate-pass kernel alone (CUDA events
recorded inside the library around the kernel), over sizes / precisions / regimes / alpha.

    python tools/bench_softmap.py [--quick]
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import _lib, ops, synthetic  # noqa: E402

PEAK = 1603.8                  # measured bf16 burst peak: these are stand-alone launches
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"]


def run(n, b, prec, alpha, regime, soft=True, iters=5):
    lib = _lib.load()
    d = synthetic.make_batch(b, n, n, regime=regime)
    x = torch.cat([d["feat1"], d["feat2"]]).cuda()
    y = torch.cat([d["feat2"], d["feat1"]]).cuda()
    v = torch.cat([d["xyz2"], d["xyz1"]]).cuda()
    for _ in range(2):
        o = ops.softmap_fwd(x, y, v if soft else None, alpha=alpha, topk=10 if soft else 1, soft=soft, prec=prec, want_stats=True)
    torch.cuda.synchronize()
    lib.dvm_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        o = ops.softmap_fwd(x, y, v if soft else None, alpha=alpha, topk=10 if soft else 1, soft=soft, prec=prec, want_stats=True)
    e1.record()
    torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
    lib.dvm_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
    lib.dvm_profile_enable(0)
    call_ms = e0.elapsed_time(e1) / iters
    cand_ms = tot.value / max(1, cnt.value)
    flops = 2.0 * n * n * 128 * 2 * b
    st = o.stats.cpu().tolist()
    return dict(n=n, problems=2 * b, prec=prec, alpha=alpha, regime=regime, soft=soft, call_ms=round(call_ms, 3), cand_ms=round(cand_ms, 3),
                cand_tflops=round(flops / cand_ms / 1e9, 1), frac_peak=round(flops / cand_ms / 1e9 / PEAK, 4),
                call_tflops=round(flops / call_ms / 1e9, 1), uncertified=st[0], rows=2 * b * n)


def main():
    quick = "--quick" in sys.argv
    if "--headline" in sys.argv:        # the bench configuration only: 4 problems of 50k, plus 5k and 20k at alpha = 100
        for n, b in ((50000, 2), (20000, 2), (4995, 16)):
            for soft in (True, False):
                print(json.dumps(run(n, b, "f16", 100.0, "structured", soft=soft)), flush=True)
        print(json.dumps(run(50000, 2, "f16", 100.0, "unstructured")), flush=True)
        print(json.dumps(run(50000, 1, "f16", 10.0, "structured", iters=2)), flush=True)
        return
    rows = []
    cfgs = [(4995, 8), (20000, 2)] if quick else [(4995, 8), (20000, 2), (50000, 1)]
    for n, b in cfgs:
        for regime in ("structured", "unstructured"):
            for alpha in (100.0, 10.0):
                for prec in ("f16",) if quick else ("f16", "bf16"):
                    rows.append(run(n, b, prec, alpha, regime))
                    print(json.dumps(rows[-1]), flush=True)
        rows.append(run(n, b, "f16", 100.0, "structured", soft=False))
        print(json.dumps(rows[-1]), flush=True)
        if n <= 20000:
            rows.append(run(n, b, "fp32", 100.0, "structured", iters=2))
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
