#!/usr/bin/env python
"""Micro-benchmark of the 3-D k-NN / Chamfer / graph kernels: grid vs brute force.  python tools/bench_geom.py"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic  # noqa: E402
from dv_matcher_b200.deformation_graph import build_graphs  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for n, b in ((4995, 16), (20000, 4), (50000, 2), (200000, 1)):
    d = synthetic.make_batch(b, n, n)
    a, c = d["xyz1"].cuda(), d["xyz2"].cuda()
    row = dict(n=n, clouds=b)
    for algo in ("auto", "brute"):
        if algo == "brute" and n > 50000:
            continue
        row[f"knn10_{algo}_ms"] = round(timeit(lambda: ops.knn3(a, a, 10, algo=algo), 5), 4)
        row[f"chamfer_{algo}_ms"] = round(timeit(lambda: ops.chamfer_fwd(a, c, algo=algo), 5), 4)
    t0 = time.perf_counter()
    g = build_graphs(a, torch.zeros(b, dtype=torch.long))
    torch.cuda.synchronize()
    row["graph_build_cold_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    row["fps_ms"] = round(timeit(lambda: ops.fps(a, n // 2, torch.zeros(b, dtype=torch.long)), 2), 3)
    row["graph_weights_ms"] = round(timeit(lambda: ops.graph_weights(a, g.nodes_idx), 3), 3)
    print(json.dumps(row), flush=True)
