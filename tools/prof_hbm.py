#!/usr/bin/env python
"""One launch of each memory-bound deformation kernel at N = 200k x B = 32 for ncu captures:  python tools/prof_hbm.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dv_matcher_b200 import ops, synthetic
from dv_matcher_b200.deformation_graph import build_graphs, BatchedGraph
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(99)
npts, nb, base = 200000, 32, 2
clouds = torch.stack([synthetic.ellipsoid_cloud(npts, gen) for _ in range(base)]).to(dev)
g0 = build_graphs(clouds, torch.arange(base))
rep = nb // base
graphs = BatchedGraph.from_tensors([t.repeat(rep, *([1] * (t.dim() - 1))).contiguous() for t in g0.tensors()])
verts = clouds.repeat(rep, 1, 1).contiguous()
d9 = (0.05 * torch.randn(nb, npts // 2, 9, device=dev)).contiguous()
for _ in range(2):
    table = ops.node_table_from_d9(d9, graphs.pack.nodes_xyz)
    w = ops.skin_fwd_packed(verts, graphs.pack, table)
    a, s = ops.arap_fwd_packed(graphs.pack, table, want_sr=False)
    a2, s2 = ops.arap_fwd_packed(graphs.pack, table, want_sr=True)
torch.cuda.synchronize()
print("done", float(a[0]))
