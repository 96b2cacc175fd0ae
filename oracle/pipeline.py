"""Oracle: the reference's CPU composition of one "match + deform" pair, timed on a bounded row sample.

TEST / BASELINE INFRASTRUCTURE (only bench.py's cpu_baseline / --impl reference legs and tests call it).
Every step is the reference's own torch-op sequence (dense N x M matrices, models/loss.py:1404-1409,
1228-1282; models/model.py:464-478), restricted to a slab of R source rows so that a 50k-point pair --
which the reference cannot hold in memory -- can still be costed: slab work is
scaled by N / R, per-cloud work (skinning, ARAP, MLP on all K nodes) is run in full.
Graph construction is excluded, exactly as in the GPU timed region ("warm" graphs).
"""
import time

import torch
import torch.nn.functional as F

from . import geometry as og
from . import graph as ogr
from . import maps as om


def _slab_direction(feat_src, feat_tgt, verts_src, verts_tgt, rows, alpha, deformer_params, k=10):
    """Reference ops for ONE direction on `rows` of the source cloud. Returns dict of results."""
    x = feat_src[:, rows]
    pi = om.knnsearch_t_grad(x, feat_tgt, alpha)                       # cdist (GEMM form) + softmax   loss.py:1404
    pi = om.topk_pi(pi)                                                # top-10 + dense re-scatter     loss.py:1406
    verts_t = torch.matmul(pi, verts_tgt)                              # Pi @ verts2                   loss.py:1408
    t12 = om.knnsearch_t(x, feat_tgt)                                  # hard map (exact cdist + topk) loss.py:91-95
    idx11 = og.knn_grad(verts_src[:, rows], verts_src, k)              # xyz 10-NN                     loss.py:1229
    feat1_conv = og.index_points(feat_src, idx11)                      # [B,R,10,128] gather           loss.py:1253
    w = deformer_params["conv_layer.weight"].reshape(-1)
    b = deformer_params["conv_layer.bias"]
    feat1 = (feat1_conv * w[None, None, :, None]).sum(2) + b           # conv_layer                    model.py:468
    cd1 = og.sqdist_exact(verts_t, verts_tgt).min(-1)                  # chamfer rows->target (x2 calls: deformed, verts12)
    cd2 = og.sqdist_exact(verts_src[:, rows], verts_tgt).min(-1)
    return dict(pi=pi, verts_t=verts_t, t12=t12, feat1=feat1, cd=(cd1, cd2))


def _full_cloud_parts(feat_tgt, verts_src, verts_tgt, graph, deformer_params, k=10):
    """Per-cloud work that does not scale with the row slab: target-side conv gather, MLP, skinning + ARAP."""
    idx22 = og.knn_grad(verts_tgt[:, :256], verts_tgt, k)              # (slab of the target k-NN; scaled by caller)
    K = graph["nodes_idx"].shape[0]
    dev = verts_src.device
    x = torch.zeros(1, K, 262, device=dev)
    lin = "deformation_decoder_layer.linear."
    for i in (0, 2, 4):
        x = F.elu(F.linear(x, deformer_params[f"{lin}{i}.weight"], deformer_params[f"{lin}{i}.bias"]))
    d9 = F.linear(x, deformer_params[f"{lin}6.weight"], deformer_params[f"{lin}6.bias"])
    iden = torch.tensor([1, 0, 0, 0, 1, 0], dtype=torch.float32, device=dev)
    R = og.rotation_6d_to_matrix(d9[..., 3:] + iden)
    warped, arap, sr = ogr.dg_forward(verts_src[0], graph["nodes_idx"], graph["influence"], graph["weights"],
                                      graph["one_ring"], R, d9[..., :3])
    return idx22, warped, arap


def time_pair_sample(batch, graph, deformer_params, alpha=100.0, rows=2048, repeats=1, sync=None, results=None):
    """Seconds the reference's op sequence needs for ONE pair (both directions), extrapolated from a slab of `rows` source
    rows per direction (rows >= N: the full, un-extrapolated sequence).  batch: dict of [1,N,*] tensors (one pair) on the
    device to time (CPU = the reference's own CPU path; CUDA = "stock PyTorch on the same B200") -- `sync` is
    called before every clock read (torch.cuda.synchronize for CUDA tensors).  `results`, if a dict, receives the first
    direction's slab outputs (t12, verts_t, rows) and the raw wall time of the sample (sample_s) for parity reporting."""
    f1, f2, v1, v2 = batch["feat1"], batch["feat2"], batch["xyz1"], batch["xyz2"]
    N, M = f1.shape[1], f2.shape[1]
    dev = f1.device
    r1 = torch.arange(min(rows, N), device=dev)
    r2 = torch.arange(min(rows, M), device=dev)
    sync = sync or (lambda: None)
    best = float("inf")
    for _ in range(repeats):
        sync()
        t0 = time.perf_counter()
        with torch.no_grad():
            d12 = _slab_direction(f1, f2, v1, v2, r1, alpha, deformer_params)
            sync()
            t1 = time.perf_counter()
            _slab_direction(f2, f1, v2, v1, r2, alpha, deformer_params)
            sync()
            t2 = time.perf_counter()
            _full_cloud_parts(f2, v1, v2, graph, deformer_params)
            _full_cloud_parts(f1, v2, v1, graph, deformer_params)
            sync()
            t3 = time.perf_counter()
        est = (t1 - t0) * (N / len(r1)) + (t2 - t1) * (M / len(r2)) + (t3 - t2)
        if est < best:
            best = est
            if results is not None:
                results.update(t12=d12["t12"], verts_t=d12["verts_t"], rows=r1, sample_s=t3 - t0)
    return best


def make_cpu_graph(verts, start=0, max_nodes=None):
    """Small real graph for the per-cloud parts (node selection itself is excluded from the timing)."""
    n = verts.shape[0]
    verts = verts.cpu()
    if n > 6000:           # FPS on the CPU oracle is O(N^2): build the timing graph from strided nodes instead
        K = n // 2
        nodes = torch.arange(0, n, 2)[:K]
        infl = torch.randint(0, K, (n, 3))
        w = torch.full((n, 3), 1.0 / 3.0)
        ring = torch.randint(0, K, (K, 9))
        return dict(nodes_idx=nodes, influence=infl, weights=w, one_ring=ring)
    return ogr.construct_graph_euclidean(verts, start, exact=True)
