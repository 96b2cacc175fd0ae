"""Oracle: deformation graph (node sampling, skinning weights, ring, warp, ARAP).  TEST INFRASTRUCTURE.

Follows lib/deformation_graph_point.py:18-33 (FPS), 177-201 (construct_graph_euclidean),
233-261 (forward); lib/deformation_graph.py:89-116 (axis-angle variant);
models/loss.py:1325-1337 (driver); models/model.py:454-478 (Deformer).
"""
import numpy as np
import torch

from .geometry import batch_rodrigues, index_points_idx, sqdist_exact


def farthest_point_sample(xyz, npoint, start):
    """FPS with the start index injected.  lib/deformation_graph_point.py:18-33.

    xyz [N,3] fp32.  distance init 1e10; dist = (dx*dx + dy*dy) + dz*dz unfused; masked min update;
    next = first arg-max.  Returns int64 [npoint].
    """
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype=np.float32))
    N = xyz.shape[0]
    out = np.empty(npoint, dtype=np.int64)
    distance = np.full(N, np.float32(1e10), dtype=np.float32)
    far = int(start)
    for i in range(npoint):
        out[i] = far
        diff = xyz - xyz[far]
        sq = diff * diff
        dist = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
        np.minimum(distance, dist, out=distance)
        far = int(np.argmax(distance))
    return out


def knn_f64(q, r, k):
    """Exact k-NN on fp32 coordinates evaluated in fp64 (what SciPy's KDTree.query returns).

    Returns (dist f64 [Nq,k], idx i64 [Nq,k]) ascending distance, ties -> lower index.
    """
    q64 = torch.as_tensor(np.asarray(q), dtype=torch.float64)
    r64 = torch.as_tensor(np.asarray(r), dtype=torch.float64)
    out_d = torch.empty(q64.shape[0], k, dtype=torch.float64)
    out_i = torch.empty(q64.shape[0], k, dtype=torch.int64)
    for s in range(0, q64.shape[0], 1024):
        d2 = sqdist_exact(q64[None, s:s + 1024], r64[None], torch.float64)[0]
        ds, order = torch.sort(d2, dim=-1, stable=True)
        out_d[s:s + 1024] = ds[:, :k].sqrt()
        out_i[s:s + 1024] = order[:, :k]
    return out_d, out_i


def construct_graph_euclidean(vertices, start, k=3, ring=9, exact=True):
    """Graph tensors for one cloud.  lib/deformation_graph_point.py:177-201 with models/loss.py:1333.

    vertices [N,3] fp32 torch.  `exact=True` evaluates the vertex->node distances with the exact
    direct-difference fp32 form (the policy of SURVEY section 7); `exact=False` reproduces the reference's
    GEMM-form `torch.cdist` matrix verbatim.
    Returns dict(nodes_idx i64[K], one_ring i64[K,ring], influence i64[N,k], dists f32[N,k],
                 weights f32[N,k], sigma f64 scalar).
    """
    v = vertices.float().contiguous()
    N = v.shape[0]
    K = N // 2
    nodes_idx = farthest_point_sample(v.numpy(), K, start)
    nodes = v[torch.from_numpy(nodes_idx)]
    _, one_ring = knn_f64(nodes, nodes, ring)                      # KDTree(nodes).query(nodes, 9)   :181-183
    if exact:
        d = sqdist_exact(v[None], nodes[None])[0].sqrt()           # [N,K]
    else:
        d = torch.cdist(v, v, p=2.0)[torch.from_numpy(nodes_idx)].t()  # -geod[nodes].T              :186
    ds, order = torch.sort(d, dim=-1, stable=True)                 # topk(k) of the negated matrix   :187-188
    dists, influence = ds[:, :k].contiguous(), order[:, :k].contiguous()
    nn2, _ = knn_f64(v, v, 2)                                      # KDTree(vertices).query(.,2)     :190-191
    sigma = 20.0 * nn2[:, 1].mean()                                # float64                         :192
    w = torch.exp(-(dists ** 2) / (2 * sigma * sigma))             # fp32 tensor / 0-dim f64 -> fp32 :195-197
    w = w / w.sum(1, keepdim=True)                                 #                                  :198
    return dict(nodes_idx=torch.from_numpy(nodes_idx), one_ring=one_ring, influence=influence,
                dists=dists, weights=w.float(), sigma=sigma)


def dg_forward(vertices, nodes_idx, influence, weights, one_ring, R, t):
    """Skinning warp + ARAP + rotation smoothness for one cloud.  lib/deformation_graph_point.py:233-261.

    vertices [N,3], R [1,K,3,3], t [1,K,3].  Returns ([1,N,3], arap, sr).
    """
    nodes = vertices[nodes_idx]
    K = nodes.shape[0]
    ring = one_ring.shape[1]
    flat = influence.reshape(-1)
    inf_v = nodes[flat]
    r = R[0, flat]
    tt = t[0, flat]
    warped = (torch.einsum("bij,bkj->bki", r, (vertices.repeat_interleave(influence.shape[1], dim=0) - inf_v).unsqueeze(1)).squeeze(1)
              + inf_v + tt).reshape(vertices.shape[0], influence.shape[1], 3) * weights.unsqueeze(-1)
    warped = warped.sum(1).float()
    rflat = one_ring.reshape(-1)
    diff = (nodes + t[0]).repeat_interleave(ring, dim=0) - (nodes[rflat] + t[0][rflat]) - \
        torch.einsum("bij,bkj->bki", R[0].repeat_interleave(ring, dim=0),
                     (nodes.repeat_interleave(ring, dim=0) - nodes[rflat]).unsqueeze(1)).squeeze(1)
    sr_term = R[0].repeat_interleave(ring, dim=0) - R[0][rflat]
    sr = torch.mean(sr_term ** 2)
    arap = torch.sum(diff ** 2) / K
    return warped.unsqueeze(0), arap, sr


def dg_forward_axis_angle(vertices, nodes_idx, influence, weights, one_ring, axis_angle, t):
    """lib/deformation_graph.py:89-116: same with R = batch_rodrigues(axis_angle[0])."""
    R = batch_rodrigues(axis_angle[0]).unsqueeze(0)
    return dg_forward(vertices, nodes_idx, influence, weights, one_ring, R, t)


def deformer_forward(params, feat1_conv, feat2_conv, verts1, verts12, pi_dense, fps1):
    """Deformer.forward.  models/model.py:464-478.  `params` = the checkpoint state dict."""
    import torch.nn.functional as F
    w = params["conv_layer.weight"].reshape(-1)   # [k]
    b = params["conv_layer.bias"]
    feat1 = (feat1_conv * w[None, None, :, None]).sum(2) + b
    feat2 = (feat2_conv * w[None, None, :, None]).sum(2) + b
    feat2 = torch.matmul(pi_dense, feat2)
    x = torch.cat([index_points_idx(verts1, fps1), index_points_idx(feat1, fps1),
                   index_points_idx(verts12, fps1), index_points_idx(feat2, fps1)], dim=-1)
    lin = "deformation_decoder_layer.linear."
    for i in (0, 2, 4):
        x = F.elu(F.linear(x, params[f"{lin}{i}.weight"], params[f"{lin}{i}.bias"]))
    return F.linear(x, params[f"{lin}6.weight"], params[f"{lin}6.bias"])
