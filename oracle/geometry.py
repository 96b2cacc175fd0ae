"""Oracle: xyz k-NN, gathers, Chamfer, rotation parametrisations.  TEST INFRASTRUCTURE.

Follows models/loss.py:39-45, 97-101, 451-473, 1216-1226, 867-882; lib/utils.py:70-112.
Chamfer restates ThibaultGROUEIX/ChamferDistancePytorch chamfer3D (un-vendored, unpinned:
parity unpinned -- see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def knn_grad(x, y, k):
    """Reference form: GEMM-mode cdist + topk, int64 [B,N,k].  models/loss.py:97-101.

    NOTE: for D=3 the GEMM form is not reproducible near zero distance; the CUDA
    path is judged against `knn_exact` and disagreements with this function are shown to lie
    inside the reference's own rounding bound.
    """
    distance = torch.cdist(x.float(), y.float())
    _, idx = distance.topk(k=k, dim=-1, largest=False)
    return idx


def sqdist_exact(x, y, dtype=torch.float32):
    """[B,N,M] squared distance, direct differences, left-to-right channel sum, no FMA."""
    x = x.to(dtype)
    y = y.to(dtype)
    D = x.shape[-1]
    acc = None
    for c in range(D):
        diff = x[..., :, None, c] - y[..., None, :, c]
        sq = diff * diff
        acc = sq if acc is None else acc + sq
    return acc


def knn_exact(x, y, k, dtype=torch.float32):
    """k-NN by exact squared distance; ties -> lower index; returns (idx i64 [B,N,k], d2 [B,N,k])."""
    d2 = sqdist_exact(x, y, dtype)
    ds, order = torch.sort(d2, dim=-1, stable=True)
    return order[..., :k].contiguous(), ds[..., :k].contiguous()


def knn_feature(a, b, k):
    """Feature-space k-NN by -||a-b||^2 in the hand-written GEMM form.  models/loss.py:451-462."""
    inner = -2 * torch.matmul(a, b.transpose(2, 1))
    aa = torch.sum(a ** 2, dim=2, keepdim=True)
    bb = torch.sum(b ** 2, dim=2, keepdim=True)
    pairwise_distance = -aa - inner - bb.transpose(2, 1)
    return pairwise_distance.topk(k=k, dim=-1)[1]


def index_points(points, idx):
    """Batched gather [B,N,C] x [B,S,K] -> [B,S,K,C].  models/loss.py:464-473."""
    raw_shape = idx.shape
    idx = idx.reshape(raw_shape[0], -1)
    res = torch.gather(points, 1, idx[..., None].expand(-1, -1, points.shape[-1]))
    return res.view(*raw_shape, -1)


def index_points_idx(points, idx):
    """points[b, idx[b]] for idx [B,K].  models/loss.py:440-449."""
    B = points.shape[0]
    return points[torch.arange(B)[:, None], idx, :]


def rotation_6d_to_matrix(d6):
    """Gram-Schmidt 6D -> R with rows b1,b2,b3.  models/loss.py:39-45."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def batch_rodrigues(axisang):
    """Axis-angle [K,3] -> R [K,3,3] via quaternion.  lib/utils.py:70-112."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=1).unsqueeze(-1)
    n = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * n], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


# ----------------------------------------------------------------------------------------------
# Chamfer (chamfer_3DDist): squared distance, both directions, int32 arg-min, lowest index on ties
# ----------------------------------------------------------------------------------------------
def chamfer_3d(a, b, chunk=2048):
    """dist1[b,i]=min_j||a_i-b_j||^2, dist2, idx1, idx2 (int32).  Call sites models/loss.py:1223, 874."""
    B, N, _ = a.shape
    M = b.shape[1]
    dist1 = torch.empty(B, N)
    idx1 = torch.empty(B, N, dtype=torch.int32)
    dist2 = torch.empty(B, M)
    idx2 = torch.empty(B, M, dtype=torch.int32)
    for bi in range(B):
        for s in range(0, N, chunk):
            d2 = sqdist_exact(a[bi:bi + 1, s:s + chunk], b[bi:bi + 1])[0]
            v, i = torch.min(d2, dim=1)  # first index on ties == ascending scan with strict '<'
            dist1[bi, s:s + chunk] = v
            idx1[bi, s:s + chunk] = i.int()
        for s in range(0, M, chunk):
            d2 = sqdist_exact(b[bi:bi + 1, s:s + chunk], a[bi:bi + 1])[0]
            v, i = torch.min(d2, dim=1)
            dist2[bi, s:s + chunk] = v
            idx2[bi, s:s + chunk] = i.int()
    return dist1, dist2, idx1, idx2


def chamfer_3d_backward(a, b, idx1, idx2, g1, g2):
    """grad wrt a and b: 2*g*(a - b[idx]) gathered + scattered from the other direction."""
    B = a.shape[0]
    da = torch.zeros_like(a)
    db = torch.zeros_like(b)
    for bi in range(B):
        i1 = idx1[bi].long()
        i2 = idx2[bi].long()
        t1 = 2 * g1[bi][:, None] * (a[bi] - b[bi][i1])
        da[bi] += t1
        db[bi].index_add_(0, i1, -t1)
        t2 = 2 * g2[bi][:, None] * (b[bi] - a[bi][i2])
        db[bi] += t2
        da[bi].index_add_(0, i2, -t2)
    return da, db


def chamfer_loss_full(a, b):
    """mean(dist1) + mean(dist2).  models/loss.py:1216-1226."""
    d1, d2, _, _ = chamfer_3d(a, b)
    return d1.mean() + d2.mean()


def chamfer_loss_partial(a, b):
    """One-sided mean from the smaller cloud.  models/loss.py:867-882."""
    d1, d2, _, _ = chamfer_3d(a, b)
    return d1.mean() if d1.shape[1] <= d2.shape[1] else d2.mean()
