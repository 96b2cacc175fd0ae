"""Oracle: hard / soft correspondence maps.  TEST INFRASTRUCTURE (see oracle/__init__.py).

CPU torch fp32 restatement of models/loss.py:91-124, 1339-1347, 1408-1409 and
deform.py:63-77, 86-90, plus an fp64 arbiter and row-chunked forms that never
hold an N x M matrix (so the checker still runs at N = M = 50k).
"""
import numpy as np
import torch
import torch.nn.functional as F

TOPK = 10  # hard-coded in the reference: models/loss.py:1340, deform.py:70


# ----------------------------------------------------------------------------------------------
# verbatim restatements (dense N x M; small sizes)
# ----------------------------------------------------------------------------------------------
def knnsearch_t(x, y):
    """Hard NN map, 0-based, int64 [B,N,1].  models/loss.py:91-95 (exact direct-difference cdist)."""
    distance = torch.cdist(x.float(), y.float(), compute_mode="donot_use_mm_for_euclid_dist")
    _, idx = distance.topk(k=1, dim=-1, largest=False)
    return idx


def search_t(a1, a2, one_based=False):
    """models/loss.py:121-124 (0-based); test.py:19-28, deform.py:86-95 return idx+1."""
    t12 = knnsearch_t(a1, a2)
    return t12 + 1 if one_based else t12


def knnsearch_t_grad(x, y, alpha=100):
    """Dense soft map softmax(-alpha * cdist) [B,N,M].  models/loss.py:110-114 (GEMM-form cdist)."""
    distance = torch.cdist(x.float(), y.float())
    return F.softmax(-alpha * distance, dim=-1)


def topk_pi(a, k=TOPK):
    """Keep the k largest per row, un-renormalised, dense.  models/loss.py:1339-1347."""
    vals, idx = torch.topk(a, k, dim=-1)
    out = torch.zeros_like(a)
    out.scatter_(-1, idx, vals)
    return out


def transfer(pi, v):
    """Pi @ V.  models/loss.py:1408-1409, models/model.py:471."""
    return torch.matmul(pi, v)


def transfer_neighborhood(pi, v_nb):
    """einsum('bij,bjkm->bikm').  models/loss.py:1237."""
    return torch.einsum("bij,bjkm->bikm", pi, v_nb)


# ----------------------------------------------------------------------------------------------
# exact-form, sparse-output oracle (what the CUDA path is compared with)
# ----------------------------------------------------------------------------------------------
def _exact_d2_rows(x_rows, y, dtype):
    """[R,M] squared direct-difference distances, accumulated in `dtype`."""
    xr = x_rows.to(dtype)
    yy = y.to(dtype)
    return ((xr[:, None, :] - yy[None, :, :]) ** 2).sum(-1)


def softmap_sparse(x, y, alpha, k=TOPK, v=None, dtype=torch.float64, chunk=256, with_softmax=True):
    """Row-chunked soft map in sparse form; never materialises N x M.

    Semantics = knnsearch_t (arg-min) + topk_pi(knnsearch_t_grad) + Pi @ V, evaluated with the
    *exact* direct-difference distance in `dtype` (fp64 = the arbiter, fp32 = the fp32 restatement).
    Returns dict(argmin i64[B,N], idx i64[B,N,k] ascending distance, w [B,N,k], d [B,N,k],
    row_sum [B,N] (= sum_j exp(-alpha (d_j - d_min))), gap [B,N] (d of rank k+1 minus rank k),
    piv [B,N,Dv] or None).  Ties resolve to the lower index (torch.min / stable sort).
    """
    B, N, _ = x.shape
    M = y.shape[1]
    kk = min(k + 1, M)
    out = dict(argmin=torch.empty(B, N, dtype=torch.int64), idx=torch.empty(B, N, k, dtype=torch.int64),
               w=torch.empty(B, N, k, dtype=dtype), d=torch.empty(B, N, k, dtype=dtype),
               row_sum=torch.empty(B, N, dtype=dtype), gap=torch.empty(B, N, dtype=dtype),
               gap1=torch.empty(B, N, dtype=dtype))
    if v is not None:
        out["piv"] = torch.empty(B, N, v.shape[-1], dtype=dtype)
    for b in range(B):
        for s in range(0, N, chunk):
            e = min(N, s + chunk)
            d2 = _exact_d2_rows(x[b, s:e], y[b], dtype)
            d = d2.clamp_min(0).sqrt()
            ds, order = torch.sort(d, dim=-1, stable=True)
            out["argmin"][b, s:e] = order[:, 0]
            out["idx"][b, s:e] = order[:, :k]
            out["d"][b, s:e] = ds[:, :k]
            out["gap"][b, s:e] = (ds[:, kk - 1] - ds[:, k - 1]) if kk > k else float("inf")
            out["gap1"][b, s:e] = ds[:, 1] - ds[:, 0] if M > 1 else float("inf")
            if with_softmax:
                e_all = torch.exp(-alpha * (ds - ds[:, :1]))
                rs = e_all.sum(-1)
                w = e_all[:, :k] / rs[:, None]
            else:
                rs = torch.ones(e - s, dtype=dtype)
                w = torch.zeros(e - s, k, dtype=dtype)
            out["row_sum"][b, s:e] = rs
            out["w"][b, s:e] = w
            if v is not None:
                out["piv"][b, s:e] = (w[:, :, None] * v[b].to(dtype)[order[:, :k]]).sum(1)
    return out


def exact_d2_rows_blocked(x_rows, y, dtype=torch.float64, cblock=16, rblock=32):
    """[R,M] squared direct-difference distances like `_exact_d2_rows`, accumulated over channel blocks so that the
    temporary is [rblock, M, cblock] instead of [R, M, C] (M = 200k fits).  Same arithmetic form (sub, square, sum)."""
    xr = x_rows.to(dtype)
    yy = y.to(dtype)
    R, C = xr.shape
    out = torch.zeros(R, yy.shape[0], dtype=dtype)
    for r0 in range(0, R, rblock):
        acc = out[r0:r0 + rblock]
        for c0 in range(0, C, cblock):
            diff = xr[r0:r0 + rblock, None, c0:c0 + cblock] - yy[None, :, c0:c0 + cblock]
            acc += (diff * diff).sum(-1)
    return out


def softmap_from_d2(d2, alpha, k=TOPK, v=None):
    """Sparse soft-map quantities of `softmap_sparse` for a block of rows given their exact squared distances [R,M]
    (one sort serves every alpha: pass a list of alphas to get a list of results)."""
    alphas = list(alpha) if isinstance(alpha, (list, tuple)) else [alpha]
    d = d2.clamp_min(0).sqrt()
    ds, order = torch.sort(d, dim=-1, stable=True)
    M = d.shape[1]
    kk = min(k + 1, M)
    res = []
    for a in alphas:
        e_all = torch.exp(-a * (ds - ds[:, :1]))
        rs = e_all.sum(-1)
        w = e_all[:, :k] / rs[:, None]
        o = dict(argmin=order[:, 0].clone(), idx=order[:, :k].clone(), d=ds[:, :k].clone(), w=w, row_sum=rs,
                 gap=(ds[:, kk - 1] - ds[:, k - 1]) if kk > k else torch.full_like(rs, float("inf")),
                 gap1=(ds[:, 1] - ds[:, 0]) if M > 1 else torch.full_like(rs, float("inf")))
        if v is not None:
            o["piv"] = (w[:, :, None] * v.to(d.dtype)[order[:, :k]]).sum(1)
        res.append(o)
    return res if isinstance(alpha, (list, tuple)) else res[0]


def sparse_to_dense(idx, w, M):
    """scatter(idx, w) -> dense [B,N,M]; the form parity on Pi is defined on."""
    B, N, _ = idx.shape
    out = torch.zeros(B, N, M, dtype=w.dtype)
    out.scatter_(-1, idx.long(), w)
    return out


def hard_map_rows(x, y, rows, dtype=torch.float32):
    """Arg-min for a subset of rows of one pair (used at 50k where the full sweep is too slow on CPU)."""
    d2 = _exact_d2_rows(x[rows], y, dtype)
    return torch.min(d2, dim=-1)[1]


def near_tie_rows(gap, d, rel=1e-6):
    """Rows whose deciding gap is below fp32 resolution: index parity is undefined there."""
    return gap <= rel * d.abs().clamp_min(1e-30)


# ----------------------------------------------------------------------------------------------
# soft-map backward oracle (autograd through the exact-form dense expression; small sizes)
# ----------------------------------------------------------------------------------------------
def softmap_topk_dense_exact(x, y, alpha, k=TOPK):
    """Differentiable dense top-k soft map with the exact distance (fp64 inputs recommended)."""
    d = torch.cdist(x, y, compute_mode="donot_use_mm_for_euclid_dist")
    pi = F.softmax(-alpha * d, dim=-1)
    return topk_pi(pi, k)
