"""xyz k-NN, gathers, rotation parametrisation and Chamfer with the reference's call surface.

Mirrors models/loss.py:39-45 (`rotation_6d_to_matrix`), 97-101 (`knn_grad`), 451-473 (`knn`,
`index_points`), 440-449 (`index_points_idx`) and the `dist_chamfer_3D.chamfer_3DDist` module of
ChamferDistancePytorch that models/loss.py:14,1099,1223 imports.
"""
import torch
import torch.nn as nn

from . import ops


def knn_grad(x, y, k):
    """k nearest neighbours of x in y, int64 [B,N,k] (models/loss.py:97-101, deform.py:24-28).

    The reference calls this on xyz only.  3-D inputs run the exact brute-force kernel (direct
    differences -- the reference's GEMM-form cdist is not reproducible near zero distance, SURVEY
    section 7); feature-space inputs (C % 4 == 0, k <= 10) go through the fused similarity kernel.
    """
    if x.shape[-1] == 3 and k <= 16:
        return ops.knn3(x, y, k)
    if x.shape[-1] % 4 == 0 and k <= 10:
        return ops.softmap_fwd(x, y, None, topk=k, soft=False, prec="fp32").top_idx.long()
    raise NotImplementedError(f"knn_grad: unsupported feature width {x.shape[-1]} / k={k}")


def knn(a, b, k, chunk=4096):
    """Feature-space k-NN by -||a-b||^2 in the reference's hand-written GEMM form (models/loss.py:451-462), int64 [B,S,k]
    in descending score order like `topk`.

    k <= 10: the fused similarity kernel (exact distances).  Larger k (500 / 300 for the 1000 sampled queries of the dist loss,
    models/loss.py:1367,1380): the scores 2 a.b - |b|^2 on tcgen05 (3xTF32, fp32-equivalent) + dvm_topk_select (radix selection
    of the k best per row, one CTA per row).  Rows that are too long for the GEMM's bias staging are scored in column chunks by
    stock matmul and selected by the same kernel."""
    if k <= 10 and a.shape[-1] % 4 == 0:
        return ops.softmap_fwd(a, b, None, topk=k, soft=False, prec="fp32").top_idx.long()
    if a.is_cuda and a.shape[-1] % 4 == 0 and k <= 1024:
        return ops.knn_feature_large(a.detach(), b.detach(), k)
    bb = torch.sum(b ** 2, dim=2, keepdim=True).transpose(2, 1)
    out = []
    for s in range(0, a.shape[1], chunk):
        ac = a[:, s:s + chunk]
        inner = -2 * torch.matmul(ac, b.transpose(2, 1))
        aa = torch.sum(ac ** 2, dim=2, keepdim=True)
        out.append((-aa - inner - bb).topk(k=k, dim=-1)[1])
    return torch.cat(out, dim=1)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx2):
        ctx.save_for_backward(idx2)
        ctx.n = points.shape[1]
        return ops.gather_rows_fwd(points, idx2)

    @staticmethod
    def backward(ctx, d_out):
        (idx2,) = ctx.saved_tensors
        return ops.gather_rows_bwd(d_out.contiguous(), idx2, ctx.n), None


def index_points(points, idx):
    """[B,N,C] gathered by idx [B,S,K] -> [B,S,K,C] (models/loss.py:464-473): dvm_gather_rows_fwd/bwd."""
    raw_shape = idx.shape
    idx2 = idx.reshape(raw_shape[0], -1).contiguous()
    if not points.is_cuda or points.dtype != torch.float32:
        res = torch.gather(points, 1, idx2[..., None].expand(-1, -1, points.shape[-1]))
    else:
        res = _GatherRows.apply(points.contiguous(), idx2.long())
    return res.view(*raw_shape, -1)


class _PairDist(torch.autograd.Function):
    """(feat [B,N,C], q [S], nbr [B,S,k]) -> |feat[nbr] - feat[q]| [B,S,k]; geo gathered alongside when given."""

    @staticmethod
    def forward(ctx, feat, qidx, nbr, geo):
        feat = feat.float().contiguous()
        d, g = ops.pair_dist_fwd(feat, qidx, nbr, geo)
        ctx.save_for_backward(feat, qidx, nbr, d)
        if g is None:
            g = torch.empty(0, device=feat.device)
        ctx.mark_non_differentiable(g)
        return d, g

    @staticmethod
    def backward(ctx, gd, _gg):
        feat, qidx, nbr, d = ctx.saved_tensors
        return ops.pair_dist_bwd(feat, qidx, nbr, d, gd.contiguous()), None, None, None


def pair_dist(feat, qidx, nbr, geo=None):
    """torch.norm(index_points(feat, nbr) - feat[:, qidx][:, :, None, :], dim=-1) and geo[b, nbr, qidx] (models/loss.py:1366-1378)."""
    d, g = _PairDist.apply(feat, qidx.long().contiguous(), nbr.long().contiguous(), geo)
    return d, (g if geo is not None else None)


def index_points_idx(points, idx):
    """points[b, idx[b]] (models/loss.py:440-449)."""
    B = points.shape[0]
    return points[torch.arange(B, device=points.device)[:, None], idx, :]


class _Rot6D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d6):
        d6 = d6.float().contiguous()
        ctx.save_for_backward(d6)
        return ops.rot6d_fwd(d6)

    @staticmethod
    def backward(ctx, dR):
        (d6,) = ctx.saved_tensors
        return ops.rot6d_bwd(d6, dR.contiguous())


def rotation_6d_to_matrix(d6):
    """Gram-Schmidt 6D -> rotation matrix with rows b1, b2, b3 (models/loss.py:39-45)."""
    return _Rot6D.apply(d6)


class _GatherConv(torch.autograd.Function):
    """index_points(feat, idx) followed by Conv2d(k -> 1, 1x1) over the neighbour axis, fused."""

    @staticmethod
    def forward(ctx, feat, idx, weight, bias):
        feat = feat.float().contiguous()
        ctx.save_for_backward(feat, idx, weight)
        ctx.has_bias = bias is not None
        return ops.gather_conv_fwd(feat, idx, weight, bias)

    @staticmethod
    def backward(ctx, d_out):
        feat, idx, weight = ctx.saved_tensors
        d_feat, d_w, d_b = ops.gather_conv_bwd(feat, idx, weight, d_out.contiguous())
        return d_feat, None, d_w.reshape(weight.shape), (d_b if ctx.has_bias else None)


def gather_conv(feat, idx, conv_weight, conv_bias):
    """conv_layer(index_points(feat, idx).permute(0,2,1,3)).squeeze(1) of models/model.py:468-469
    without the [B,N,k,C] intermediate (256 MB per shape at N = 50k)."""
    return _GatherConv.apply(feat, idx, conv_weight, conv_bias)


# --------------------------------------------------------------------------------------------------
# Chamfer
# --------------------------------------------------------------------------------------------------
class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = a.float().contiguous(), b.float().contiguous()
        d1, d2, i1, i2 = ops.chamfer_fwd(a, b)
        ctx.save_for_backward(a, b, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2, i1, i2

    @staticmethod
    def backward(ctx, g1, g2, _i1, _i2):
        a, b, i1, i2 = ctx.saved_tensors
        da, db = ops.chamfer_bwd(a, b, i1, i2, g1.contiguous(), g2.contiguous())
        return da, db


class chamfer_3DDist(nn.Module):
    """Drop-in for ChamferDistancePytorch's `dist_chamfer_3D.chamfer_3DDist()`:
    forward(a [B,N,3], b [B,M,3]) -> (dist1 [B,N], dist2 [B,M], idx1 int32, idx2 int32), squared distances."""

    def forward(self, input1, input2):
        return _Chamfer.apply(input1, input2)


class _DistChamfer3DNamespace:
    """So that `dist_chamfer_3D.chamfer_3DDist()` (models/loss.py:1099) resolves after install()."""
    chamfer_3DDist = chamfer_3DDist


dist_chamfer_3D = _DistChamfer3DNamespace()


def chamfer_loss(pos1, pos2):
    """mean(dist1) + mean(dist2) (models/loss.py:1216-1226)."""
    d1, d2, _, _ = _Chamfer.apply(pos1, pos2)
    return torch.mean(d1) + torch.mean(d2)


def chamfer_loss_partial(pos1, pos2):
    """One-sided mean from the cloud with fewer points (models/loss.py:867-882)."""
    d1, d2, _, _ = _Chamfer.apply(pos1, pos2)
    return torch.mean(d1) if d1.shape[1] <= d2.shape[1] else torch.mean(d2)
