"""The reference's secondary soft-map API (dead code in its entry points, mirrored for completeness):
cosine similarity -> top-k along rows AND columns -> softmax over the k -> weighted reconstruction.

    forward_source_target, forward_shape, reconstruction     test_partial.py:73-108
    measure_similarity("cosine") + get_s_t_neighbors          misc/switch_functions.py:121-135, misc/correspondence_utils.py:4-48
    cross_construct                                           test_partial.py:134-144

The reference materialises the N x M similarity matrix and two `topk`s over it.  Here the matrix only exists as row chunks of
tensor-core scores (`dvm_linear_act_fwd`, 3xTF32) that the radix selection (`dvm_topk_select`) consumes; the kept pairs'
similarities are recomputed exactly (`dvm_pair_dist_fwd`: cos = 1 - |a^ - b^|^2 / 2) and the reconstruction is the sparse
transfer kernel (`dvm_sparse_transfer_fwd`).  Inference only (the reference never differentiates through these).
"""
import torch

from . import ops


def _unit(f):
    """Rows scaled to unit length; the channel count is padded with zeros to a multiple of 4 (the kernels read float4 rows;
    zeros change neither dot products nor distances)."""
    f = f.float()
    f = f / f.norm(dim=-1)[:, :, None]
    pad = (-f.shape[-1]) % 4
    if pad:
        f = torch.nn.functional.pad(f, (0, pad))
    return f.contiguous()


def cosine_topk(a, b, k):
    """For every row of a [B,N,C]: the k rows of b [B,M,C] with the largest cosine similarity.
    Returns (sim [B,N,k] descending, idx int64 [B,N,k]) == `P.topk(k, dim=2)` of P = measure_similarity("cosine", a, b)."""
    an, bn = _unit(a), _unit(b)
    B, N, _ = an.shape
    k = min(k, bn.shape[1])
    idx = ops.knn_feature_large(an, bn, k)                      # |b^| = 1: the score a^.b^ - 1/2 orders like the cosine
    both = torch.cat([an, bn], 1)
    q = torch.arange(N, device=an.device).expand(B, N).contiguous()
    d, _ = ops.pair_dist_fwd(both, q, idx + N)
    return 1.0 - 0.5 * d * d, idx


def get_s_t_neighbors(k, feat_source, feat_target, s_only=False, ignore_first=False):
    """misc/correspondence_utils.py:30-48 with sim_normalization="softmax", taking the FEATURES instead of the dense matrix:
    (s_nn_weight, s_nn_sim, s_nn_idx, t_nn_weight, t_nn_sim, t_nn_idx)."""
    s_sim, s_idx = cosine_topk(feat_source, feat_target, k)
    if ignore_first:
        s_sim, s_idx = s_sim[:, :, 1:], s_idx[:, :, 1:]
    out = [torch.softmax(s_sim, dim=2), s_sim, s_idx]
    if s_only:
        return out + [None, None, None]
    t_sim, t_idx = cosine_topk(feat_target, feat_source, k)     # P.topk(k, dim=1) transposed: the k best sources of every target
    if ignore_first:
        t_sim, t_idx = t_sim[:, :, 1:], t_idx[:, :, 1:]
    return out + [torch.softmax(t_sim, dim=2), t_sim, t_idx]


def reconstruction(pos, nn_idx, nn_weight, k=None):
    """test_partial.py:73-80: (sum_k w * pos[idx], pos[idx[..., 0]])."""
    if k is not None:
        nn_idx, nn_weight = nn_idx[:, :, :k], nn_weight[:, :, :k]
    recon = ops.sparse_transfer_fwd(nn_idx.to(torch.int32).contiguous(), nn_weight.float().contiguous(), pos)
    hard = torch.gather(pos, 1, nn_idx[:, :, :1].expand(-1, -1, pos.shape[-1]))
    return recon, hard


def forward_source_target(feat_source, feat_target, vert_source, vert_target, k=40):
    """test_partial.py:82-96 -> (source_cross_recon [B,M,3], target_cross_recon [B,N,3])."""
    w_s, _, idx_s, w_t, _, idx_t = get_s_t_neighbors(k, feat_source, feat_target)
    source_cross_recon, _ = reconstruction(vert_source, idx_t, w_t)
    target_cross_recon, _ = reconstruction(vert_target, idx_s, w_s)
    return source_cross_recon, target_cross_recon


def forward_shape(feat, verts, k=40):
    """test_partial.py:98-108: reconstruction of every point from its k most similar OTHER points."""
    w, _, idx, _, _, _ = get_s_t_neighbors(k + 1, feat, feat, s_only=True, ignore_first=True)
    return reconstruction(verts, idx, w)[0]


def cross_construct(x, y, verts2, k_num):
    """test_partial.py:134-144: the k_num nearest columns (exact-form distance), softmax of their cosine similarities."""
    x, y = x.float().contiguous(), y.float().contiguous()
    B, N, _ = x.shape
    pad = (-x.shape[-1]) % 4
    if pad:                                          # zero channels change no distance
        x, y = torch.nn.functional.pad(x, (0, pad)), torch.nn.functional.pad(y, (0, pad))
    if k_num <= 10:
        idx = ops.softmap_fwd(x, y, None, alpha=1.0, topk=k_num, soft=True, prec="fp32").top_idx.long()     # exact, ascending distance
    else:
        idx = ops.knn_feature_large(x, y, k_num)
    both = torch.cat([_unit(x), _unit(y)], 1)
    q = torch.arange(N, device=x.device).expand(B, N).contiguous()
    d, _ = ops.pair_dist_fwd(both, q, idx + N)
    w = torch.softmax(1.0 - 0.5 * d * d, dim=2)
    return ops.sparse_transfer_fwd(idx.to(torch.int32).contiguous(), w.contiguous(), verts2)


__all__ = ["cosine_topk", "get_s_t_neighbors", "reconstruction", "forward_source_target", "forward_shape", "cross_construct"]
