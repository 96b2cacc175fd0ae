"""TEST INFRASTRUCTURE -- CPU restatement of the reference's secondary soft-map API (dead code in the
reference's entry points, kept for completeness).  Dense torch, small sizes only.  Pinned against outputs of the unmodified
reference functions (tests/golden/make_golden_secondary.py -> ref_secondary.npz).

    measure_similarity("cosine")            misc/switch_functions.py:121-135
    get_s_t_topk / get_s_t_neighbors        misc/correspondence_utils.py:4-48   (top-k along rows AND columns, softmax over the k)
    reconstruction                          test_partial.py:73-80
    forward_source_target / forward_shape   test_partial.py:82-108
    cross_construct                         test_partial.py:134-144
"""
import torch
import torch.nn.functional as F


def cosine_similarity_matrix(a, b):
    """misc/switch_functions.py:131-134."""
    an = a / a.norm(dim=-1)[:, :, None]
    bn = b / b.norm(dim=-1)[:, :, None]
    return torch.bmm(an, bn.transpose(1, 2))


def s_t_neighbors(P, k, ignore_first=False, s_only=False):
    """misc/correspondence_utils.py:4-48 with sim_normalization="softmax": (w_s, sim_s, idx_s, w_t, sim_t, idx_t)."""
    s_val, s_idx = P.topk(k=min(k, P.shape[2]), dim=2)
    if ignore_first:
        s_val, s_idx = s_val[:, :, 1:], s_idx[:, :, 1:]
    out = [F.softmax(s_val, dim=2), s_val, s_idx]
    if s_only:
        return out + [None, None, None]
    t_val, t_idx = P.topk(k=k, dim=1)
    t_val, t_idx = t_val.transpose(2, 1), t_idx.transpose(2, 1)
    if ignore_first:
        t_val, t_idx = t_val[:, :, 1:], t_idx[:, :, 1:]
    return out + [F.softmax(t_val, dim=2), t_val, t_idx]


def reconstruction(pos, nn_idx, nn_weight):
    """test_partial.py:73-80: weighted sum of the neighbours' positions, and the first neighbour's position."""
    B = pos.shape[0]
    nn_pos = torch.stack([pos[b][nn_idx[b]] for b in range(B)])          # [B,R,k,3]
    return (nn_pos * nn_weight.unsqueeze(3)).sum(2), nn_pos[:, :, 0, :]


def forward_source_target(feat_source, feat_target, vert_source, vert_target, k=40):
    """test_partial.py:82-96 -> (source_cross_recon [B,M,3], target_cross_recon [B,N,3])."""
    P = cosine_similarity_matrix(feat_source, feat_target)
    w_s, _, idx_s, w_t, _, idx_t = s_t_neighbors(P, k)
    src_recon, _ = reconstruction(vert_source, idx_t, w_t)
    tgt_recon, _ = reconstruction(vert_target, idx_s, w_s)
    return src_recon, tgt_recon


def forward_shape(feat, verts, k=40):
    """test_partial.py:98-108: self reconstruction from the k most similar OTHER points."""
    P = cosine_similarity_matrix(feat, feat)
    w, _, idx, _, _, _ = s_t_neighbors(P, k + 1, ignore_first=True, s_only=True)
    return reconstruction(verts, idx, w)[0]


def cross_construct(x, y, verts2, k_num):
    """test_partial.py:134-144: k_num nearest (exact-form distance) columns, softmax of their cosine similarities."""
    d = torch.cdist(x.float(), y.float(), compute_mode="donot_use_mm_for_euclid_dist")
    _, idx = d.topk(k=k_num, dim=-1, largest=False)
    B = x.shape[0]
    v = torch.stack([verts2[b][idx[b]] for b in range(B)])
    f2 = torch.stack([y[b].float()[idx[b]] for b in range(B)])
    sim = F.cosine_similarity(x.float().unsqueeze(2).expand_as(f2), f2, dim=3)
    return (v * F.softmax(sim, dim=2).unsqueeze(-1)).sum(2)
