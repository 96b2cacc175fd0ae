"""Training losses with the reference's constructor arguments, forward signature and 5-tuple result.

Mirrors models/loss.py:1075-1435 (`GraphDeformLoss_Neural`) and 726-1073 (`GraphDeformLoss_Neural_Partial`),
`FrobeniusLoss` (476-482).  The orchestration is the reference's; what changes underneath:

  * graphs: `deformation_graph_node` builds all B graphs in batched kernels (no N x N cdist, no host copy);
  * Pi_12 / Pi_21: one fused launch sequence per direction, `SparseSoftMap` instead of dense [B,N,M] tensors;
  * `deform`: the per-batch-element Python loop over `dg(...)` (models/loss.py:1269-1273) is one batched
    skinning + ARAP launch; gathers + 1x1 conv of the Deformer are fused; Chamfer is `dvm_chamfer_*`;
  * the print at :1282 (which costs a third Chamfer evaluation) and the four OFF dumps at :1284-1295 only
    happen with DVM_REFERENCE_SIDE_EFFECTS=1.

Every differentiable step has a CUDA backward (dvm_softmap_bwd, dvm_sparse_transfer_bwd, dvm_gather_conv_bwd,
dvm_rot6d_bwd, dvm_skin_bwd, dvm_arap_bwd, dvm_chamfer_bwd); the MLP and the dist-loss cosine stay torch.
"""
import os
import random

import torch
import torch.nn as nn

from . import geometry, maps
from .deformation_graph import BatchedGraph, build_graphs, deform_batched, draw_fps_start
from .geometry import chamfer_3DDist, index_points, knn, knn_grad, rotation_6d_to_matrix

_IDEN6 = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)


def _side_effects():
    return os.environ.get("DVM_REFERENCE_SIDE_EFFECTS", "0") == "1"


class FrobeniusLoss(nn.Module):
    """models/loss.py:476-482: sum of squares over axes (1, 2), mean over what is left."""

    def forward(self, a, b):
        loss = torch.sum(torch.abs(a - b) ** 2, axis=(1, 2))
        return torch.mean(loss)


def save_off_file(path, pts):
    """deform.py:79-84 / models/loss.py OFF writer: vertices only."""
    with open(path, "w") as f:
        f.write("OFF\n")
        f.write(f"{pts.shape[0]} 0 0\n")
        for p in pts:
            f.write(f"{p[0]} {p[1]} {p[2]}\n")


def dist_loss_term(feat, dist, n_sample, k, numbers=None):
    """One shape's half of the dist loss (models/loss.py:1362-1374 / 1376-1390).

    feat [B,N,C], dist [B,N,N] (geodesic matrix, any float dtype).  `numbers` = the N_dist sampled query
    indices (the reference draws them with random.sample on the host RNG).  Returns sum(1 - |cos|).
    """
    B, num = dist.shape[0], dist.shape[1]
    if numbers is None:
        numbers = random.sample(range(num), n_sample)
    rn = torch.as_tensor(numbers, device=feat.device, dtype=torch.long)
    f1 = feat[:, rn]                                               # [B,S,C]
    idx = knn(f1, feat, k)                                         # [B,S,k]  tcgen05 scores + radix selection (dvm_topk_select)
    if feat.is_cuda and feat.shape[-1] % 4 == 0 and dist.is_contiguous() and dist.dtype in (torch.float32, torch.float64):
        # ||feat[idx] - f1|| and dist[i, idx[i], idx_num[i]] in one gather kernel with its own backward: the reference builds
        # f2 = index_points(feat, idx) ([B,S,k,C]: 256 MB per shape at S = 1000, k = 500) and loops over B in Python
        dist_result, dist_f = geometry.pair_dist(feat, rn, idx, dist)
    else:
        dist_result = torch.gather(torch.cdist(f1, feat, compute_mode="donot_use_mm_for_euclid_dist"), 2, idx)   # [B,S,k]
        bidx = torch.arange(B, device=feat.device)[:, None, None]
        dist_f = dist[bidx, idx, rn[None, :, None]].float()        # torch.zeros_like(idx, dtype=float) in the reference
    return torch.sum(1 - torch.abs(torch.nn.functional.cosine_similarity(dist_result, dist_f, dim=2)))


class _GraphDeformBase(nn.Module):
    def __init__(self, k_deform=10, w_dist=1, w_map=1, k_dist=1000, N_dist=1000, partial=False, w_deform=1, w_img=1,
                 w_rank=1, w_self_rec=1, w_cd=1, w_arap=1, save_name=None):
        super().__init__()
        self.device = "cuda:0"
        self.input_pts = 4995
        self.w_dist, self.w_map, self.w_deform, self.w_self_rec = w_dist, w_map, w_deform, w_self_rec
        self.w_cd, self.w_arap, self.w_rank, self.w_img = w_cd, w_arap, w_rank, w_img
        self.k_dist, self.N_dist, self.k_deform = k_dist, N_dist, k_deform
        self.dist_loss = 0
        self.deform_loss = 0
        self.self_rec_loss = 0
        self.img_loss = 0
        self.rank_loss = 0
        self.map_loss = 0
        self.partial = partial
        self.frob_loss = FrobeniusLoss()
        self.chamfer_dist_3d = chamfer_3DDist()
        self.save_name = save_name
        # graphs are a pure function of the geometry + the FPS start index: with cache_graphs = True they are cached under
        # the keys the CALLER supplies per forward through `graph_keys = (key of verts1, key of verts2)` (e.g. the dataset
        # indices of the two shape batches) -- tensor addresses are reused by the caching allocator and identify nothing.
        # Off by default so that every forward redraws the start like the reference does.  Bounded (FIFO).
        self.cache_graphs = False
        self.graph_keys = None
        self.max_cached_graphs = 256
        self._graph_cache = {}
        self._graph_slot = 0
        self._iden6 = {}
        # set by training.CapturedTrainStep: graphs and sampled query indices live in STATIC device buffers that are refilled
        # before every replay of the captured step (nothing inside forward may touch the host then)
        self.static_graphs = None        # [(nodes_idx float [B,K], BatchedGraph) for verts1, same for verts2]
        self.static_draws = None         # (int64 [N_dist] for shape 1, same for shape 2)

    # -- pieces of the reference API -----------------------------------------------------------------
    def topk_pi(self, A):
        return maps.topk_pi(A)

    def _batch_frobenius_norm(self, matrix1, matrix2):
        return torch.norm((matrix1 - matrix2), dim=(1, 2))

    def deformation_graph_node(self, verts1):
        """Batched graphs of models/loss.py:1325-1337.  Returns (nodes_idx [B,K] float like the reference's
        `num_nodes_all`, BatchedGraph) -- the second item replaces the reference's list of per-cloud objects."""
        if self.static_graphs is not None:
            out = self.static_graphs[self._graph_slot % 2]
            self._graph_slot += 1
            return out
        key = None
        if self.cache_graphs:
            if self.graph_keys is None:
                raise RuntimeError("cache_graphs = True needs graph_keys = (key1, key2) identifying the two shape batches of this forward")
            key = (self.graph_keys[self._graph_slot % 2], tuple(verts1.shape))
            self._graph_slot += 1
            hit = self._graph_cache.get(key)
            if hit is not None:
                return hit
        g = build_graphs(verts1.detach(), draw_fps_start(verts1.shape[0], verts1.shape[1]))
        out = (g.nodes_idx.float(), g)
        if key is not None:
            if len(self._graph_cache) >= self.max_cached_graphs:
                self._graph_cache.pop(next(iter(self._graph_cache)))
            self._graph_cache[key] = out
        return out

    def _identity6(self, ref):
        """[1,0,0,0,1,0] of models/loss.py:1259-1262 as a cached device constant (a per-step torch.tensor(list, device=...)
        is a blocking host-to-device copy)."""
        k = (ref.device, ref.dtype)
        t = self._iden6.get(k)
        if t is None:
            t = torch.tensor(_IDEN6, device=ref.device, dtype=ref.dtype)
            self._iden6[k] = t
        return t

    def _dist_loss(self, feat1, feat2, dist1, dist2):
        if self.static_draws is not None:
            numbers1, numbers2 = self.static_draws
        else:
            numbers1 = random.sample(range(dist1.shape[1]), self.N_dist)      # same host-RNG draw order as :1361-1364
            numbers2 = random.sample(range(dist2.shape[1]), self.N_dist)
        s1 = dist_loss_term(feat1, dist1, self.N_dist, self.k_dist, numbers1)
        s2 = dist_loss_term(feat2, dist2, self.N_dist, self.k_dist, numbers2)
        return (s1 + s2) * self.w_dist

    def _deform_common(self, verts12, verts1, Pi_12, verts2, k, fps1, graph, feat1, feat2, deformer):
        """Shared part of both `deform` variants: Deformer -> R, t -> warped source, ARAP."""
        idx11 = knn_grad(verts1, verts1, k)
        idx22 = knn_grad(verts2, verts2, k)
        if hasattr(deformer, "forward_fused"):
            deformations = deformer.forward_fused(feat1, feat2, idx11, idx22, verts1, verts12, Pi_12, fps1)
        else:                                   # a foreign Deformer: the reference's call, gathers materialised
            deformations = deformer(index_points(feat1, idx11), index_points(feat2, idx22), verts1, verts12, Pi_12, fps1)
        rotations = deformations[:, :, 3:] + self._identity6(deformations)
        T1 = deformations[:, :, :3]
        R1 = rotation_6d_to_matrix(rotations)
        deformed_points1, arap, _sr = deform_batched(verts1, graph, R1, T1)
        return idx11, idx22, deformed_points1, arap.sum()

    def _dump(self, deformed_points1, verts1, verts2, verts12, cd_term, arap_term):
        n = str(random.randint(0, 10))
        print(f"Rand:{n}, Deform_Result: cd_loss:{cd_term}, arap_loss:{arap_term}")
        save_path_t = "visual_result/" + str(self.save_name)
        os.makedirs(save_path_t, exist_ok=True)
        for tag, t in (("deform_", deformed_points1), ("target_", verts2), ("source_", verts1), ("pi_verts2_", verts12)):
            save_off_file(f"{save_path_t}/{tag}{n}.off", t[0].detach().cpu().squeeze().numpy())


class GraphDeformLoss_Neural(_GraphDeformBase):
    """Full-shape loss (models/loss.py:1075-1435)."""

    def chamfer_loss(self, pos1, pos2):
        return geometry.chamfer_loss(pos1, pos2)

    def deform(self, verts12, verts1, Pi_12, verts2, k, fps1, dg_list1, feat1, feat2, deformer):
        idx11, idx22, deformed_points1, arap_all = self._deform_common(verts12, verts1, Pi_12, verts2, k, fps1, dg_list1,
                                                                        feat1, feat2, deformer)
        if self.w_map > 0:
            verts2_corr_neighbor = index_points(verts12, idx11)                                  # [B,N,k,3]
            verts2_neighbor = index_points(verts2, idx22)
            verts2_neighbor_corr = torch.einsum("bij, bjkm->bikm", Pi_12, verts2_neighbor)      # 10-sparse transfer
            map_loss12 = self.frob_loss(verts2_corr_neighbor, verts2_neighbor_corr)
        else:
            map_loss12 = 0
        cd = self.chamfer_loss(deformed_points1, verts2)
        cross_deform_loss12 = cd * self.w_cd + arap_all * self.w_arap
        self_rec_loss12 = self.chamfer_loss(verts12, verts2)
        if _side_effects():
            self._dump(deformed_points1, verts1, verts2, verts12, cd * self.w_cd, arap_all * self.w_arap)
        return map_loss12, cross_deform_loss12, self_rec_loss12

    def forward(self, feat1, feat2, dist1, dist2, verts1, verts2, alpha_i, deformer):
        loss = 0
        if self.w_dist > 0:
            self.dist_loss = self._dist_loss(feat1, feat2, dist1, dist2)
            loss += self.dist_loss
        B, N, _ = verts1.shape
        k = self.k_deform
        self._graph_slot = 0
        num_nodes_all1, dg1 = self.deformation_graph_node(verts1)
        num_nodes_all2, dg2 = self.deformation_graph_node(verts2)
        Pi_12 = self.topk_pi(maps.knnsearch_t_grad(feat1, feat2, alpha=alpha_i))
        Pi_21 = self.topk_pi(maps.knnsearch_t_grad(feat2, feat1, alpha=alpha_i))
        verts12 = torch.matmul(Pi_12, verts2)
        verts21 = torch.matmul(Pi_21, verts1)
        map_loss12, cross12, self_rec12 = self.deform(verts12, verts1, Pi_12, verts2, k, num_nodes_all1.long(), dg1, feat1, feat2, deformer)
        map_loss21, cross21, self_rec21 = self.deform(verts21, verts2, Pi_21, verts1, k, num_nodes_all2.long(), dg2, feat2, feat1, deformer)
        self.deform_loss = (cross12 + cross21) * N * self.w_deform / 2
        loss += self.deform_loss
        if self.w_map > 0:
            self.map_loss = self.w_map * (map_loss12 + map_loss21) / 2
            loss += self.map_loss
        if self.w_self_rec > 0:
            self.self_rec_loss = (self_rec12 + self_rec21) * N * self.w_self_rec / 2
            loss += self.self_rec_loss
        if self.w_rank > 0:                    # weight 0 in every shipped config; dense N x N, reference formula
            I_N = torch.eye(n=N, device=verts1.device).unsqueeze(0).repeat(B, 1, 1)
            P12, P21 = Pi_12.to_dense(), Pi_21.to_dense()
            self.rank_loss = (torch.mean(self._batch_frobenius_norm(torch.bmm(P12, P12.transpose(2, 1).contiguous()), I_N.float()))
                              + torch.mean(self._batch_frobenius_norm(torch.bmm(P21, P21.transpose(2, 1).contiguous()), I_N.float()))) * self.w_rank / 2
            loss += self.rank_loss
        return loss, self.dist_loss, self.deform_loss, self.map_loss, self.self_rec_loss


class GraphDeformLoss_Neural_Partial(_GraphDeformBase):
    """Partial-to-full loss (models/loss.py:726-1073): one-sided Chamfer, no map term, no x N scaling."""

    def chamfer_loss(self, pos1, pos2):
        return geometry.chamfer_loss_partial(pos1, pos2)

    def deform(self, verts12, verts1, Pi_12, verts2, k, fps1, dg_list1, feat1, feat2, deformer):
        _, _, deformed_points1, arap_all = self._deform_common(verts12, verts1, Pi_12, verts2, k, fps1, dg_list1,
                                                                feat1, feat2, deformer)
        cd = self.chamfer_loss(deformed_points1, verts2)
        cross_deform_loss12 = cd * self.w_cd + arap_all * self.w_arap
        self_rec_loss12 = self.chamfer_loss(verts12, verts2)
        if _side_effects():
            self._dump(deformed_points1, verts1, verts2, verts12, cd * self.w_cd, arap_all * self.w_arap)
        return cross_deform_loss12, self_rec_loss12

    def forward(self, feat1, feat2, dist1, dist2, verts1, verts2, alpha_i, deformer):
        loss = 0
        if self.w_dist > 0:
            self.dist_loss = self._dist_loss(feat1, feat2, dist1, dist2)
            loss += self.dist_loss
        self_rec12 = self_rec21 = None
        if self.w_deform > 0:
            k = self.k_deform
            self._graph_slot = 0
            num_nodes_all1, dg1 = self.deformation_graph_node(verts1)
            num_nodes_all2, dg2 = self.deformation_graph_node(verts2)
            Pi_12 = self.topk_pi(maps.knnsearch_t_grad(feat1, feat2, alpha=alpha_i))
            Pi_21 = self.topk_pi(maps.knnsearch_t_grad(feat2, feat1, alpha=alpha_i))
            verts12 = torch.matmul(Pi_12, verts2)
            verts21 = torch.matmul(Pi_21, verts1)
            cross12, self_rec12 = self.deform(verts12, verts1, Pi_12, verts2, k, num_nodes_all1.long(), dg1, feat1, feat2, deformer)
            cross21, self_rec21 = self.deform(verts21, verts2, Pi_21, verts1, k, num_nodes_all2.long(), dg2, feat2, feat1, deformer)
            self.deform_loss = (cross12 + cross21) * self.w_deform / 2
            loss += self.deform_loss
        if self.w_self_rec > 0:
            if self_rec12 is None:             # the reference raises NameError here (w_deform == 0, :1057); be explicit
                raise RuntimeError("GraphDeformLoss_Neural_Partial: w_self_rec > 0 needs w_deform > 0 (as in the reference)")
            self.self_rec_loss = (self_rec12 + self_rec21) * self.w_self_rec / 2
            loss += self.self_rec_loss
        if self.w_rank > 0:
            B, N, _ = verts1.shape
            M = verts2.shape[1]
            P12, P21 = Pi_12.to_dense(), Pi_21.to_dense()
            I_N = torch.eye(n=N, device=verts1.device).unsqueeze(0).repeat(B, 1, 1)
            I_M = torch.eye(n=M, device=verts1.device).unsqueeze(0).repeat(B, 1, 1)
            self.rank_loss = (torch.mean(self._batch_frobenius_norm(torch.bmm(P12, P12.transpose(2, 1).contiguous()), I_N.float()))
                              + torch.mean(self._batch_frobenius_norm(torch.bmm(P21, P21.transpose(2, 1).contiguous()), I_M.float()))) * self.w_rank / 2
            loss += self.rank_loss
        return loss, self.dist_loss, self.deform_loss, self.map_loss, self.self_rec_loss


__all__ = ["FrobeniusLoss", "GraphDeformLoss_Neural", "GraphDeformLoss_Neural_Partial", "BatchedGraph", "dist_loss_term", "save_off_file"]
