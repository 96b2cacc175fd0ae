"""Deformation graph with the reference's call surface, everything on the GPU.

Mirrors lib/deformation_graph_point.py: `farthest_point_sample` (18-33), `DeformationGraph_geod`
(71-261: `construct_graph_euclidean` 177-201, `forward` 233-261), lib/deformation_graph.py's
axis-angle `DeformationGraph` (17-116) and the driver `deformation_graph_node`
(models/loss.py:1325-1337, deform.py:41-53).

What changes underneath: no N x N `cdist` matrix, no device->host copy of it, no SciPy KD-trees, no
2497-iteration Python FPS loop.  Node selection (dvm_fps), the 3 influencing nodes + Gaussian weights,
the 9-NN node ring and sigma (dvm_graph_weights) and the warp/ARAP forward+backward (dvm_skin_*,
dvm_arap_*) are CUDA kernels, batched over the clouds of a batch.
"""
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .geometry import rotation_6d_to_matrix  # noqa: F401  (re-exported for callers of the reference's helper)


@dataclass
class GraphPack:
    """HBM-shaped copy of the graph tensors the per-step kernels stream (static per shape; see csrc/deform_packed.cu).
    Nodes are RENUMBERED in Morton order of their positions: node_perm[new] = old (the reference's FPS order)."""
    node_perm: torch.Tensor     # int32 [B,K]     new -> old node number
    nodes_idx_m: torch.Tensor   # int64 [B,K]     vertex index of every node, NEW order (what the Deformer is asked for: `fps1`)
    nodes_xyz: torch.Tensor     # f32   [B,K,3]   node positions, new order
    vorder: torch.Tensor        # int32 [B,N]     vertices in Morton order
    s_xyz: torch.Tensor         # f32   [B,N,3]   coordinates of vertex vorder[i]
    s_infl: torch.Tensor        # int32 [B,3,N]   influence lists (new node numbers) of vertex vorder[i], slot-major
    s_w: torch.Tensor           # f32   [B,3,N]
    s_ring: torch.Tensor        # int32 [B,9,K]   ring of new node i (new numbers), slot-major
    csr_ptr: torch.Tensor       # int32 [B,K+1]   vertices influenced by each new node ...
    csr_vert: torch.Tensor      # int32 [B,3N]    ... ascending vertex id
    csr_w: torch.Tensor         # f32   [B,3N]

    def tensors(self):
        return (self.node_perm, self.nodes_idx_m, self.nodes_xyz, self.vorder, self.s_xyz, self.s_infl, self.s_w, self.s_ring,
                self.csr_ptr, self.csr_vert, self.csr_w)

    def to_old_order(self, per_node):
        """[B,K,...] tensor with rows in the packed (new) node order -> the reference's node order."""
        idx = self.node_perm.long().reshape(*self.node_perm.shape, *([1] * (per_node.dim() - 2))).expand_as(per_node)
        return torch.empty_like(per_node).scatter_(1, idx, per_node)


@dataclass
class BatchedGraph:
    """Graph tensors of B clouds (same N, K): everything stays on the device."""
    nodes_idx: torch.Tensor     # int64 [B,K]  vertex index of every node (== `num_nodes_all` of the reference driver)
    influence: torch.Tensor     # int64 [B,N,3] node index space
    dists: torch.Tensor         # f32  [B,N,3]
    weights: torch.Tensor       # f32  [B,N,3]
    ring: torch.Tensor          # int64 [B,K,9] node index space, self first
    sigma: torch.Tensor         # f64  [B]
    pack: GraphPack = None      # streaming layout of the same graph (built by `pack_graph`)

    def tensors(self):
        base = (self.nodes_idx, self.influence, self.dists, self.weights, self.ring, self.sigma)
        return base + (self.pack.tensors() if self.pack is not None else ())

    @staticmethod
    def from_tensors(ts):
        ts = list(ts)
        return BatchedGraph(*ts[:6], pack=GraphPack(*ts[6:]) if len(ts) > 6 else None)

    @staticmethod
    def cat(graphs):
        return BatchedGraph.from_tensors([torch.cat(parts) for parts in zip(*[g.tensors() for g in graphs])])

    def select(self, i):
        """Graph of cloud i as a batch of one."""
        return BatchedGraph.from_tensors([t[i:i + 1] for t in self.tensors()])


def draw_fps_start(B, N):
    """The reference's RNG draw for the first centroid (lib/deformation_graph_point.py:24):
    one `torch.randint(0, N, (1,))` on the global CPU generator per cloud, in batch order."""
    return torch.cat([torch.randint(0, N, (1,), dtype=torch.long) for _ in range(B)])


def farthest_point_sample(xyz, npoint, start=None):
    """lib/deformation_graph_point.py:18-33 for xyz [B,N,3]; `start` [B] injects the first index."""
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    return ops.fps(xyz, npoint, start)


def _morton_order(p):
    """argsort of 30-bit Morton codes of points p [B,n,3] (per-cloud bounding box), int32 [B,n]."""
    lo = p.amin(1, keepdim=True)
    ext = (p.amax(1, keepdim=True) - lo).clamp_min(1e-20)
    q = ((p - lo) / ext * 1023.0).long().clamp_(0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        return (v | (v << 2)) & 0x09249249

    code = spread(q[..., 0]) | (spread(q[..., 1]) << 1) | (spread(q[..., 2]) << 2)
    return torch.argsort(code, dim=1, stable=True).to(torch.int32)


def pack_graph(verts, nodes_idx, influence, weights, ring):
    """Streaming layout of a graph (once per shape; plain torch ops -- this is graph construction, not the per-step path)."""
    B, N, _ = verts.shape
    K = nodes_idx.shape[1]
    g_old = torch.gather(verts, 1, nodes_idx[..., None].expand(B, K, 3))
    node_perm = _morton_order(g_old)                                      # new -> old
    no = node_perm.long()
    inv = torch.empty_like(no).scatter_(1, no, torch.arange(K, device=verts.device).expand(B, K))     # old -> new
    nodes_idx_m = torch.gather(nodes_idx, 1, no).contiguous()
    nodes_xyz = torch.gather(g_old, 1, no[..., None].expand(B, K, 3)).contiguous()
    infl_new = torch.gather(inv, 1, influence.reshape(B, 3 * N)).reshape(B, N, 3)                     # new node numbers
    vorder = _morton_order(verts)
    vo = vorder.long()
    s_xyz = torch.gather(verts, 1, vo[..., None].expand(B, N, 3)).contiguous()
    s_infl = torch.gather(infl_new, 1, vo[..., None].expand(B, N, 3)).transpose(1, 2).to(torch.int32).contiguous()
    s_w = torch.gather(weights, 1, vo[..., None].expand(B, N, 3)).transpose(1, 2).contiguous()
    rk = ring.shape[2]
    ring_rows = torch.gather(ring, 1, no[..., None].expand(B, K, rk))                                  # rings of the nodes in new order ...
    s_ring = torch.gather(inv, 1, ring_rows.reshape(B, K * rk)).reshape(B, K, rk).transpose(1, 2).to(torch.int32).contiguous()   # ... new numbers
    flat = infl_new.reshape(B, 3 * N)
    order = torch.argsort(flat, dim=1, stable=True)                       # entries (v, k) sorted by new node, ascending v within a node
    csr_vert = (order // 3).to(torch.int32).contiguous()
    csr_w = torch.gather(weights.reshape(B, 3 * N), 1, order).contiguous()
    counts = torch.zeros(B, K, dtype=torch.int64, device=verts.device).scatter_add_(1, flat, torch.ones_like(flat))
    csr_ptr = torch.zeros(B, K + 1, dtype=torch.int32, device=verts.device)
    csr_ptr[:, 1:] = counts.cumsum(1).to(torch.int32)
    return GraphPack(node_perm, nodes_idx_m, nodes_xyz, vorder, s_xyz, s_infl, s_w, s_ring, csr_ptr, csr_vert, csr_w)


def build_graphs(verts, start=None):
    """construct_graph_euclidean for a batch verts [B,N,3] (K = N // 2 nodes, 3 influences, ring of 9)."""
    B, N, _ = verts.shape
    if start is None:
        start = draw_fps_start(B, N)
    verts = verts.float().contiguous()
    nodes_idx = ops.fps(verts, N // 2, start)
    infl, dists, wts, ring, sigma = ops.graph_weights(verts, nodes_idx)
    return BatchedGraph(nodes_idx, infl, dists, wts, ring, sigma, pack_graph(verts, nodes_idx, infl, wts, ring))


class _Deform(torch.autograd.Function):
    """(R [B,K,3,3], t [B,K,3]) -> warped [B,N,3], arap [B], sr [B]; geometry and graph carry no gradient."""

    @staticmethod
    def forward(ctx, verts, R, t, graph, want_sr):
        verts, R, t = verts.float().contiguous(), R.float().contiguous(), t.float().contiguous()
        table = ops.node_table(R, t, graph.pack.nodes_xyz, node_perm=graph.pack.node_perm)      # R, t arrive in the reference's node order
        warped = ops.skin_fwd_packed(verts, graph.pack, table)
        arap, sr = ops.arap_fwd_packed(graph.pack, table, want_sr)
        ctx.save_for_backward(verts, R, t)
        ctx.graph = graph
        if sr is None:
            sr = torch.zeros_like(arap)
        ctx.mark_non_differentiable(sr)
        return warped, arap, sr

    @staticmethod
    def backward(ctx, d_warped, d_arap, _d_sr):
        verts, R, t = ctx.saved_tensors
        g = ctx.graph
        dR, dt = ops.skin_bwd_csr(verts, g.pack, d_warped.contiguous())
        dR, dt = g.pack.to_old_order(dR), g.pack.to_old_order(dt)                                     # packed (Morton) -> reference node order
        ops.arap_bwd(verts, g.nodes_idx, g.ring, R, t, d_arap.contiguous(), dR, dt)
        return None, dR, dt, None, None


def deform_batched(verts, graph, R, t, want_sr=True):
    """Batched DeformationGraph_geod.forward: returns (warped [B,N,3], arap [B], sr [B]; zeros when want_sr is False --
    no caller of the reference ever reads the smoothness term)."""
    if graph.pack is None:
        graph.pack = pack_graph(verts.float().contiguous(), graph.nodes_idx, graph.influence, graph.weights, graph.ring)
    return _Deform.apply(verts, R, t, graph, want_sr)


def deform_from_d9(verts, graph, d9, want_sr=False, packed_order=True):
    """Inference form fused with models/loss.py:1257-1264: d9 [B,K,9] is the Deformer output (t, 6D residual); the
    identity offset, 6D -> R and the node-record packing are one kernel.  packed_order: the rows of d9 follow
    graph.pack.nodes_idx_m (the Deformer was called with fps1 = pack.nodes_idx_m), else graph.nodes_idx.
    Returns (warped, arap, sr or None)."""
    if graph.pack is None:
        graph.pack = pack_graph(verts.float().contiguous(), graph.nodes_idx, graph.influence, graph.weights, graph.ring)
    table = ops.node_table_from_d9(d9, graph.pack.nodes_xyz, node_perm=None if packed_order else graph.pack.node_perm)
    warped = ops.skin_fwd_packed(verts, graph.pack, table)
    arap, sr = ops.arap_fwd_packed(graph.pack, table, want_sr)
    return warped, arap, sr


class DeformationGraph_geod(nn.Module):
    """Per-cloud object with the reference's attributes and methods (lib/deformation_graph_point.py:71)."""

    def __init__(self, radius=0.1, k=3, sampling_strategy="qslim"):
        super().__init__()
        self.radius = radius
        self.k = k
        self.max_neigh_num = 18
        self.sampling_strategy = sampling_strategy
        self.one_ring_neigh = []
        self.nodes_idx = None
        self.weights = None
        self.influence_nodes_idx = []
        self.dists = []
        self._graph = None

    def construct_graph_euclidean(self, vertices=None, geod=None, device=None, start=None):
        """`geod` (the reference's N x N Euclidean matrix) is accepted and ignored: distances are
        evaluated inside the kernels with the exact direct-difference form."""
        if self.k != 3:
            raise NotImplementedError("the reference hard-codes k = 3 influencing nodes")
        v = torch.as_tensor(vertices).float()
        if device is None:
            device = v.device if v.is_cuda else torch.device("cuda")
        v = v.to(device)[None]
        g = build_graphs(v, start=None if start is None else torch.as_tensor([int(start)]))
        self._graph = g
        self._device_verts = None
        self.max_neigh_num = 9
        # reference-visible attributes (numpy int64 where the reference has numpy)
        self.nodes_idx = g.nodes_idx[0].cpu().numpy()
        self.nodes = v[0][g.nodes_idx[0]].cpu()
        self.one_ring_neigh = g.ring[0].cpu().numpy()
        self.influence_nodes_idx = g.influence[0]
        self.dists = g.dists[0]
        self.weights = g.weights[0]
        self.sigma = g.sigma[0].cpu()
        return self

    def forward(self, vertices, opt_d_rotations, opt_d_translations):
        """vertices [N,3], R [1,K,3,3], t [1,K,3] -> ([1,N,3], arap, sr) (lib/deformation_graph_point.py:233-261)."""
        if self._graph is None:
            raise RuntimeError("construct_graph_euclidean must be called first")
        warped, arap, sr = deform_batched(vertices[None], self._graph, opt_d_rotations, opt_d_translations)
        return warped, arap[0], sr[0]


class DeformationGraph(DeformationGraph_geod):
    """Axis-angle variant (lib/deformation_graph.py:17-116): rotations arrive as [1,K,3] axis-angle."""

    def forward(self, vertices, opt_d_rotations, opt_d_translations):
        R = batch_rodrigues(opt_d_rotations[0]).unsqueeze(0)
        return super().forward(vertices, R, opt_d_translations)


def batch_rodrigues(axisang):
    """lib/utils.py:70-112 (angle = ||a + 1e-8||, quaternion, normalised, 3x3); small per-node torch math."""
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


def deformation_graph_node(verts1):
    """Driver of models/loss.py:1325-1337 / deform.py:41-53: returns (num_nodes_all [B,K] float tensor
    like the reference -- callers apply .long() --, list of per-cloud graph objects)."""
    B, N, _ = verts1.shape
    g = build_graphs(verts1)
    dg_list = []
    for i in range(B):
        dg = DeformationGraph_geod()
        dg._graph = g.select(i)
        dg.max_neigh_num = 9
        dg.nodes_idx = _LazyNumpy(g.nodes_idx[i])
        dg.one_ring_neigh = _LazyNumpy(g.ring[i])
        dg.influence_nodes_idx = g.influence[i]
        dg.dists = g.dists[i]
        dg.weights = g.weights[i]
        dg.sigma = g.sigma[i]
        dg_list.append(dg)
    return g.nodes_idx.to(verts1.dtype if verts1.dtype.is_floating_point else torch.float32), dg_list


def deformation_graph_node_list(verts1):
    """Same as `deformation_graph_node`; the name the entry scripts' local copies are rebound to (deform.py:41-53)."""
    return deformation_graph_node(verts1)


class _LazyNumpy:
    """Device tensor that turns into the reference's numpy array only if somebody asks (no sync otherwise)."""

    def __init__(self, t):
        self._t = t
        self.shape = tuple(t.shape)

    def __array__(self, dtype=None, copy=None):
        a = self._t.cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, i):
        return np.asarray(self)[i]
