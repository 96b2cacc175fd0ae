"""The "match + deform" unit of work on a batch of shape pairs, as one call.

match : Pi_12, Pi_21 (top-10 idx + w), verts12 = Pi_12 @ verts2, verts21, hard maps T12, T21
        -> ONE batched dvm_softmap_fwd over the 2B (source, target) problems (both directions).
deform: xyz 10-NN of every cloud, Deformer (torch MLP; gathers fused), 6D -> R, skinning + ARAP,
        Chamfer(deformed, target) and Chamfer(verts12, verts2), both directions
        -> the orchestration of GraphDeformLoss_Neural.deform (models/loss.py:1228-1282) / deform.py:229-257
           without the prints, OFF dumps and Python per-batch-element loops.
Graphs are built once per shape (`build_graphs`) and passed in ("warm"); `build_graphs` timed alone is
the "cold" cost the reference pays on the CPU every step (models/loss.py:1401-1402).
"""
import torch

from . import maps, ops
from .deformation_graph import BatchedGraph, build_graphs, deform_batched
from .geometry import rotation_6d_to_matrix

_IDEN6 = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0)
_IDEN6_DEV = {}


def _iden6(device):
    """The identity offset of models/loss.py:1259-1262 as a cached device constant: building it per step from a Python
    list is a blocking host-to-device copy, i.e. a full stream synchronisation in the middle of every step."""
    t = _IDEN6_DEV.get(device)
    if t is None:
        t = torch.tensor(_IDEN6, device=device, dtype=torch.float32)
        _IDEN6_DEV[device] = t
    return t


def cat_graphs(g1, g2):
    return BatchedGraph(*[torch.cat([a, b]) for a, b in zip(
        (g1.nodes_idx, g1.influence, g1.dists, g1.weights, g1.ring, g1.sigma),
        (g2.nodes_idx, g2.influence, g2.dists, g2.weights, g2.ring, g2.sigma))])


def match(feat1, feat2, verts1, verts2, alpha=100.0, prec=None):
    """Both directions of the fused soft/hard map for B pairs with N == M (one launch sequence).

    Returns (SparseSoftMap over the 2B stacked problems [12-direction first], verts_transferred [2B,N,3])."""
    if feat1.shape[1] == feat2.shape[1]:
        X = torch.cat([feat1, feat2])
        Y = torch.cat([feat2, feat1])
        V = torch.cat([verts2, verts1])
        return maps.soft_map(X, Y, alpha, v=V, prec=prec)
    sm12, v12 = maps.soft_map(feat1, feat2, alpha, v=verts2, prec=prec)
    sm21, v21 = maps.soft_map(feat2, feat1, alpha, v=verts1, prec=prec)
    return (sm12, sm21), (v12, v21)


def match_deform(feat1, feat2, verts1, verts2, graphs, deformer, alpha=100.0, k_deform=10, prec=None):
    """One pass of the hot path over B pairs (N == M).  `graphs` = BatchedGraph over cat([verts1, verts2]).

    Returns a dict of device tensors; nothing is synchronised or copied to the host here."""
    B, N, _ = verts1.shape
    src = torch.cat([verts1, verts2])                                      # source cloud of each of the 2B problems
    tgt = torch.cat([verts2, verts1])
    fsrc = torch.cat([feat1, feat2])
    ftgt = torch.cat([feat2, feat1])
    sm, vt = maps.soft_map(fsrc, ftgt, alpha, v=tgt, prec=prec)            # [2B,...]: rows 0..B-1 = 1->2, B..2B-1 = 2->1
    idx_self = ops.knn3(src, src, k_deform)                                # idx11 | idx22  (models/loss.py:1229-1230)
    idx_tgt = torch.cat([idx_self[B:], idx_self[:B]])
    fps = graphs.nodes_idx
    deformations = deformer.forward_fused(fsrc, ftgt, idx_self, idx_tgt, src, vt, sm, fps)     # [2B,K,9]
    iden = _iden6(src.device)
    R = rotation_6d_to_matrix(deformations[..., 3:] + iden)               # models/loss.py:1258-1264
    T = deformations[..., :3].contiguous()
    deformed, arap, sr = deform_batched(src, graphs, R, T)                 # models/loss.py:1269-1273
    cd_d1, cd_d2, _, _ = ops.chamfer_fwd(deformed, tgt)                    # chamfer(deformed, target)   :1279
    cd_s1, cd_s2, _, _ = ops.chamfer_fwd(vt, tgt)                          # chamfer(verts12, verts2)    :1280
    return dict(T=sm.argmin, top_idx=sm.idx, top_w=sm.w, verts_t=vt, deformed=deformed, arap=arap,
                cd_deform=cd_d1.mean(1) + cd_d2.mean(1), cd_self=cd_s1.mean(1) + cd_s2.mean(1))


class MatchDeformEngine:
    """Public end-to-end entry: host (pinned) buffers in, host results out, on the current device.

    step(feat1, feat2, verts1, verts2) copies the step's inputs H2D, runs match_deform and reads the step's
    results back (hard maps T12/T21 int64 [2B,N] and the per-problem losses) -- the call bench.py's `e2e`
    number times.  Staging buffers are double-buffered: `prefetch(...)` (or step(..., next_inputs=...)) starts the
    H2D copy of the NEXT step's inputs on a copy stream while the current step computes, so the PCIe transfer of
    step i+1 overlaps the kernels of step i; every step's inputs are still copied inside that step's call sequence.
    """

    def __init__(self, deformer, alpha=100.0, k_deform=10, prec=None, device=None):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.deformer = deformer.to(self.device).eval()
        self.alpha, self.k_deform, self.prec = alpha, k_deform, prec
        self._dev = [{}, {}]            # two staging sets
        self._slot = 0
        self._pending = None            # (slot, event, host tensors identity) of a prefetched batch
        self._copy_stream = torch.cuda.Stream(self.device)
        self._host_out = {}
        self._graph_cache = {}

    def _stage(self, slot, name, host, stream):
        buf = self._dev[slot].get(name)
        if buf is None or buf.shape != host.shape or buf.dtype != host.dtype:
            buf = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            self._dev[slot][name] = buf
        with torch.cuda.stream(stream):
            buf.copy_(host, non_blocking=True)
        return buf

    def _copy_in(self, slot, inputs, stream):
        names = ("f1", "f2", "v1", "v2")
        bufs = [self._stage(slot, n, h, stream) for n, h in zip(names, inputs)]
        ev = torch.cuda.Event()
        ev.record(stream)
        return bufs, ev

    def prefetch(self, feat1, feat2, verts1, verts2):
        """Start the H2D copy of the next step's inputs (pinned host tensors) on the copy stream."""
        slot = 1 - self._slot
        # the staging set may still be read by kernels of the step before last: order the copy after the compute stream
        self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        bufs, ev = self._copy_in(slot, (feat1, feat2, verts1, verts2), self._copy_stream)
        self._pending = (slot, ev, bufs, tuple(id(t) for t in (feat1, feat2, verts1, verts2)))

    def graphs_for(self, key, verts_cat, start=None):
        """Per-shape graph cache ("warm" path). key identifies the batch of shapes."""
        g = self._graph_cache.get(key)
        if g is None:
            g = build_graphs(verts_cat, start)
            self._graph_cache[key] = g
        return g

    @torch.no_grad()
    def step(self, feat1, feat2, verts1, verts2, graph_key="default", fps_start=None, next_inputs=None):
        cur = torch.cuda.current_stream(self.device)
        ids = tuple(id(t) for t in (feat1, feat2, verts1, verts2))
        if self._pending is not None and self._pending[3] == ids:           # inputs already on their way
            slot, ev, bufs, _ = self._pending
            cur.wait_event(ev)
        else:
            slot = 1 - self._slot
            bufs, ev = self._copy_in(slot, (feat1, feat2, verts1, verts2), cur)
        self._pending = None
        self._slot = slot
        f1, f2, v1, v2 = bufs
        graphs = self.graphs_for(graph_key, torch.cat([v1, v2]), fps_start)
        if next_inputs is not None:
            self.prefetch(*next_inputs)                                     # overlaps with the kernels launched below
        out = match_deform(f1, f2, v1, v2, graphs, self.deformer, self.alpha, self.k_deform, self.prec)
        res = {}
        for name in ("T", "cd_deform", "cd_self", "arap"):
            t = out[name]
            h = self._host_out.get(name)
            if h is None or h.shape != t.shape:
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                self._host_out[name] = h
            h.copy_(t, non_blocking=True)
            res[name] = h
        cur.synchronize()
        return res

    @staticmethod
    def h2d_bytes(feat1, feat2, verts1, verts2):
        return sum(t.numel() * t.element_size() for t in (feat1, feat2, verts1, verts2))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._host_out.values())
