"""The training step of the hot path as ONE CUDA graph (reference loop: train.py:93-112).

A step of `GraphDeformLoss_Neural` is ~600 kernel launches for ~6 ms of GPU work at the reference's training size (B = 2,
N = 4995): issued eagerly it is host-bound (`tools/prof_train.py`).  `CapturedTrainStep` captures feature head -> loss ->
backward once and replays it; everything a step reads lives in static device buffers that are refilled before the replay:

  * the batch (features / base inputs of the feature function, geodesic matrices, vertices),
  * the two deformation graphs (`BatchedGraph.tensors()`; built by the caller, per shape, as the graph cache does),
  * the query indices of the dist loss -- drawn on the HOST with `random.sample`, in the reference's order
    (models/loss.py:1361-1364), staged through pinned memory.

Gradients land in the parameters' `.grad` (static as well: never zero them with set_to_none=True after the capture); the gradient
all-reduce (`distributed.allreduce_gradients`) and the optimizer step stay outside the graph.
"""
import random

import torch

from .deformation_graph import BatchedGraph, build_graphs, draw_fps_start


class CapturedTrainStep:
    def __init__(self, crit, feature_fn, deformer, params, alpha, warmup=2):
        """crit: losses.GraphDeformLoss_Neural(_Partial); feature_fn(x1, x2) -> (feat1, feat2) (the network);
        params: the parameters that receive gradients; alpha: the soft-map temperature of the step."""
        self.crit, self.feature_fn, self.deformer, self.params = crit, feature_fn, deformer, list(params)
        self.alpha, self.warmup = alpha, warmup
        self.graph = None
        self.static = None
        self.out = None
        self.launches_per_step = 0

    # ---- static buffers -------------------------------------------------------------------------------------
    def _bind(self, batch, graphs):
        dev = batch["xyz1"].device
        self.static = {k: torch.empty_like(v) for k, v in batch.items()}
        self.g_static = [BatchedGraph.from_tensors([torch.empty_like(t) for t in g.tensors()]) for g in graphs]
        self.nodes_f = [torch.empty(g.nodes_idx.shape, dtype=torch.float32, device=dev) for g in graphs]
        n_dist = self.crit.N_dist
        self.draws = (torch.empty(n_dist, dtype=torch.int64, device=dev), torch.empty(n_dist, dtype=torch.int64, device=dev))
        self.draws_pinned = torch.empty(2, n_dist, dtype=torch.int64).pin_memory()
        self.draws_done = torch.cuda.Event()
        self.draws_done.record()
        self.crit.static_graphs = [(self.nodes_f[0], self.g_static[0]), (self.nodes_f[1], self.g_static[1])]
        self.crit.static_draws = self.draws

    def _fill(self, batch, graphs):
        for k, v in batch.items():
            self.static[k].copy_(v, non_blocking=True)
        for gs, nf, g in zip(self.g_static, self.nodes_f, graphs):
            for dst, src in zip(gs.tensors(), g.tensors()):
                dst.copy_(src, non_blocking=True)
            nf.copy_(g.nodes_idx)
        if self.crit.w_dist > 0:
            self.draws_done.synchronize()                      # the previous step's copy has left the pinned buffer
            n1, n2 = self.static["dist1"].shape[1], self.static["dist2"].shape[1]
            self.draws_pinned[0] = torch.tensor(random.sample(range(n1), self.crit.N_dist))     # the reference's draw order
            self.draws_pinned[1] = torch.tensor(random.sample(range(n2), self.crit.N_dist))
            self.draws[0].copy_(self.draws_pinned[0], non_blocking=True)
            self.draws[1].copy_(self.draws_pinned[1], non_blocking=True)
            self.draws_done.record()

    def _step(self):
        s = self.static
        f1, f2 = self.feature_fn(s["feat1"], s["feat2"])
        out = self.crit(f1, f2, s["dist1"], s["dist2"], s["xyz1"], s["xyz2"], self.alpha, self.deformer)
        out[0].backward()
        return tuple(o.detach() if torch.is_tensor(o) else o for o in out)

    # ---- one step ---------------------------------------------------------------------------------------------
    def __call__(self, batch, graphs=None):
        """batch: dict feat1, feat2 (inputs of feature_fn), dist1, dist2, xyz1, xyz2 (device tensors of fixed shapes);
        graphs: (BatchedGraph of xyz1, of xyz2) or None (built here, with a freshly drawn FPS start like the reference).
        Returns the loss 5-tuple (static tensors: read or copy them before the next call)."""
        from . import _lib
        if graphs is None:
            graphs = tuple(build_graphs(batch[k].detach(), draw_fps_start(batch[k].shape[0], batch[k].shape[1])) for k in ("xyz1", "xyz2"))
        if self.static is None:
            self._bind(batch, graphs)
        self._fill(batch, graphs)
        if self.graph is None:
            cur = torch.cuda.current_stream()
            side = self._capture_stream = torch.cuda.Stream()       # kept alive: the library's workspaces are keyed by stream, and the
            side.wait_stream(cur)                                   # captured kernels hold raw pointers into this stream's buffers
            with torch.cuda.stream(side):                      # warm-up on the capture stream: workspaces, lazy initialisation
                for _ in range(self.warmup):
                    for p in self.params:
                        p.grad = None
                    self._step()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            for p in self.params:
                p.grad = None
            _lib.workspace.keep_retired = True
            self.graph = torch.cuda.CUDAGraph()
            l0 = _lib.load().dvm_launch_count()
            with torch.cuda.graph(self.graph, stream=side):
                self.out = self._step()
            self.launches_per_step = int(_lib.load().dvm_launch_count() - l0)     # library kernels one replay launches
        self.graph.replay()
        return self.out


__all__ = ["CapturedTrainStep"]
