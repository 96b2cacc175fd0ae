"""Training-step benchmark (BASELINE.json config 4; reference call stack train.py:93-112, models/loss.py:1349-1435).

    python bench.py --mode train --gpus N --steps K --warmup W            (torchrun for N > 1, one rank per GPU)

One step per rank = a batch of `--pairs` (default 2, config/scape_r.yaml) synthetic pairs at N = M = 4995:
    features = stand-in feature head (LG-Net itself is out of scope, SURVEY 8f1: its parameter COUNT -- 1,822,592 fp32 -- is
               reproduced so that the gradient all-reduce moves the 8.49 MB the real model would)
    loss     = dv_matcher_b200.losses.GraphDeformLoss_Neural(.forward) with the shipped weights (config/scape_r.yaml:35-50),
               graphs cached per shape, soft maps on the tcgen05 pass (DVM_TRAIN_PREC, default f16) or the fp32 pass
    backward = every CUDA backward of the hot path (dvm_softmap_bwd, dvm_sparse_transfer_bwd, dvm_gather_conv_bwd,
               dvm_rot6d_bwd, dvm_skin_bwd_csr, dvm_arap_bwd, dvm_chamfer_bwd) + torch autograd for the MLP / feature head
    exchange = dv_matcher_b200.distributed.allreduce_gradients: ONE flattened NCCL all-reduce (the path's only collective)
    update   = Adam (train.py:64-68)
Timed with CUDA events, W warm-up + K timed steps between barriers, max over ranks; the phases are bracketed separately.
Before timing, one step with IDENTICAL data on every rank checks that the all-reduced gradient equals the local one
(the 1-GPU vs N-GPU gradient equality of SURVEY section 4).
"""
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_PTS = 4995
C = 128
LGNET_PARAMS = 1822592          # Uni3FC parameter count


class FeatureHead(nn.Module):
    """Stand-in for LG-Net's last layers: per-point MLP 128 -> 128 -> 128 on fixed base features, padded with an (unused in
    the forward, but gradient-carrying through a zero-weight sum) parameter block up to LG-Net's parameter count."""

    def __init__(self):
        super().__init__()
        self.l1 = nn.Linear(C, C)
        self.l2 = nn.Linear(C, C)
        used = sum(p.numel() for p in (*self.l1.parameters(), *self.l2.parameters()))
        self.rest = nn.Parameter(torch.zeros(LGNET_PARAMS - used))

    def forward(self, base):
        h = torch.nn.functional.leaky_relu(self.l1(base), 0.2)
        return base + 0.1 * self.l2(h) + 0.0 * self.rest.sum()


def run(args):
    import torch.distributed as dist_mod
    from dv_matcher_b200 import _lib, distributed as dd, synthetic
    from dv_matcher_b200.deformer import Deformer
    from dv_matcher_b200.losses import GraphDeformLoss_Neural
    from bench import ClockSampler, measured_peaks, physical_gpu_index, METRIC, UNIT
    rank, world, device = dd.init()
    dist = dist_mod if world > 1 else None
    lib = _lib.load()
    B = args.pairs
    n = N_PTS if args.n == 50000 else args.n
    prec = os.environ.get("DVM_TRAIN_PREC", "f16")
    os.environ["DVM_TRAIN_PREC"] = prec
    torch.manual_seed(0)
    head = FeatureHead().to(device)
    deformer = Deformer(10).to(device)
    params = list(head.parameters()) + list(deformer.parameters())
    opt = torch.optim.Adam(params, lr=2e-3, betas=(0.9, 0.99))
    crit = GraphDeformLoss_Neural(k_deform=10, w_dist=0.02, w_map=0.005, k_dist=500, N_dist=1000, partial=False, w_deform=0.5,
                                  w_img=0, w_rank=0, w_self_rec=0.5, w_cd=0.1, w_arap=0.01, save_name="bench")
    crit.cache_graphs = True

    def make(q, same=False):
        d = synthetic.make_batch(B, n, n, first_pair=(0 if same else rank * 64) + q * B, regime=args.regime)
        g = {k: v.to(device) for k, v in d.items()}
        # stand-in geodesic matrices (the dataset's heat-method matrices, models/dataset.py:49-54): Euclidean, resident fp32
        g["dist1"] = torch.cdist(g["xyz1"], g["xyz1"])
        g["dist2"] = torch.cdist(g["xyz2"], g["xyz2"])
        return g

    nring = 4
    ring = [make(q) for q in range(nring)]
    use_graph = not args.no_cuda_graph
    bucket = [None]
    ev = {k: [] for k in ("fwd", "bwd", "ar", "opt")}
    captured = None
    if use_graph:
        # forward + backward as ONE CUDA graph (dv_matcher_b200.training): batch, deformation graphs and the dist-loss query
        # indices are refilled into static buffers before every replay; the graphs of the ring's shapes are built once
        # (what crit.cache_graphs does on the eager path)
        from dv_matcher_b200 import training
        from dv_matcher_b200.deformation_graph import build_graphs, draw_fps_start
        ring_graphs = [tuple(build_graphs(g[k], draw_fps_start(B, n)) for k in ("xyz1", "xyz2")) for g in ring]
        captured = training.CapturedTrainStep(crit, lambda a, b: (head(a), head(b)), deformer, params, args.alpha)
        ev = {k: [] for k in ("fwdbwd", "ar", "opt")}

    def batch_of(g):
        return {k: g[k] for k in ("feat1", "feat2", "dist1", "dist2", "xyz1", "xyz2")}

    def step(i, data=None, record=False):
        g = data or ring[i % nring]
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        if record:
            marks[0].record()
        if captured is not None:
            graphs = same_graphs if data is not None else ring_graphs[i % nring]
            out = captured(batch_of(g), graphs)
            if record:
                marks[2].record()
        else:
            crit.graph_keys = (("s1", rank if data is None else -1, i % nring), ("s2", rank if data is None else -1, i % nring))
            f1, f2 = head(g["feat1"]), head(g["feat2"])
            out = crit(f1, f2, g["dist1"], g["dist2"], g["xyz1"], g["xyz2"], args.alpha, deformer)
            if record:
                marks[1].record()
            opt.zero_grad(set_to_none=False)
            out[0].backward()
            if record:
                marks[2].record()
        bucket[0] = dd.allreduce_gradients(params, world=world, bucket=bucket[0])
        if record:
            marks[3].record()
        opt.step()
        if record:
            marks[4].record()
            if captured is not None:
                ev["fwdbwd"].append((marks[0], marks[2]))
            else:
                ev["fwd"].append((marks[0], marks[1])); ev["bwd"].append((marks[1], marks[2]))
            ev["ar"].append((marks[2], marks[3])); ev["opt"].append((marks[3], marks[4]))
        return out

    # ---- gradient equality: identical data on every rank => the averaged gradient is the local gradient
    same = make(0, same=True)
    torch.manual_seed(1)
    import random
    random.seed(1)
    if captured is not None:
        same_graphs = tuple(build_graphs(same[k], torch.zeros(B, dtype=torch.int64, device=device)) for k in ("xyz1", "xyz2"))
        captured(batch_of(same), same_graphs)
    else:
        opt.zero_grad(set_to_none=False)
        f1, f2 = head(same["feat1"]), head(same["feat2"])
        crit.graph_keys = ("eq1", "eq2")
        out = crit(f1, f2, same["dist1"], same["dist2"], same["xyz1"], same["xyz2"], args.alpha, deformer)
        out[0].backward()
        del out, f1, f2
    local = torch.cat([p.grad.reshape(-1) for p in params]).clone()
    dd.allreduce_gradients(params, world=world)
    after = torch.cat([p.grad.reshape(-1) for p in params])
    grad_eq = float((after - local).abs().max() / local.abs().max().clamp_min(1e-30))
    # run-to-run differences of the 16-bit softmax mass (queue order) reach the gradients at ~1e-6; ranks see the same data
    assert grad_eq <= 1e-3, f"all-reduced gradient differs from the local one by {grad_eq}"
    if captured is None:
        opt.zero_grad(set_to_none=False)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    sampler = ClockSampler(physical_gpu_index(int(os.environ.get("LOCAL_RANK", "0"))))
    sampler.start()
    l0 = lib.dvm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i, record=True)
    e1.record()
    torch.cuda.synchronize(device)
    launches = lib.dvm_launch_count() - l0
    if captured is not None:
        launches += captured.launches_per_step * args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = e0.elapsed_time(e1)
    rank_ms = [ms]
    if dist is not None:
        t = torch.tensor([ms], device=device)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        rank_ms = [a.item() for a in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    phases = {k: sum(a.elapsed_time(b) for a, b in v) / max(1, len(v)) for k, v in ev.items()}
    n_params = sum(p.numel() for p in params)
    if rank == 0:
        peaks = measured_peaks()
        line = dict(
            metric=METRIC.replace("match+deform", "training step: loss fwd+bwd + gradient all-reduce + Adam"), value=B * world * args.steps / (ms * 1e-3),
            unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype=f"{prec} soft-map forward + backward on tcgen05 (fp32 accumulate, exact fp32 top-k terms)" if prec != "fp32" else "f32", data="synthetic",
            config=dict(workload=f"config 4 training step: GraphDeformLoss_Neural fwd+bwd, N=M={n}, C={C}, alpha={args.alpha}, B={B} pairs/GPU, "
                                 f"stand-in feature head with LG-Net's parameter count, Deformer, Adam",
                        pairs_per_step_per_gpu=B, parallelism=f"data parallel over {world} GPU(s): one flattened NCCL all-reduce of {n_params * 4 / 1e6:.2f} MB per step",
                        graphs="warm (cached per shape)", prec=prec,
                        launch="forward + backward replayed as one CUDA graph (static buffers refilled per step); all-reduce and Adam eager" if captured is not None else "eager"),
            phases_ms=(dict(forward_backward_one_cuda_graph=phases["fwdbwd"], allreduce=phases["ar"], optimizer=phases["opt"]) if captured is not None
                       else dict(forward=phases["fwd"], backward=phases["bwd"], allreduce=phases["ar"], optimizer=phases["opt"])),
            allreduce=dict(bytes=n_params * 4, ms=phases["ar"], backend="nccl" if world > 1 else "none (1 GPU)",
                           gradient_equality_rel_err=grad_eq),
            gpu_launches=int(launches), clocks=sampler.summary(),
            extra=dict(ms_per_step_by_rank=[round(v / args.steps, 3) for v in rank_ms], n_params=n_params,
                       reference_cpu="5.02 s forward + 6.23 s backward for the loss alone at B=2 on 8 cores (BASELINE.md section 2)"))
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
