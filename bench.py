#!/usr/bin/env python
"""Benchmark of the match + deform hot path (BASELINE.json metric: shape pairs/s, sim TFLOP/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (B200, one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU op sequence (oracle port)
    python bench.py --mode train --gpus N ...                 # config 4: training step (loss fwd + bwd + NCCL all-reduce)

One "step" = one pass of match + deform over a batch of `--pairs` synthetic pairs per GPU at
N = M = `--n` points (default 50000), C = 128.  Pairs are independent: ranks take disjoint pairs, there is
no data-path collective (torch.distributed is only used for the barrier and the max-over-ranks time).
Prints ONE JSON line (rank 0).

What is timed how:
  value          K replays of the warm step captured as ONE CUDA graph per input batch (inputs resident in HBM); consecutive batches
                 alternate over `--streams` compute streams (own workspaces and graph pools), so the issue-bound tail of one step
                 overlaps the tensor-bound sweep of the next; every step's work is inside the timed region (K steps between two
                 events bracketed by device synchronisation), ms_per_step = that time / K;
  roofline*      the same K steps launched eagerly with CUDA-event brackets inside the library, on the launching stream:
                 `roofline` = the dominant kernel (priming + sweep of the similarity pass), `roofline_fused` = the whole
                 fused op dvm_softmap_fwd (prep + prime + sweep + finalize + rescue), both against the measured bf16 peak;
  roofline_hbm   the memory-bound kernels alone at N = 200k x B = 32 (inputs > L2), algorithmic bytes / CUDA-event time;
  e2e            MatchDeformEngine.submit/result with pinned HOST buffers, up to four steps in flight: H2D of the inputs and D2H of the step's results
                 (hard maps, soft map idx + w, transferred and deformed coordinates, losses) inside the timed region;
  cpu_baseline / torch_cuda_baseline   the reference's torch op sequence (oracle port) on the host cores / on the same B200.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "shape pairs/sec (match+deform)"
UNIT = "pairs/s"
C = 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"])
    ap.add_argument("--n", type=int, default=50000, help="points per cloud (N = M)")
    ap.add_argument("--pairs", type=int, default=2, help="pairs per GPU per step")
    ap.add_argument("--alpha", type=float, default=100.0)
    ap.add_argument("--regime", default="structured", choices=["structured", "unstructured"])
    ap.add_argument("--prec", default="f16", choices=["f16", "bf16", "fp32"])
    ap.add_argument("--cpu-rows", type=int, default=0, help="row-slab size of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-5k", action="store_true")
    ap.add_argument("--no-hbm", action="store_true", help="skip the memory-bound kernel microbench (roofline_hbm)")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="compute streams the ring's batches are spread over (CUDA-graph mode)")
    return ap.parse_args()


def bench_config(args, world):
    """`config` of the JSON line: a pure function of the command line, so both arms (ours and --impl reference) print the same dict."""
    return dict(workload=f"match+deform, N=M={args.n}, C={C}, alpha={args.alpha}, {args.regime} features (config 5 of BASELINE.json)",
                pairs_per_step_per_gpu=args.pairs, parallelism=f"pairs sharded over {world} GPU(s), no data-path collective",
                l2_policy="inputs rotate through >= 2 distinct batches (> 126 MB L2 in total)",
                graphs="warm (built once per shape, outside the timed region)", prec=args.prec,
                launch="eager launches" if args.no_cuda_graph else
                       ("the warm step is one CUDA-graph launch per batch" + (f"; consecutive batches alternate over {args.streams} compute streams "
                        "(the tail of one step overlaps the sweep of the next)" if args.streams > 1 else "")))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_sustained=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], source="measured")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 50 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.power_w, self.mem_mhz, self.temp_c = [], [], []

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {getattr(pynvml, n): n for n in dir(pynvml) if n.startswith("nvmlClocksThrottleReason") and not n.endswith("All")}
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:                                   # diagnostics for a slow rank: board power, memory clock, temperature
                    self.power_w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    self.mem_mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM))
                    self.temp_c.append(pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU))
                except Exception:
                    pass
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if bit and (r & bit):
                        self.reasons.add(n.replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.05)
        except Exception as e:  # clocks are evidence, not a dependency
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        bad = {"HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown"}
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting")),
                    rejected=bool(bad & self.reasons))


def _median(v):
    v = sorted(v)
    return v[len(v) // 2] if v else None


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline / torch_cuda_baseline: the oracle port of the reference's torch op sequence
# ------------------------------------------------------------------------------------------------------
def _deformer_params(device="cpu"):
    from dv_matcher_b200.deformer import Deformer
    torch.manual_seed(0)
    return {k: v.detach().to(device) for k, v in Deformer(10).state_dict().items()}


def reference_pairs_per_s(n, alpha, regime, rows, steps=1, warmup=0, device="cpu", results=None):
    """(pairs/s, cores, sample description, estimated s/pair, measured wall s of one sample) of the oracle port."""
    from dv_matcher_b200 import synthetic
    from oracle import pipeline as opipe
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = rows < 0 or n <= 5000
    if full:
        rows = n
    elif rows == 0:
        rows = max(128, min(n, int(2048 * 5000 / max(n, 5000)) * 4))
    batch = synthetic.make_batch(1, n, n, first_pair=0, regime=regime)
    params = _deformer_params(device)
    graph = opipe.make_cpu_graph(batch["xyz1"][0])
    sync = None
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
        graph = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in graph.items()}
        sync = torch.cuda.synchronize
    for _ in range(warmup):
        opipe.time_pair_sample(batch, graph, params, alpha, rows, sync=sync)
    res = results if results is not None else {}
    ts = [opipe.time_pair_sample(batch, graph, params, alpha, rows, sync=sync, results=res) for _ in range(max(1, steps))]
    sec = min(ts)
    where = "host cores" if device == "cpu" else "the same B200 (stock PyTorch kernels)"
    if rows >= n:
        sample = (f"oracle port of the reference's torch op sequence on {where}: the FULL sequence for one pair of N=M={n} "
                  f"(both directions, no extrapolation), graphs warm; best of {max(1, steps)}")
    else:
        sample = (f"oracle port of the reference's torch op sequence on {where}: a slab of {rows} of {n} source rows per direction "
                  f"(x N/rows -- the reference cannot materialise {n} x {n}), per-cloud parts in full, graphs warm; best of {max(1, steps)}")
    return 1.0 / sec, cores, sample, sec, res.get("sample_s", sec)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = min(args.warmup, 1)
    v, cores, sample, sec, wall = reference_pairs_per_s(args.n, args.alpha, args.regime, args.cpu_rows, steps=args.steps, warmup=warm)
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=wall * 1e3,                 # wall time of ONE bounded sample step (what actually ran)
                ms_per_pair_estimated=sec * 1e3,        # the sample scaled to the whole workload (slab x N/rows)
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=bench_config(args, args.gpus),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    if args.n > 5000 and not args.no_5k:
        v5, _, s5, sec5, _ = reference_pairs_per_s(4995, args.alpha, args.regime, -1, steps=1, warmup=0)
        line["also_5k"] = dict(workload="match+deform, N=M=4995, one pair, full un-extrapolated sequence (configs 1/3/4 scale)",
                               value=v5, unit=UNIT, ms_per_pair=sec5 * 1e3, sample=s5)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def timed_loop(fn, steps, warmup, dist, device, finish=None):
    """W warm-up calls, barrier + synchronize, EXACTLY `steps` calls between two CUDA events on the current stream,
    synchronize; max over ranks.  `finish` runs after the last call, inside the timed region (pipelined e2e drains there)."""
    for i in range(warmup):
        fn(i)
    if finish is not None and warmup:
        finish()
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    if finish is not None:
        finish()
    torch.cuda.synchronize(device)      # e2e work runs on the engine's own streams: the closing event must come after all of it
    e1.record()
    torch.cuda.synchronize(device)
    ms = max(e0.elapsed_time(e1), 0.0)
    wall_ms = (time.perf_counter() - t0) * 1e3
    timed_loop.wall_ms = wall_ms
    timed_loop.rank_ms = [ms]
    if dist is not None:
        t = torch.tensor([ms], device=device)
        allr = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(allr, t)
        timed_loop.rank_ms = [a.item() for a in allr]          # diagnostic: the spread over ranks (the value uses the max)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        dist.barrier()
    return ms


def read_profile(lib, ch):
    import ctypes
    tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
    lib.dvm_profile_read_channel(ch, ctypes.byref(tot), ctypes.byref(cnt))
    return tot.value, cnt.value


def hbm_microbench(device, peaks, npts=200000, nb=32, iters=10):
    """The memory-bound kernels alone, at a size where the HBM fraction means something: N = 200k points,
    B = 32 clouds (every operand set > 126 MB L2), algorithmic bytes / CUDA-event time on the launching stream."""
    from dv_matcher_b200 import ops, synthetic
    from dv_matcher_b200.deformation_graph import build_graphs, BatchedGraph
    gen = torch.Generator().manual_seed(99)
    base = 2                                                # distinct clouds built for real (FPS at 200k is ~0.4 s per cloud) ...
    clouds = torch.stack([synthetic.ellipsoid_cloud(npts, gen) for _ in range(base)]).to(device)
    g0 = build_graphs(clouds, torch.arange(base))
    rep = nb // base                                        # ... then replicated into separate memory: same access pattern per cloud
    graphs = BatchedGraph.from_tensors([t.repeat(rep, *([1] * (t.dim() - 1))).contiguous() for t in g0.tensors()])
    verts = clouds.repeat(rep, 1, 1).contiguous()
    K = npts // 2
    d9 = (0.05 * torch.randn(nb, K, 9, device=device)).contiguous()
    table, R, t = ops.node_table_from_d9(d9, graphs.pack.nodes_xyz, want_rt=True)
    d6 = (d9[..., 3:] + torch.tensor([1.0, 0, 0, 0, 1, 0], device=device)).contiguous()
    # soft-map-like sparse operands: 10 random-but-local columns per row
    nbs = 8
    idx = (torch.arange(npts, device=device)[None, :, None] + torch.randint(-2000, 2000, (nbs, npts, 10), device=device)).clamp_(0, npts - 1).int().contiguous()
    w = torch.rand(nbs, npts, 10, device=device)
    y3 = torch.randn(nbs, npts, 3, device=device)
    y128 = torch.randn(nbs, npts, 128, device=device)
    knn = ops.knn3(verts[:nbs], verts[:nbs], 10)
    conv_w = torch.randn(10, device=device)
    conv_b = torch.zeros(1, device=device)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / iters

    nv, nn = nb * npts, nb * K
    cases = [
        ("skin_fwd_packed_kernel", lambda: ops.skin_fwd_packed(verts, graphs.pack, table), nv * 78.0,
         "78 B/vertex: v 12 + 3 idx 12 + 3 w 12 + out 12 + node table 60 B/node (SURVEY 8d; this layout moves 52 + 64 B/node = 84)"),
        ("arap_fwd_packed_kernel<0>", lambda: ops.arap_fwd_packed(graphs.pack, table, want_sr=False), nn * 96.0,
         "96 B/node: g 12 + t 12 + R 36 + ring 36 (SURVEY 8d); neighbour records are L2 gathers"),
        ("arap_fwd_packed_kernel<1> (+ smoothness)", lambda: ops.arap_fwd_packed(graphs.pack, table, want_sr=True), nn * 96.0, "same bytes; reads both sectors of every neighbour record"),
        ("node_table_kernel<1> (identity offset + 6D->R + pack)", lambda: ops.node_table_from_d9(d9, graphs.pack.nodes_xyz), nn * (36.0 + 12.0 + 64.0),
         "d9 36 + g 12 in, 64-byte record out per node (replaces rot6d 24 + 36, the offset add 24 + 24 and the t copy 12 + 12)"),
        ("rot6d_fwd_kernel", lambda: ops.rot6d_fwd(d6), nn * 60.0, "24 B in + 36 B out per node (SURVEY 8d)"),
        ("sparse_transfer_fwd_kernel D=3", lambda: ops.sparse_transfer_fwd(idx, w, y3), nbs * npts * (40.0 + 40.0 + 12.0 + 12.0),
         "idx 40 + w 40 + out 12 per row + Y read once 12 B/column"),
        ("sparse_transfer_fwd_kernel D=128", lambda: ops.sparse_transfer_fwd(idx, w, y128), nbs * npts * (40.0 + 40.0 + 512.0 + 512.0),
         "idx 40 + w 40 + out 512 per row + Y read once 512 B/column; the 10 gathered rows per output row (5 KB) come from L2"),
        ("gather_conv_fwd_kernel", lambda: ops.gather_conv_fwd(y128, knn, conv_w, conv_b), nbs * npts * (80.0 + 512.0 + 512.0),
         "idx 80 (int64, the reference's dtype) + out 512 per row + feat read once 512 B/point; 10 gathered rows per output row come from L2"),
    ]
    out = []
    for name, fn, nbytes, note in cases:
        ms = timeit(fn)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append(dict(kernel=name, bound="hbm", algorithmic_bytes=nbytes, avg_launch_ms=ms, achieved=gbs, peak=peaks["hbm"], unit="GB/s",
                        frac=gbs / peaks["hbm"], bytes_model=note))
    return dict(workload=f"N={npts} points x B={nb} clouds (sparse transfers / gather-conv: B={nbs}), K=N/2 nodes, inputs larger than L2; CUDA events, {iters} launches",
                peak_source=peaks["source"], kernels=out)


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    from dv_matcher_b200 import _lib, pipeline, synthetic
    from dv_matcher_b200.deformation_graph import build_graphs
    from dv_matcher_b200.deformer import Deformer
    lib = _lib.load()

    peaks = measured_peaks()
    torch.manual_seed(0)
    deformer = Deformer(10).to(device).eval()
    n, B = args.n, args.pairs
    use_graph = not args.no_cuda_graph

    def make_ring(npts, nb, nring):
        """`nring` distinct input batches (pinned host + device copies); rank r owns pairs r*nring*nb ..."""
        ring = []
        for q in range(nring):
            host = synthetic.make_batch(nb, npts, npts, first_pair=(rank * nring + q) * nb, regime=args.regime, pin=True)
            dev = {k: v.to(device) for k, v in host.items()}
            ring.append((host, dev))
        return ring

    def bench_size(npts, nb, steps, warmup, with_e2e):
        bytes_per_batch = nb * 2 * npts * (C + 3) * 4
        nring = max(2, min(8, int(130e6 * 1.5 / bytes_per_batch) + 1))      # ring of inputs > 126 MB L2
        nring = max(nring, 4 if with_e2e else min(args.streams, 4))        # one batch per engine slot (4) / compute stream
        ring = make_ring(npts, nb, nring)
        t0 = time.perf_counter()
        graphs = [build_graphs(torch.cat([d["xyz1"], d["xyz2"]]), torch.arange(2 * nb) % npts) for _, d in ring]
        torch.cuda.synchronize(device)
        graph_cold_s = (time.perf_counter() - t0) / (nring * nb)

        def eager_step(i):
            _, d = ring[i % nring]
            with torch.no_grad():
                return pipeline.match_deform(d["feat1"], d["feat2"], d["xyz1"], d["xyz2"], graphs[i % nring], deformer,
                                             alpha=args.alpha, prec=args.prec)

        # ---- warm-up (also: every lazy one-time initialisation happens here, before any capture)
        # `--streams 2`: the two input batches of the ring live on two streams (own workspaces, own graph pools), so the
        # latency-bound tail of one step (k-NN, finalize, decoder MLP, transfers) overlaps the tensor-bound sweep of the next
        nstreams = max(1, min(args.streams, nring)) if use_graph else 1
        sides = [torch.cuda.Stream(device) for _ in range(nstreams)]
        for sd in sides:
            sd.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(sd):                # a capture stream: warm its workspaces with the eager step
                for i in range(max(warmup, nring)):
                    eager_step(i)
            torch.cuda.current_stream(device).wait_stream(sd)
        for i in range(warmup):
            eager_step(i)
        torch.cuda.synchronize(device)

        # ---- value: the warm step replayed as one CUDA graph per input batch
        cgs, keep = None, []
        if use_graph:
            _lib.workspace.keep_retired = True
            pools = [torch.cuda.graph_pool_handle() for _ in range(nstreams)]
            cgs = []
            for q in range(nring):
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, pool=pools[q % nstreams], stream=sides[q % nstreams]):
                    keep.append(eager_step(q))
                cgs.append(cg)

            if nstreams == 1:
                def step(i):
                    cgs[i % nring].replay()
            else:
                def step(i):
                    q = i % nring
                    with torch.cuda.stream(sides[q % nstreams]):
                        cgs[q].replay()
        else:
            step = eager_step
        l0 = lib.dvm_launch_count()
        for i in range(nring):                        # launches per step (a replayed graph launches what its capture recorded)
            eager_step(i)
        launches_per_step = (lib.dvm_launch_count() - l0) / nring
        sampler = ClockSampler(physical_gpu_index(local))
        sampler.start()
        ms = timed_loop(step, steps, warmup, dist, device)
        rank_ms = [round(v / steps, 3) for v in timed_loop.rank_ms]
        sampler.stop_flag = True
        sampler.join(timeout=2)

        # ---- rooflines: the same steps launched eagerly with the library's CUDA-event brackets
        lib.dvm_profile_enable(1)
        eager_ms = timed_loop(eager_step, steps, 1, dist, device)
        cand_ms, cand_cnt = read_profile(lib, 0)
        fused_ms, fused_cnt = read_profile(lib, 1)
        lib.dvm_profile_enable(0)
        clocks_by_rank = None
        if dist is not None:                       # diagnostic: what every rank's GPU did (a slow rank sets the whole-job value)
            objs = [None] * world
            mine = dict(sampler.summary(), power_w=_median(sampler.power_w), mem_mhz=_median(sampler.mem_mhz), temp_c=_median(sampler.temp_c),
                        sweep_ms=round(cand_ms / max(cand_cnt, 1), 4), softmap_ms=round(fused_ms / max(fused_cnt, 1), 4),
                        gpu=physical_gpu_index(local))
            dist.all_gather_object(objs, mine)
            clocks_by_rank = [{k: o.get(k) for k in ("gpu", "sm_mhz", "reasons", "power_w", "mem_mhz", "temp_c", "sweep_ms", "softmap_ms")} for o in objs]
        # the warm-up step of that loop is bracketed too: per-launch averages are what is used below
        res = dict(ms=ms, eager_ms=eager_ms, launches=int(round(launches_per_step * steps)), nring=nring, graph_cold_s=graph_cold_s,
                   clocks=sampler.summary(), rank_ms=rank_ms, clocks_by_rank=clocks_by_rank,
                   cand_avg_ms=cand_ms / max(cand_cnt, 1), fused_avg_ms=fused_ms / max(fused_cnt, 1),
                   launch_mode="cuda_graph" if use_graph else "eager")
        # stats of one step: rows the 16-bit pass could not certify
        from dv_matcher_b200 import ops
        _, d = ring[0]
        o = ops.softmap_fwd(torch.cat([d["feat1"], d["feat2"]]), torch.cat([d["feat2"], d["feat1"]]), None, alpha=args.alpha,
                            prec=args.prec, want_stats=True)
        st = o.stats.cpu().tolist()
        res["uncertified_rows_frac"] = st[0] / float(2 * nb * npts)
        res["fp32_pass_rows"] = st[2]
        if with_e2e:
            eng = pipeline.MatchDeformEngine(deformer, alpha=args.alpha, prec=args.prec, device=device, use_cuda_graph=use_graph)
            for q in range(nring):
                eng.put_graphs(q, graphs[q])
            pending = []

            def estep(i):
                h, _ = ring[i % nring]
                pending.append(eng.submit(h["feat1"], h["feat2"], h["xyz1"], h["xyz2"], graph_key=i % nring))
                if len(pending) == 4:                 # four steps in flight (one per engine slot): copies of i+3 / i-1 under the kernels of i .. i+2
                    eng.result(pending.pop(0))

            def drain():
                while pending:
                    eng.result(pending.pop(0))

            e2e_steps = max(4, steps)                 # the pipeline's fill and drain (one step latency) are inside the timed region
            ems = timed_loop(estep, e2e_steps, max(9, warmup), dist, device, finish=drain)      # every slot (4) is used twice before timing: eager, then captured
            h0 = ring[0][0]
            res["e2e"] = dict(ms=ems, steps=e2e_steps, h2d=eng.h2d_bytes(h0["feat1"], h0["feat2"], h0["xyz1"], h0["xyz2"]), d2h=eng.d2h_bytes(),
                              launch_mode=eng.launch_mode)
            del eng
        res["sample"] = (ring[0][1], keep[0] if keep else eager_step(0))           # for the parity leg of cpu_baseline
        del cgs
        return res, ring, graphs

    main, ring, graphs = bench_size(n, B, args.steps, max(args.warmup, 3), with_e2e=True)
    total_pairs = B * world * args.steps
    value = total_pairs / (main["ms"] * 1e-3)
    flops_per_launch = 2.0 * n * n * C * (2 * B)                       # both directions of B pairs in one candidate launch
    achieved = flops_per_launch / (main["cand_avg_ms"] * 1e-3) / 1e12 if main["cand_avg_ms"] else float("nan")
    achieved_fused = flops_per_launch / (main["fused_avg_ms"] * 1e-3) / 1e12 if main["fused_avg_ms"] else float("nan")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("n") == n and tj.get("prec") == args.prec:          # ncu capture of the same kernel / size: scale to this launch
            traffic = tj.get("dram_bytes_per_problem", 0) * 2 * B or None
    e2e = main["e2e"]
    kern = "softmap_cand_tc_kernel (priming + sweep)" if args.prec != "fp32" else "softmap_cand_simt_kernel"
    step_ms = main["ms"] / args.steps
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype=("f16 similarity (tcgen05, fp32 accumulate) + fp32 exact re-scoring" if args.prec == "f16" else
               "bf16 similarity (tcgen05, fp32 accumulate) + fp32 exact re-scoring" if args.prec == "bf16" else "f32"),
        data="synthetic",
        config=bench_config(args, world),
        roofline=dict(bound="tensor", achieved=achieved, peak=peaks["tflops_burst"], unit="TFLOP/s", frac=achieved / peaks["tflops_burst"],
                      traffic=traffic, kernel=kern, peak_source=f"{peaks['source']} bf16 burst (the timed loop is {main['ms'] * 1e-3:.2f} s)",
                      peak_sustained=peaks["tflops_sustained"], frac_of_sustained=achieved / peaks["tflops_sustained"],
                      flops_per_launch=flops_per_launch, avg_launch_ms=main["cand_avg_ms"],
                      share_of_step=main["cand_avg_ms"] * args.steps / main["eager_ms"] if main["eager_ms"] else None,
                      timing="CUDA events inside the library on the launching stream, eager replay of the same steps"),
        roofline_fused=dict(bound="tensor", achieved=achieved_fused, peak=peaks["tflops_burst"], unit="TFLOP/s",
                            frac=achieved_fused / peaks["tflops_burst"], frac_of_sustained=achieved_fused / peaks["tflops_sustained"],
                            kernel="dvm_softmap_fwd: operand prep + priming + sweep + finalize + rescue (the fused similarity->softmax->soft-map op)",
                            flops_per_launch=flops_per_launch, avg_call_ms=main["fused_avg_ms"]),
        e2e=dict(value=B * world * e2e["steps"] / (e2e["ms"] * 1e-3), unit=UNIT, h2d_bytes_per_step=e2e["h2d"], d2h_bytes_per_step=e2e["d2h"],
                 results="hard maps T, soft map (top-10 idx + w), transferred and deformed coordinates, ARAP, Chamfer terms",
                 launch=e2e["launch_mode"]),
        gpu_launches=int(main["launches"]), clocks=main["clocks"],
        extra=dict(input_ring=main["nring"], uncertified_rows_frac=main["uncertified_rows_frac"], fp32_pass_rows=main["fp32_pass_rows"],
                   graph_build_cold_s_per_pair=main["graph_cold_s"], sim_tflops=achieved,
                   ms_per_step_eager=main["eager_ms"] / args.steps, ms_per_step_by_rank=main["rank_ms"], clocks_by_rank=main["clocks_by_rank"]),
    )
    sample = main.pop("sample")
    del ring, graphs
    torch.cuda.empty_cache()
    if not args.no_5k:
        s5, r5, g5 = bench_size(4995, 16, max(10, args.steps), 3, with_e2e=False)
        s5.pop("sample")
        del r5, g5
        torch.cuda.empty_cache()
        f5 = 2.0 * 4995 * 4995 * C * 32
        a5 = f5 / (s5["cand_avg_ms"] * 1e-3) / 1e12
        a5f = f5 / (s5["fused_avg_ms"] * 1e-3) / 1e12
        line["also_5k"] = dict(workload="match+deform, N=M=4995, 16 pairs/step/GPU (configs 1/3/4 scale)",
                               value=16 * world * max(10, args.steps) / (s5["ms"] * 1e-3), unit=UNIT, ms_per_step=s5["ms"] / max(10, args.steps),
                               ms_per_step_eager=s5["eager_ms"] / max(10, args.steps),
                               sim_tflops=a5, frac_of_peak=a5 / peaks["tflops_burst"], fused_tflops=a5f, fused_frac_of_peak=a5f / peaks["tflops_burst"],
                               uncertified_rows_frac=s5["uncertified_rows_frac"], graph_build_cold_s_per_pair=s5["graph_cold_s"])
    if rank == 0 and world == 1 and not args.no_hbm:
        line["roofline_hbm"] = hbm_microbench(device, peaks)
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_torch_baseline:
        tb = {}
        for nn_ in (4995, 20000):
            try:
                v, _, smp, sec, _ = reference_pairs_per_s(nn_, args.alpha, args.regime, -1, steps=2, warmup=1, device=str(device))
                tb[f"n{nn_}"] = dict(value=v, unit=UNIT, ms_per_pair=sec * 1e3, sample=smp)
            except RuntimeError as e:                   # the dense N x M chain does not fit / is not supported at this size
                tb[f"n{nn_}"] = dict(value=None, error=str(e)[:200])
            torch.cuda.empty_cache()
        tb["kind"] = "port (oracle restatement of the reference's torch op sequence run with device='cuda'; /root/reference does not travel to the GPU box)"
        line["torch_cuda_baseline"] = tb
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = {}
        v, cores, smp, _, _ = reference_pairs_per_s(n, args.alpha, args.regime, args.cpu_rows, steps=1, results=res)
        cb = dict(value=v, unit=UNIT, cores=cores, kind="port", sample=smp)
        # parity leg: the oracle's slab results (reference semantics, fp32 CPU) against this run's GPU results on the same pair
        if rank == 0 and "t12" in res:
            dev_in, out = sample
            from dv_matcher_b200 import synthetic
            chk = synthetic.make_batch(1, n, n, first_pair=0, regime=args.regime)
            if torch.equal(chk["feat1"][0], dev_in["feat1"][0].cpu()):
                rows = res["rows"]
                t_gpu = out["T"][0].cpu()[rows]
                mism = int((t_gpu != res["t12"][0, :, 0]).sum())
                perr = float((out["verts_t"][0].cpu()[rows] - res["verts_t"][0]).abs().max() / chk["xyz2"].abs().max())
                cb["parity_vs_gpu"] = dict(rows=int(len(rows)), hard_map_mismatches=mism, verts12_max_err_over_scale=perr,
                                           note="GPU f16 path vs the oracle's fp32 reference-form ops on the slab rows of pair 0")
        line["cpu_baseline"] = cb
        if n > 5000 and not args.no_5k:
            v5, _, smp5, sec5, _ = reference_pairs_per_s(4995, args.alpha, args.regime, -1, steps=1)
            line["cpu_baseline_5k"] = dict(value=v5, unit=UNIT, cores=cores, kind="port", ms_per_pair=sec5 * 1e3, sample=smp5)
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        if args.mode == "train":
            from tools import bench_train
            bench_train.run(args)
        else:
            run_b200(args)


if __name__ == "__main__":
    main()
