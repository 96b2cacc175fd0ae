#!/usr/bin/env python
"""Benchmark of the match + deform hot path (BASELINE.json metric: shape pairs/s, sim TFLOP/s).

    python bench.py --gpus N --steps K --warmup W            # our arm (B200, one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU op sequence (oracle port)

One "step" = one pass of match + deform over a batch of `--pairs` synthetic pairs per GPU at
N = M = `--n` points (default 50000), C = 128.  Pairs are independent: ranks take disjoint pairs, there is
no data-path collective (torch.distributed is only used for the barrier and the max-over-ranks time).
Prints ONE JSON line (rank 0).
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
    ap.add_argument("--n", type=int, default=50000, help="points per cloud (N = M)")
    ap.add_argument("--pairs", type=int, default=2, help="pairs per GPU per step")
    ap.add_argument("--alpha", type=float, default=100.0)
    ap.add_argument("--regime", default="structured", choices=["structured", "unstructured"])
    ap.add_argument("--prec", default="f16", choices=["f16", "bf16", "fp32"])
    ap.add_argument("--cpu-rows", type=int, default=0, help="row-slab size of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-5k", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"], source="measured (sustained)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU every 50 ms during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {getattr(pynvml, n): n for n in dir(pynvml) if n.startswith("nvmlClocksThrottleReason") and not n.endswith("All")}
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
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


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU op sequence on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(n, alpha, regime, rows, steps=1, warmup=0):
    from dv_matcher_b200 import synthetic
    from dv_matcher_b200.deformer import Deformer
    from oracle import pipeline as opipe
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if rows <= 0:
        rows = max(128, min(n, int(2048 * 5000 / max(n, 5000)) * 4)) if n > 5000 else min(n, 1024)
    batch = synthetic.make_batch(1, n, n, first_pair=0, regime=regime)
    torch.manual_seed(0)
    params = {k: v.detach() for k, v in Deformer(10).state_dict().items()}
    graph = opipe.make_cpu_graph(batch["xyz1"][0])
    for _ in range(warmup):
        opipe.time_pair_sample(batch, graph, params, alpha, rows)
    ts = [opipe.time_pair_sample(batch, graph, params, alpha, rows) for _ in range(max(1, steps))]
    sec = min(ts)
    sample = (f"oracle port of the reference's torch-CPU op sequence on a slab of {min(rows, n)} of {n} source rows per direction "
              f"(x N/rows), per-cloud parts in full, graphs warm; best of {max(1, steps)}")
    return 1.0 / sec, cores, sample, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sample, sec = cpu_reference_pairs_per_s(args.n, args.alpha, args.regime, args.cpu_rows, steps=args.steps, warmup=min(args.warmup, 1))
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=f"match+deform, N=M={args.n}, C={C}, alpha={args.alpha}, {args.regime} features (config 5 of BASELINE.json)",
                            pairs_per_step=1),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def timed_loop(fn, steps, warmup, dist, device):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
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
    import ctypes

    peaks = measured_peaks()
    torch.manual_seed(0)
    deformer = Deformer(10).to(device).eval()
    n, B = args.n, args.pairs

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
        ring = make_ring(npts, nb, nring)
        t0 = time.perf_counter()
        graphs = [build_graphs(torch.cat([d["xyz1"], d["xyz2"]]), torch.arange(2 * nb) % npts) for _, d in ring]
        torch.cuda.synchronize(device)
        graph_cold_s = (time.perf_counter() - t0) / (nring * nb)

        def step(i):
            _, d = ring[i % nring]
            with torch.no_grad():
                return pipeline.match_deform(d["feat1"], d["feat2"], d["xyz1"], d["xyz2"], graphs[i % nring], deformer,
                                             alpha=args.alpha, prec=args.prec)

        sampler = ClockSampler(physical_gpu_index(local))
        l0 = lib.dvm_launch_count()
        # warm-up outside the profile/clock window
        for i in range(warmup):
            step(i)
        torch.cuda.synchronize(device)
        lib.dvm_profile_enable(1)
        sampler.start()
        lw = lib.dvm_launch_count()
        ms = timed_loop(step, steps, 0, dist, device)
        rank_ms = [round(v / steps, 3) for v in timed_loop.rank_ms]
        launches = lib.dvm_launch_count() - lw
        sampler.stop_flag = True
        sampler.join(timeout=2)
        tot, cnt = ctypes.c_double(0), ctypes.c_int(0)
        lib.dvm_profile_read(ctypes.byref(tot), ctypes.byref(cnt))
        lib.dvm_profile_enable(0)
        clocks_by_rank = None
        if dist is not None:                       # diagnostic: median SM clock and throttle reasons of every rank's GPU
            objs = [None] * world
            dist.all_gather_object(objs, sampler.summary())
            clocks_by_rank = [[o.get("sm_mhz"), o.get("reasons")] for o in objs]
        res = dict(ms=ms, launches=launches, nring=nring, graph_cold_s=graph_cold_s, clocks=sampler.summary(), rank_ms=rank_ms,
                   clocks_by_rank=clocks_by_rank,
                   cand_ms=tot.value, cand_launches=cnt.value)
        # stats of the last step: rows the 16-bit pass could not certify
        from dv_matcher_b200 import ops
        _, d = ring[0]
        o = ops.softmap_fwd(torch.cat([d["feat1"], d["feat2"]]), torch.cat([d["feat2"], d["feat1"]]), None, alpha=args.alpha,
                            prec=args.prec, want_stats=True)
        st = o.stats.cpu().tolist()
        res["uncertified_rows_frac"] = st[0] / float(2 * nb * npts)
        if with_e2e:
            eng = pipeline.MatchDeformEngine(deformer, alpha=args.alpha, prec=args.prec, device=device)
            for q in range(nring):
                eng._graph_cache[q] = graphs[q]

            def estep(i):
                h, _ = ring[i % nring]
                hn, _ = ring[(i + 1) % nring]                        # next step's inputs: their H2D copy overlaps this step's kernels
                return eng.step(h["feat1"], h["feat2"], h["xyz1"], h["xyz2"], graph_key=i % nring,
                                next_inputs=(hn["feat1"], hn["feat2"], hn["xyz1"], hn["xyz2"]))

            e2e_steps = max(3, steps // 2)
            ems = timed_loop(estep, e2e_steps, min(warmup, 3), dist, device)
            h0 = ring[0][0]
            res["e2e"] = dict(ms=ems, steps=e2e_steps, h2d=eng.h2d_bytes(h0["feat1"], h0["feat2"], h0["xyz1"], h0["xyz2"]), d2h=eng.d2h_bytes())
        del ring, graphs
        torch.cuda.empty_cache()
        return res

    main = bench_size(n, B, args.steps, max(args.warmup, 3), with_e2e=True)
    total_pairs = B * world * args.steps
    value = total_pairs / (main["ms"] * 1e-3)
    flops_per_launch = 2.0 * n * n * C * (2 * B)                       # both directions of B pairs in one candidate launch
    cand_avg_s = (main["cand_ms"] * 1e-3 / main["cand_launches"]) if main["cand_launches"] else float("nan")
    achieved = flops_per_launch / cand_avg_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("n") == n and tj.get("prec") == args.prec:          # ncu capture of the same kernel / size: scale to this launch
            traffic = tj.get("dram_bytes_per_problem", 0) * 2 * B or None
    e2e = main["e2e"]
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=main["ms"] / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype=("f16 similarity (tcgen05, fp32 accumulate) + fp32 exact re-scoring" if args.prec == "f16" else
               "bf16 similarity (tcgen05, fp32 accumulate) + fp32 exact re-scoring" if args.prec == "bf16" else "f32"),
        data="synthetic",
        config=dict(workload=f"match+deform, N=M={n}, C={C}, alpha={args.alpha}, {args.regime} features (config 5 of BASELINE.json)",
                    pairs_per_step_per_gpu=B, parallelism=f"pairs sharded over {world} GPU(s), no data-path collective",
                    l2_policy=f"inputs rotate through {main['nring']} distinct batches (> 126 MB L2 in total)",
                    graphs="warm (built once per shape, outside the timed region)", prec=args.prec),
        roofline=dict(bound="tensor", achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"],
                      traffic=traffic, kernel="softmap_cand_tc_kernel" if args.prec != "fp32" else "softmap_cand_simt_kernel",
                      peak_source=peaks["source"], flops_per_launch=flops_per_launch, avg_launch_ms=cand_avg_s * 1e3,
                      share_of_step=main["cand_ms"] / main["ms"] if main["ms"] else None),
        e2e=dict(value=B * world * e2e["steps"] / (e2e["ms"] * 1e-3), unit=UNIT, h2d_bytes_per_step=e2e["h2d"], d2h_bytes_per_step=e2e["d2h"]),
        gpu_launches=int(main["launches"]), clocks=main["clocks"],
        extra=dict(uncertified_rows_frac=main["uncertified_rows_frac"], graph_build_cold_s_per_pair=main["graph_cold_s"],
                   sim_tflops=achieved, ms_per_step_by_rank=main["rank_ms"], clocks_by_rank=main["clocks_by_rank"]),
    )
    if not args.no_5k:
        s5 = bench_size(4995, 16, max(10, args.steps), 3, with_e2e=False)
        f5 = 2.0 * 4995 * 4995 * C * 32
        a5 = f5 / (s5["cand_ms"] * 1e-3 / max(1, s5["cand_launches"])) / 1e12
        line["also_5k"] = dict(workload="match+deform, N=M=4995, 16 pairs/step/GPU (configs 1/3/4 scale)",
                               value=16 * world * max(10, args.steps) / (s5["ms"] * 1e-3), unit=UNIT, ms_per_step=s5["ms"] / max(10, args.steps),
                               sim_tflops=a5, frac_of_peak=a5 / peaks["tflops"], uncertified_rows_frac=s5["uncertified_rows_frac"],
                               graph_build_cold_s_per_pair=s5["graph_cold_s"])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_pairs_per_s(n, args.alpha, args.regime, args.cpu_rows, steps=1)
        line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind="port", sample=sample)
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
        run_b200(args)


if __name__ == "__main__":
    main()
