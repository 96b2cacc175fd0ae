"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the B200 box, gloo in CPU tests).

The hot path shards over independent shape pairs: rank r takes pairs r::world and inference / evaluation
needs NO collective.  Training has exactly one exchange step: the gradient all-reduce of LG-Net + Deformer
(8.5 MB fp32), done here as ONE flattened bucket per step -- at that size the cost is launch latency, so one
collective beats per-parameter calls.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, world, device)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if use_cuda else "gloo")
        kw = dict(device_id=device) if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, device


def shard_pairs(n_pairs, rank, world):
    """Indices of the shape pairs rank `rank` owns (round-robin: balanced for any n_pairs)."""
    return list(range(rank, n_pairs, world))


def allreduce_gradients(params, world=None, bucket=None):
    """Average the gradients of `params` over all ranks with ONE all-reduce of a flat fp32 bucket.

    `bucket` may be a preallocated flat tensor (reused across steps).  Parameters without a gradient contribute
    zeros (every rank must pass the same parameter list).  Returns the bucket."""
    params = [p for p in params if p.requires_grad]
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    n = sum(p.numel() for p in params)
    if n == 0:
        return bucket
    dev = params[0].device
    if bucket is None or bucket.numel() != n or bucket.device != dev:
        bucket = torch.empty(n, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            bucket[off:off + k].zero_()
        else:
            bucket[off:off + k].copy_(p.grad.reshape(-1))
        off += k
    if world > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        bucket.div_(world)
    off = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            p.grad = bucket[off:off + k].reshape(p.shape).clone()
        else:
            p.grad.copy_(bucket[off:off + k].reshape(p.shape))
        off += k
    return bucket


def gather_pair_results(local_results, world=None):
    """Host-side gather of per-pair python results (evaluation bookkeeping; not on the data path)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [local_results]
    out = [None] * world
    dist.all_gather_object(out, local_results)
    return out


# --------------------------------------------------------------------------------------------------
# one giant pair (config 5, N = M = 200k): rows shard, columns replicate, nothing is exchanged on the data path
# --------------------------------------------------------------------------------------------------
def shard_rows(n_rows, rank, world):
    """[lo, hi) of the contiguous row slab rank `rank` owns (balanced to within one row)."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def match_rows_sharded(feat1, feat2, verts2, alpha=100.0, rank=None, world=None, prec=None, gather=False, soft_map_fn=None):
    """Direction 1 -> 2 of ONE pair with the source rows sharded over the ranks: every rank holds all of
    feat2 / verts2 (102 MB + 2.4 MB at 200k: replicated), takes its slab of feat1 rows and runs the fused kernel on it --
    row softmax, top-10, arg-min and Pi @ verts2 need no exchange.  The reverse direction is the same call with the
    roles swapped.  feat1 [1,N,C], feat2 [1,M,C], verts2 [1,M,3].

    Returns dict(rows=(lo, hi), argmin, top_idx, top_w, verts_t) for the local slab; with gather=True every entry is
    all-gathered into the full [1,N,...] result (row order preserved; a host-side convenience, not part of the data path)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if soft_map_fn is None:
        from . import maps
        soft_map_fn = maps.soft_map
    lo, hi = shard_rows(feat1.shape[1], rank, world)
    sm, vt = soft_map_fn(feat1[:, lo:hi].contiguous(), feat2, alpha, v=verts2, prec=prec)
    out = dict(rows=(lo, hi), argmin=sm.argmin, top_idx=sm.idx, top_w=sm.w, verts_t=vt)
    if gather and world > 1:
        n = feat1.shape[1]
        most = max(shard_rows(n, r, world)[1] - shard_rows(n, r, world)[0] for r in range(world))
        for k in ("argmin", "top_idx", "top_w", "verts_t"):
            t = out[k]
            pad = torch.zeros(t.shape[0], most, *t.shape[2:], dtype=t.dtype, device=t.device)
            pad[:, : hi - lo] = t
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            out[k] = torch.cat([parts[r][:, : shard_rows(n, r, world)[1] - shard_rows(n, r, world)[0]] for r in range(world)], dim=1)
        out["rows"] = (0, n)
    return out
