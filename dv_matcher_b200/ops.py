"""Functional (non-autograd) wrappers: one Python function per C entry point of libdvm_b200.

Each takes/returns torch CUDA tensors, allocates the outputs, fetches scratch from the grow-only
workspace cache, launches on torch's current stream and raises RuntimeError on a non-zero code.
"""
from collections import namedtuple

import torch

from . import _lib
from ._lib import check, f32c, ptr, require_device, stream_ptr

SoftMapOut = namedtuple("SoftMapOut", "argmin top_idx top_w top_d row_min row_sum piv stats")

DEFAULT_PREC = "f16"


def softmap_fwd(x, y, v=None, alpha=100.0, topk=10, soft=True, prec=None, want_stats=False):
    """Fused similarity -> softmax -> top-k -> Pi.V -> arg-min (dvm_softmap_fwd).

    x [B,N,C], y [B,M,C], v [B,M,Dv] or None.  Returns SoftMapOut (argmin i64 [B,N]; top_idx i32,
    top_w, top_d [B,N,topk]; row_min, row_sum [B,N]; piv [B,N,Dv] or None; stats i32[4] or None).
    """
    lib = _lib.load()
    x, y = f32c(x), f32c(y)
    require_device(x)
    B, N, C = x.shape
    M = y.shape[1]
    if y.shape[0] != B or y.shape[2] != C:
        raise RuntimeError(f"softmap_fwd: shape mismatch x{tuple(x.shape)} y{tuple(y.shape)}")
    prec_i = _lib.PREC[prec or DEFAULT_PREC]
    dev = x.device
    Dv = 0
    if v is not None:
        v = f32c(v)
        Dv = v.shape[-1]
    argmin = torch.empty(B, N, dtype=torch.int64, device=dev)
    top_idx = torch.empty(B, N, topk, dtype=torch.int32, device=dev)
    top_d = torch.empty(B, N, topk, dtype=torch.float32, device=dev)
    top_w = torch.empty(B, N, topk, dtype=torch.float32, device=dev) if soft else None
    row_min = torch.empty(B, N, dtype=torch.float32, device=dev)
    row_sum = torch.empty(B, N, dtype=torch.float32, device=dev) if soft else None
    piv = torch.empty(B, N, Dv, dtype=torch.float32, device=dev) if (soft and v is not None) else None
    stats = torch.empty(4, dtype=torch.int32, device=dev) if want_stats else None
    nbytes = lib.dvm_softmap_workspace_bytes(B, N, M, C, prec_i)
    ws = _lib.workspace.get(nbytes, dev, "softmap")
    rc = lib.dvm_softmap_fwd(ptr(x), ptr(y), ptr(v) if piv is not None else None, B, N, M, C, Dv if piv is not None else 0,
                             float(alpha), int(topk), _lib.MODE_SOFT if soft else _lib.MODE_HARD, prec_i,
                             ptr(argmin), ptr(top_idx), ptr(top_w), ptr(top_d), ptr(row_min), ptr(row_sum), ptr(piv), ptr(stats),
                             ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dvm_softmap_fwd")
    return SoftMapOut(argmin, top_idx, top_w, top_d, row_min, row_sum, piv, stats)


def softmap_bwd(x, y, alpha, out, d_w, prec="fp32"):
    """Gradient of the kept top-k weights w.r.t. x and y. Returns (dx, dy).
    prec "fp32": dvm_softmap_bwd (CUDA cores, exact form); "f16" / "bf16": dvm_softmap_bwd_tc (dense part on tcgen05; C <= 128)."""
    lib = _lib.load()
    x, y = f32c(x), f32c(y)
    B, N, C = x.shape
    M = y.shape[1]
    topk = out.top_idx.shape[-1]
    d_w = f32c(d_w)
    dx = torch.empty_like(x)
    dy = torch.zeros_like(y)
    if prec != "fp32" and C <= 128:
        nbytes = lib.dvm_softmap_bwd_tc_workspace_bytes(B, N, M, C)
        ws = _lib.workspace.get(nbytes, x.device, "softmap_bwd_tc")
        rc = lib.dvm_softmap_bwd_tc(ptr(x), ptr(y), B, N, M, C, float(alpha), topk,
                                    ptr(out.top_idx), ptr(out.top_w), ptr(out.top_d), ptr(out.row_min), ptr(out.row_sum), ptr(d_w),
                                    ptr(dx), ptr(dy), ptr(ws), ws.numel(), stream_ptr())
        check(rc, "dvm_softmap_bwd_tc")
        return dx, dy
    nbytes = lib.dvm_softmap_bwd_workspace_bytes(B, N, M, C)
    ws = _lib.workspace.get(nbytes, x.device, "softmap_bwd")
    rc = lib.dvm_softmap_bwd(ptr(x), ptr(y), B, N, M, C, float(alpha), topk,
                             ptr(out.top_idx), ptr(out.top_w), ptr(out.top_d), ptr(out.row_min), ptr(out.row_sum), ptr(d_w),
                             ptr(dx), ptr(dy), ptr(ws), ws.numel(), stream_ptr())
    check(rc, "dvm_softmap_bwd")
    return dx, dy


def sparse_transfer_fwd(idx, w, y):
    """out[b,i,:] = sum_k w[b,i,k] y[b, idx[b,i,k], :]  (y [B,M,D])."""
    lib = _lib.load()
    y = f32c(y)
    require_device(y)
    B, N, K = idx.shape
    M, D = y.shape[1], y.shape[2]
    out = torch.empty(B, N, D, dtype=torch.float32, device=y.device)
    check(lib.dvm_sparse_transfer_fwd(ptr(idx, torch.int32), ptr(w, torch.float32), ptr(y), B, N, M, K, D, ptr(out), stream_ptr()), "dvm_sparse_transfer_fwd")
    return out


def sparse_transfer_bwd(idx, w, y, d_out, need_dw=True, need_dy=True):
    lib = _lib.load()
    y, d_out = f32c(y), f32c(d_out)
    B, N, K = idx.shape
    M, D = y.shape[1], y.shape[2]
    dw = torch.empty(B, N, K, dtype=torch.float32, device=y.device) if need_dw else None
    dy = torch.zeros_like(y) if need_dy else None
    check(lib.dvm_sparse_transfer_bwd(ptr(idx, torch.int32), ptr(w, torch.float32), ptr(y), ptr(d_out), B, N, M, K, D, ptr(dw), ptr(dy), stream_ptr()),
          "dvm_sparse_transfer_bwd")
    return dw, dy


def knn3(q, r, k, f64=False, want_d2=False, idx_dtype=torch.int64, algo="auto"):
    """Exact k-NN on 3-D points: idx [B,N,k] (+ squared distances).  algo: "auto" (uniform grid for reference
    clouds >= 1024 points, else brute force) or "brute"; both give bit-identical results."""
    lib = _lib.load()
    q, r = f32c(q), f32c(r)
    require_device(q)
    B, N, _ = q.shape
    M = r.shape[1]
    if q.shape[-1] != 3 or r.shape[-1] != 3 or r.shape[0] != B:
        raise RuntimeError(f"knn3: expected [B,N,3] / [B,M,3], got {tuple(q.shape)} / {tuple(r.shape)}")
    idx = torch.empty(B, N, k, dtype=idx_dtype, device=q.device)
    d2 = torch.empty(B, N, k, dtype=torch.float64 if f64 else torch.float32, device=q.device) if want_d2 else None
    i64 = ptr(idx) if idx_dtype == torch.int64 else None
    i32 = ptr(idx) if idx_dtype == torch.int32 else None
    ws = None
    if algo != "brute":
        nbytes = lib.dvm_knn3_workspace_bytes(B, N, M)
        ws = _lib.workspace.get(nbytes, q.device, "knn3") if nbytes else None
    check(lib.dvm_knn3(ptr(q), ptr(r), B, N, M, int(k), int(bool(f64)), i64, i32,
                       ptr(d2) if (want_d2 and not f64) else None, ptr(d2) if (want_d2 and f64) else None,
                       ptr(ws), ws.numel() if ws is not None else 0, stream_ptr()), "dvm_knn3")
    return (idx, d2) if want_d2 else idx


def chamfer_fwd(a, b, algo="auto"):
    lib = _lib.load()
    a, b = f32c(a), f32c(b)
    require_device(a)
    B, N, _ = a.shape
    M = b.shape[1]
    dev = a.device
    d1 = torch.empty(B, N, dtype=torch.float32, device=dev)
    d2 = torch.empty(B, M, dtype=torch.float32, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
    i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    ws = None
    if algo != "brute":
        nbytes = lib.dvm_chamfer_workspace_bytes(B, N, M)
        ws = _lib.workspace.get(nbytes, dev, "knn3") if nbytes else None
    check(lib.dvm_chamfer_fwd(ptr(a), ptr(b), B, N, M, ptr(d1), ptr(d2), ptr(i1), ptr(i2),
                              ptr(ws), ws.numel() if ws is not None else 0, stream_ptr()), "dvm_chamfer_fwd")
    return d1, d2, i1, i2


def chamfer_bwd(a, b, i1, i2, g1, g2):
    lib = _lib.load()
    a, b, g1, g2 = f32c(a), f32c(b), f32c(g1), f32c(g2)
    B, N, _ = a.shape
    M = b.shape[1]
    da = torch.empty_like(a)
    db = torch.empty_like(b)
    check(lib.dvm_chamfer_bwd(ptr(a), ptr(b), ptr(i1), ptr(i2), ptr(g1), ptr(g2), B, N, M, ptr(da), ptr(db), stream_ptr()), "dvm_chamfer_bwd")
    return da, db


def fps(xyz, k, start):
    """Farthest point sampling with injected start indices: xyz [B,N,3], start i64 [B] -> i64 [B,k]."""
    lib = _lib.load()
    xyz = f32c(xyz)
    require_device(xyz)
    B, N, _ = xyz.shape
    start = start.to(device=xyz.device, dtype=torch.int64).contiguous()
    out = torch.empty(B, k, dtype=torch.int64, device=xyz.device)
    ws = _lib.workspace.get(lib.dvm_fps_workspace_bytes(B, N), xyz.device, "fps")
    check(lib.dvm_fps(ptr(xyz), B, N, int(k), ptr(start), ptr(out), ptr(ws), ws.numel(), stream_ptr()), "dvm_fps")
    return out


def graph_weights(xyz, nodes_idx):
    """Graph tensors for given node lists: (influence i64 [B,N,3], dists, weights, ring i64 [B,K,9], sigma f64 [B])."""
    lib = _lib.load()
    xyz = f32c(xyz)
    require_device(xyz)
    B, N, _ = xyz.shape
    K = nodes_idx.shape[1]
    dev = xyz.device
    nodes_idx = nodes_idx.to(device=dev, dtype=torch.int64).contiguous()
    infl = torch.empty(B, N, 3, dtype=torch.int64, device=dev)
    dists = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    wts = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    ring = torch.empty(B, K, 9, dtype=torch.int64, device=dev)
    sigma = torch.empty(B, dtype=torch.float64, device=dev)
    ws = _lib.workspace.get(lib.dvm_graph_workspace_bytes(B, N, K), dev, "graph")
    check(lib.dvm_graph_weights(ptr(xyz), ptr(nodes_idx), B, N, K, ptr(infl), ptr(dists), ptr(wts), ptr(ring), ptr(sigma),
                                ptr(ws), ws.numel(), stream_ptr()), "dvm_graph_weights")
    return infl, dists, wts, ring, sigma


def rot6d_fwd(d6):
    lib = _lib.load()
    d6 = f32c(d6)
    require_device(d6)
    n = d6.numel() // 6
    R = torch.empty(*d6.shape[:-1], 3, 3, dtype=torch.float32, device=d6.device)
    check(lib.dvm_rot6d_fwd(ptr(d6), n, ptr(R), stream_ptr()), "dvm_rot6d_fwd")
    return R


def rot6d_bwd(d6, dR):
    lib = _lib.load()
    d6, dR = f32c(d6), f32c(dR)
    n = d6.numel() // 6
    dd6 = torch.empty_like(d6)
    check(lib.dvm_rot6d_bwd(ptr(d6), ptr(dR), n, ptr(dd6), stream_ptr()), "dvm_rot6d_bwd")
    return dd6


def skin_fwd(xyz, nodes_idx, infl, wts, R, t):
    lib = _lib.load()
    xyz, wts, R, t = f32c(xyz), f32c(wts), f32c(R), f32c(t)
    require_device(xyz)
    B, N, _ = xyz.shape
    K = nodes_idx.shape[1]
    out = torch.empty_like(xyz)
    check(lib.dvm_skin_fwd(ptr(xyz), ptr(nodes_idx), ptr(infl), ptr(wts), ptr(R), ptr(t), B, N, K, ptr(out), stream_ptr()), "dvm_skin_fwd")
    return out


def skin_bwd(xyz, nodes_idx, infl, wts, d_out):
    lib = _lib.load()
    xyz, wts, d_out = f32c(xyz), f32c(wts), f32c(d_out)
    B, N, _ = xyz.shape
    K = nodes_idx.shape[1]
    dR = torch.empty(B, K, 3, 3, dtype=torch.float32, device=xyz.device)
    dt = torch.empty(B, K, 3, dtype=torch.float32, device=xyz.device)
    check(lib.dvm_skin_bwd(ptr(xyz), ptr(nodes_idx), ptr(infl), ptr(wts), ptr(d_out), B, N, K, ptr(dR), ptr(dt), stream_ptr()), "dvm_skin_bwd")
    return dR, dt


def arap_fwd(xyz, nodes_idx, ring, R, t):
    lib = _lib.load()
    xyz, R, t = f32c(xyz), f32c(R), f32c(t)
    require_device(xyz)
    B, N, _ = xyz.shape
    K, rk = ring.shape[1], ring.shape[2]
    arap = torch.empty(B, dtype=torch.float32, device=xyz.device)
    sr = torch.empty(B, dtype=torch.float32, device=xyz.device)
    ws = _lib.workspace.get(lib.dvm_arap_workspace_bytes(B, K), xyz.device, "arap")
    check(lib.dvm_arap_fwd(ptr(xyz), ptr(nodes_idx), ptr(ring), ptr(R), ptr(t), B, N, K, rk, ptr(arap), ptr(sr),
                           ptr(ws), ws.numel(), stream_ptr()), "dvm_arap_fwd")
    return arap, sr


def arap_bwd(xyz, nodes_idx, ring, R, t, g_arap, dR, dt):
    """Accumulates g_arap[b] * d arap/d(R,t) into dR, dt (in place)."""
    lib = _lib.load()
    xyz, R, t, g_arap = f32c(xyz), f32c(R), f32c(t), f32c(g_arap)
    B, N, _ = xyz.shape
    K, rk = ring.shape[1], ring.shape[2]
    check(lib.dvm_arap_bwd(ptr(xyz), ptr(nodes_idx), ptr(ring), ptr(R), ptr(t), ptr(g_arap), B, N, K, rk, ptr(dR), ptr(dt), stream_ptr()),
          "dvm_arap_bwd")
    return dR, dt


def node_table(R, t, nodes_xyz, node_perm=None):
    """Pack (R [B,K,3,3], t [B,K,3], g [B,K,3]) into 64-byte node records [B,K,16] (dvm_node_table).  With node_perm
    (int32 [B,K], new -> old) R and t are in the reference's node order and the table comes out in the packed (Morton) order."""
    lib = _lib.load()
    R, t, nodes_xyz = f32c(R), f32c(t), f32c(nodes_xyz)
    require_device(R)
    B, K = nodes_xyz.shape[0], nodes_xyz.shape[1]
    table = torch.empty(B, K, 16, dtype=torch.float32, device=R.device)
    check(lib.dvm_node_table(ptr(R), ptr(t), ptr(nodes_xyz), ptr(node_perm, torch.int32), B, K, ptr(table), stream_ptr()), "dvm_node_table")
    return table


def node_table_from_d9(d9, nodes_xyz, want_rt=False, node_perm=None):
    """Deformer output d9 [B,K,9] -> node records (identity offset + 6D -> R fused; dvm_node_table_from_d9).
    Returns table or (table, R [B,K,3,3], t [B,K,3])."""
    lib = _lib.load()
    d9, nodes_xyz = f32c(d9), f32c(nodes_xyz)
    require_device(d9)
    B, K = nodes_xyz.shape[0], nodes_xyz.shape[1]
    table = torch.empty(B, K, 16, dtype=torch.float32, device=d9.device)
    R = torch.empty(B, K, 3, 3, dtype=torch.float32, device=d9.device) if want_rt else None
    t = torch.empty(B, K, 3, dtype=torch.float32, device=d9.device) if want_rt else None
    check(lib.dvm_node_table_from_d9(ptr(d9), ptr(nodes_xyz), ptr(node_perm, torch.int32), B, K, ptr(table), ptr(R), ptr(t), stream_ptr()),
          "dvm_node_table_from_d9")
    return (table, R, t) if want_rt else table


def skin_fwd_packed(xyz, pack, table):
    """Skinning warp on the packed graph layout (dvm_skin_fwd_packed). pack: deformation_graph.GraphPack."""
    lib = _lib.load()
    xyz = f32c(xyz)
    require_device(xyz)
    B, N, _ = xyz.shape
    K = table.shape[1]
    out = torch.empty_like(xyz)
    check(lib.dvm_skin_fwd_packed(ptr(pack.s_xyz, torch.float32), ptr(pack.vorder, torch.int32), ptr(pack.s_infl, torch.int32), ptr(pack.s_w, torch.float32),
                                  ptr(table, torch.float32), B, N, K, ptr(out), stream_ptr()), "dvm_skin_fwd_packed")
    return out


def skin_bwd_csr(xyz, pack, d_out):
    lib = _lib.load()
    xyz, d_out = f32c(xyz), f32c(d_out)
    B, N, _ = xyz.shape
    K = pack.nodes_xyz.shape[1]
    dR = torch.empty(B, K, 3, 3, dtype=torch.float32, device=xyz.device)
    dt = torch.empty(B, K, 3, dtype=torch.float32, device=xyz.device)
    check(lib.dvm_skin_bwd_csr(ptr(xyz), ptr(pack.nodes_xyz, torch.float32), ptr(pack.csr_ptr, torch.int32), ptr(pack.csr_vert, torch.int32),
                               ptr(pack.csr_w, torch.float32), ptr(d_out), B, N, K, ptr(dR), ptr(dt), stream_ptr()), "dvm_skin_bwd_csr")
    return dR, dt


def arap_fwd_packed(pack, table, want_sr=True):
    lib = _lib.load()
    require_device(table)
    B, K = table.shape[0], table.shape[1]
    rk = pack.s_ring.shape[1]
    arap = torch.empty(B, dtype=torch.float32, device=table.device)
    sr = torch.empty(B, dtype=torch.float32, device=table.device) if want_sr else None
    ws = _lib.workspace.get(lib.dvm_arap_packed_workspace_bytes(B, K), table.device, "arap")
    check(lib.dvm_arap_fwd_packed(ptr(pack.s_ring, torch.int32), ptr(table, torch.float32), B, K, rk,
                                  ptr(arap), ptr(sr), ptr(ws), ws.numel(), stream_ptr()), "dvm_arap_fwd_packed")
    return arap, sr


def gather_conv_fwd(feat, idx, weight, bias):
    lib = _lib.load()
    feat = f32c(feat)
    require_device(feat)
    B, N, C = feat.shape
    k = idx.shape[-1]
    R = idx.shape[1]
    idx = idx.to(torch.int64).contiguous()
    w = f32c(weight.reshape(-1))
    b = f32c(bias.reshape(-1)) if bias is not None else None
    out = torch.empty(B, R, C, dtype=torch.float32, device=feat.device)
    check(lib.dvm_gather_conv_fwd(ptr(feat), ptr(idx), ptr(w), ptr(b), B, N, R, C, k, ptr(out), stream_ptr()), "dvm_gather_conv_fwd")
    return out


def gather_conv_bwd(feat, idx, weight, d_out):
    lib = _lib.load()
    feat, d_out = f32c(feat), f32c(d_out)
    B, N, C = feat.shape
    k = idx.shape[-1]
    R = idx.shape[1]
    idx = idx.to(torch.int64).contiguous()
    w = f32c(weight.reshape(-1))
    d_feat = torch.zeros_like(feat)
    d_w = torch.zeros(k, dtype=torch.float32, device=feat.device)
    d_b = torch.zeros(1, dtype=torch.float32, device=feat.device)
    check(lib.dvm_gather_conv_bwd(ptr(feat), ptr(idx), ptr(w), ptr(d_out), B, N, R, C, k, ptr(d_feat), ptr(d_w), ptr(d_b), stream_ptr()),
          "dvm_gather_conv_bwd")
    return d_feat, d_w, d_b


def linear_act_fwd(x, weight, bias, act="none", x_cols=None):
    """act(x @ weight.T + bias) on tensor cores with fp32-equivalent accuracy (dvm_linear_act_fwd; nn.Linear weight layout).

    x: [..., K] fp32 (last-dim stride 1, a row pitch that is a multiple of 4 floats -- e.g. a padded staging buffer
    viewed through `x_cols`: x [rows, pitch] with x_cols = K valid columns)."""
    lib = _lib.load()
    require_device(x)
    K = x.shape[-1] if x_cols is None else x_cols
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype != torch.float32 or x2.stride(1) != 1 or x2.stride(0) % 4 or x2.data_ptr() % 16:
        xp = torch.empty(x2.shape[0], (K + 3) // 4 * 4, dtype=torch.float32, device=x.device)
        xp[:, :K] = x2[:, :K]
        x2 = xp
    N, Kw = weight.shape
    if Kw != K:
        raise ValueError(f"linear_act_fwd: weight is [{N},{Kw}] but x has {K} columns")
    w = weight.detach()
    if w.dtype != torch.float32 or w.stride(1) != 1 or w.stride(0) % 4 or w.data_ptr() % 16:
        wp = torch.empty(N, (K + 3) // 4 * 4, dtype=torch.float32, device=x.device)
        wp[:, :K] = w
        w = wp
    b = f32c(bias.detach()) if bias is not None else None
    out = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
    check(lib.dvm_linear_act_fwd(x2.data_ptr(), x2.shape[0], K, x2.stride(0), w.data_ptr(), w.stride(0), ptr(b), N, {"none": 0, "elu": 1}[act],
                                 ptr(out), N, stream_ptr()), "dvm_linear_act_fwd")
    return out.reshape(*lead, N)


def topk_select(scores, k, n_valid=None):
    """Indices of the k largest entries of every row of scores [R, pitch] (first n_valid columns), descending, int64 [R,k]."""
    lib = _lib.load()
    require_device(scores)
    if scores.dtype != torch.float32 or scores.dim() != 2 or scores.stride(1) != 1:
        raise RuntimeError("topk_select: scores must be a float32 [rows, pitch] matrix with unit column stride")
    R = scores.shape[0]
    N = scores.shape[1] if n_valid is None else n_valid
    idx = torch.empty(R, k, dtype=torch.int64, device=scores.device)
    check(lib.dvm_topk_select(scores.data_ptr(), R, N, scores.stride(0), int(k), ptr(idx), stream_ptr()), "dvm_topk_select")
    return idx


LINEAR_MAX_OUT = 8192          # dvm_linear_act_fwd stages its bias row in shared memory: wider outputs go in column blocks


def linear_into(x, w, bias, out):
    """out[:, :N] = x @ w.T (+ bias) with dvm_linear_act_fwd, any N: column blocks of LINEAR_MAX_OUT written straight into `out`
    ([rows, pitch >= N] float32, pitch a multiple of 4).  x [rows,K], w [N,K] float32 with row pitches that are multiples of 4."""
    lib = _lib.load()
    rows, K = x.shape
    N = w.shape[0]
    if out.stride(0) % 4 or out.stride(1) != 1 or out.data_ptr() % 16 or x.stride(0) % 4 or w.stride(0) % 4 or x.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError("linear_into: row pitches must be multiples of 4 floats, unit column stride, 16-byte aligned")
    for c0 in range(0, N, LINEAR_MAX_OUT):
        nc = min(LINEAR_MAX_OUT, N - c0)
        check(lib.dvm_linear_act_fwd(x.data_ptr(), rows, K, x.stride(0), w.data_ptr() + 4 * c0 * w.stride(0), w.stride(0),
                                     (bias.data_ptr() + 4 * c0) if bias is not None else None, nc, 0,
                                     out.data_ptr() + 4 * c0, out.stride(0), stream_ptr()), "dvm_linear_act_fwd")
    return out


def knn_feature_large(a, b, k):
    """knn(a, b, k) of models/loss.py:451-462 for k > 10: scores 2 a.b - |b|^2 on tcgen05 (3xTF32), radix selection of the k best."""
    a, b = f32c(a), f32c(b)
    B, S, C = a.shape
    N = b.shape[1]
    npad = (N + 3) // 4 * 4
    out = torch.empty(B, S, k, dtype=torch.int64, device=a.device)
    bias = ((-0.5) * (b * b).sum(-1)).contiguous()                     # [B,N]
    rows_max = max(128, (256 << 20) // (4 * npad))                     # the score matrix only exists as row chunks of <= 256 MB
    sc = torch.empty(min(S, rows_max), npad, dtype=torch.float32, device=a.device)
    for i in range(B):
        for r0 in range(0, S, rows_max):
            r = min(rows_max, S - r0)
            linear_into(a[i, r0:r0 + r], b[i], bias[i], sc[:r])
            out[i, r0:r0 + r] = topk_select(sc[:r], k, n_valid=N)
    return out


def pair_dist_fwd(feat, qidx, nbr, geo=None):
    """d [B,S,k] = |feat[b, nbr] - feat[b, qidx]| (+ geo[b, nbr, qidx] as float32 when a [B,N,N] matrix is given)."""
    lib = _lib.load()
    feat = f32c(feat)
    require_device(feat)
    B, N, C = feat.shape
    S, k = nbr.shape[1], nbr.shape[2]
    d = torch.empty(B, S, k, dtype=torch.float32, device=feat.device)
    g_out = None
    if geo is not None:
        if geo.dtype not in (torch.float32, torch.float64) or tuple(geo.shape) != (B, N, N) or not geo.is_contiguous():
            raise RuntimeError("pair_dist_fwd: geo must be a contiguous float32/float64 [B,N,N] tensor")
        g_out = torch.empty_like(d)
    check(lib.dvm_pair_dist_fwd(ptr(feat), ptr(qidx, torch.int64), ptr(nbr, torch.int64), B, N, C, S, k, ptr(d),
                                ptr(geo), int(geo is not None and geo.dtype == torch.float64), ptr(g_out), stream_ptr()), "dvm_pair_dist_fwd")
    return d, g_out


def pair_dist_bwd(feat, qidx, nbr, d, g):
    lib = _lib.load()
    feat, d, g = f32c(feat), f32c(d), f32c(g)
    B, N, C = feat.shape
    S, k = nbr.shape[1], nbr.shape[2]
    dfeat = torch.zeros_like(feat)
    check(lib.dvm_pair_dist_bwd(ptr(feat), ptr(qidx, torch.int64), ptr(nbr, torch.int64), ptr(d), ptr(g), B, N, C, S, k, ptr(dfeat), stream_ptr()),
          "dvm_pair_dist_bwd")
    return dfeat


def gather_rows_fwd(points, idx2):
    """points [B,N,C], idx2 int64 [B,R] -> [B,R,C]."""
    lib = _lib.load()
    points = f32c(points)
    require_device(points)
    B, N, C = points.shape
    R = idx2.shape[1]
    out = torch.empty(B, R, C, dtype=torch.float32, device=points.device)
    check(lib.dvm_gather_rows_fwd(ptr(points), ptr(idx2, torch.int64), B, N, R, C, ptr(out), stream_ptr()), "dvm_gather_rows_fwd")
    return out


def gather_rows_bwd(d_out, idx2, N):
    lib = _lib.load()
    d_out = f32c(d_out)
    B, R, C = d_out.shape
    d_pts = torch.zeros(B, N, C, dtype=torch.float32, device=d_out.device)
    check(lib.dvm_gather_rows_bwd(ptr(d_out), ptr(idx2, torch.int64), B, N, R, C, ptr(d_pts), stream_ptr()), "dvm_gather_rows_bwd")
    return d_pts
