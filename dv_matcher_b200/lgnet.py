"""LG-Net's N x N pieces on this library.  LG-Net itself (`Uni3FC`) stays the reference's PyTorch code;
what it spends its time and memory on at large N is

  * `knn_new(pcd, pcd, 40)` x 7 per forward (models/model.py:267-278)      -> `geometry.knn` (tensor-core scores + radix selection)
  * `SA_Layer` global attention x 4 per forward (models/model.py:113-123)  -> `sa_attention` below

`sa_attention` never materialises the N x N energy / attention matrices (400 MB each at N = 10k, 10 GB at 50k): it walks row
chunks with two fp32-equivalent tcgen05 GEMMs (`dvm_linear_act_fwd`, 3xTF32) around one fused row-softmax + transpose kernel
(`dvm_softmax_rows_transposed`, csrc/attention.cu); the column renormalisation `attention / (1e-9 + attention.sum(dim=1))` needs the
column sums of the whole matrix, which come out of the second GEMM through an appended row of ones.

Forward and backward (`_SAAttention`): the backward recomputes the chunk's energies and needs five GEMMs per chunk.
"""
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def _pad4(n):
    return (n + 3) // 4 * 4


def _attention_forward(x_q, x_k, x_v, chunk):
    """-> (x_r [B,C,N], column sums s [B,N] of the row-softmax matrix)."""
    lib = _lib.load()
    B, N, c = x_q.shape
    C = x_v.shape[1]
    dev = x_q.device
    if c % 4:
        raise RuntimeError("sa_attention: the query/key width must be a multiple of 4")
    chunk = max(32, min(chunk, N))
    rpad, npad = _pad4(chunk), _pad4(N)
    out = torch.empty(B, C, N, dtype=torch.float32, device=dev)
    colsum = torch.empty(B, N, dtype=torch.float32, device=dev)
    e = torch.empty(chunk, npad, dtype=torch.float32, device=dev)                # energy rows of one chunk
    pt = torch.empty(N, rpad, dtype=torch.float32, device=dev)                   # their softmax, transposed
    vw = torch.empty(C + 1, rpad, dtype=torch.float32, device=dev)               # [x_v chunk ; 1]: the last row yields the column sums
    uc = torch.empty(N, _pad4(C + 1), dtype=torch.float32, device=dev)
    stats = torch.empty(chunk, 2, dtype=torch.float32, device=dev)
    for b in range(B):
        q = x_q[b].float().contiguous()
        kt = x_k[b].float().t().contiguous()                                     # [N,c]
        u = torch.zeros(N, C + 1, dtype=torch.float32, device=dev)
        for r0 in range(0, N, chunk):
            r = min(chunk, N - r0)
            ops.linear_into(q[r0:r0 + r], kt, None, e[:r])                       # [r,N] = Q_chunk K^T
            check(lib.dvm_softmax_rows_transposed(e.data_ptr(), r, N, e.stride(0), ptr(pt), pt.stride(0), ptr(stats), stream_ptr()),
                  "dvm_softmax_rows_transposed")
            vw[:C, :r] = x_v[b, :, r0:r0 + r]
            vw[C, :r] = 1.0
            # [N,C+1] = P^T [x_v ; 1]^T over the chunk's rows (K = r; the operands' first r columns)
            check(lib.dvm_linear_act_fwd(pt.data_ptr(), N, r, pt.stride(0), vw.data_ptr(), vw.stride(0), None, C + 1, 0,
                                         uc.data_ptr(), uc.stride(0), stream_ptr()), "dvm_linear_act_fwd")
            u += uc[:, :C + 1]
        colsum[b] = u[:, C]
        out[b] = (u[:, :C] / (1e-9 + u[:, C:])).t()
    return out, colsum


def _attention_backward(x_q, x_k, x_v, x_r, colsum, g, chunk):
    """Gradients of x_r = x_v A w.r.t. (x_q, x_k, x_v) from row chunks (csrc/attention.cu has the algebra): per chunk five
    fp32-equivalent tcgen05 GEMMs (E, dA, dV, dQ, dK) around the softmax and softmax-backward kernels."""
    lib = _lib.load()
    B, N, c = x_q.shape
    C = x_v.shape[1]
    dev = x_q.device
    chunk = max(32, min(chunk, N, 65535))
    rpad, npad = _pad4(chunk), _pad4(N)
    d_q = torch.empty(B, N, c, dtype=torch.float32, device=dev)
    d_k = torch.empty(B, c, N, dtype=torch.float32, device=dev)
    d_v = torch.empty(B, C, N, dtype=torch.float32, device=dev)
    e = torch.empty(chunk, npad, dtype=torch.float32, device=dev)                # E, then P
    da = torch.empty(chunk, npad, dtype=torch.float32, device=dev)               # dA, then dE
    det = torch.empty(N, rpad, dtype=torch.float32, device=dev)                  # dE transposed
    gs = torch.zeros(C, npad, dtype=torch.float32, device=dev)                   # G with its columns scaled by t
    gt = torch.empty(N, C, dtype=torch.float32, device=dev)                      # G^T
    kp = torch.zeros(c, npad, dtype=torch.float32, device=dev)                   # x_k with a TMA-friendly row pitch
    qt = torch.empty(c, rpad, dtype=torch.float32, device=dev)                   # Q_chunk^T
    vt = torch.empty(chunk, C, dtype=torch.float32, device=dev)                  # x_v chunk^T
    o_rc = torch.empty(chunk, _pad4(max(C, c)), dtype=torch.float32, device=dev)
    o_nc = torch.empty(N, _pad4(c), dtype=torch.float32, device=dev)
    stats = torch.empty(chunk, 2, dtype=torch.float32, device=dev)
    rowdot = torch.empty(chunk, dtype=torch.float32, device=dev)
    for b in range(B):
        q = x_q[b].float().contiguous()
        kt = x_k[b].float().t().contiguous()
        kp[:, :N] = x_k[b]
        gb = g[b].float()
        t = (1.0 / (1e-9 + colsum[b])).contiguous()
        w = (gb * x_r[b]).sum(0).contiguous()
        gs[:, :N] = gb * t
        gt.copy_(gb.t())
        dk_acc = torch.zeros(N, c, dtype=torch.float32, device=dev)
        for r0 in range(0, N, chunk):
            r = min(chunk, N - r0)
            ops.linear_into(q[r0:r0 + r], kt, None, e[:r])
            check(lib.dvm_softmax_rows_inplace(e.data_ptr(), r, N, e.stride(0), ptr(stats), stream_ptr()), "dvm_softmax_rows_inplace")
            vt[:r] = x_v[b, :, r0:r0 + r].t()
            ops.linear_into(vt[:r], gt, None, da[:r])                            # dA = x_v_chunk^T G                       [r,N]
            # dV_chunk = P (G t)^T: K = N
            check(lib.dvm_linear_act_fwd(e.data_ptr(), r, N, e.stride(0), gs.data_ptr(), gs.stride(0), None, C, 0,
                                         o_rc.data_ptr(), o_rc.stride(0), stream_ptr()), "dvm_linear_act_fwd")
            d_v[b, :, r0:r0 + r] = o_rc[:r, :C].t()
            check(lib.dvm_attn_softmax_bwd(e.data_ptr(), da.data_ptr(), ptr(t), ptr(w), r, N, e.stride(0), ptr(det), det.stride(0),
                                           ptr(rowdot), stream_ptr()), "dvm_attn_softmax_bwd")
            # dQ_chunk = dE K^T: K = N
            check(lib.dvm_linear_act_fwd(da.data_ptr(), r, N, da.stride(0), kp.data_ptr(), kp.stride(0), None, c, 0,
                                         o_rc.data_ptr(), o_rc.stride(0), stream_ptr()), "dvm_linear_act_fwd")
            d_q[b, r0:r0 + r] = o_rc[:r, :c]
            # dK += dE^T Q_chunk: K = r
            qt[:, :r] = q[r0:r0 + r].t()
            check(lib.dvm_linear_act_fwd(det.data_ptr(), N, r, det.stride(0), qt.data_ptr(), qt.stride(0), None, c, 0,
                                         o_nc.data_ptr(), o_nc.stride(0), stream_ptr()), "dvm_linear_act_fwd")
            dk_acc += o_nc[:, :c]
        d_k[b] = dk_acc.t()
    return d_q, d_k, d_v


class _SAAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_q, x_k, x_v, chunk):
        out, colsum = _attention_forward(x_q.detach(), x_k.detach(), x_v.detach(), chunk)
        ctx.save_for_backward(x_q, x_k, x_v, out, colsum)
        ctx.chunk = chunk
        return out

    @staticmethod
    def backward(ctx, g):
        x_q, x_k, x_v, out, colsum = ctx.saved_tensors
        d_q, d_k, d_v = _attention_backward(x_q.detach(), x_k.detach(), x_v.detach(), out, colsum, g.contiguous(), ctx.chunk)
        return d_q.to(x_q.dtype), d_k.to(x_k.dtype), d_v.to(x_v.dtype), None


def sa_attention(x_q, x_k, x_v, chunk=2048):
    """x_r = x_v @ A,  A = softmax(x_q @ x_k, dim=-1) / (1e-9 + column sums)   (models/model.py:116-119), forward and backward,
    without the N x N matrices.

    x_q [B,N,c], x_k [B,c,N], x_v [B,C,N]  ->  [B,C,N] float32."""
    return _SAAttention.apply(x_q, x_k, x_v, chunk)


def sa_layer_forward(self, x):
    """Drop-in for `SA_Layer.forward` (models/model.py:113-123): same modules, same order; the bmm -> softmax -> renormalise ->
    bmm core runs through `sa_attention` (autograd-aware) on CUDA tensors."""
    x_q = self.q_conv(x).permute(0, 2, 1)          # b, n, c
    x_k = self.k_conv(x)                           # b, c, n
    x_v = self.v_conv(x)
    if not x.is_cuda:                              # the reference's dense formula
        energy = torch.bmm(x_q, x_k)
        attention = self.softmax(energy)
        attention = attention / (1e-9 + attention.sum(dim=1, keepdims=True))
        x_r = torch.bmm(x_v, attention)
    else:
        x_r = sa_attention(x_q, x_k, x_v)
    x_r = self.act(self.after_norm(self.trans_conv(x - x_r)))
    return x + x_r


__all__ = ["sa_attention", "sa_layer_forward"]
