"""LG-Net's N x N pieces on this library.  LG-Net itself (`Uni3FC`) stays the reference's PyTorch code;
what it spends its time and memory on at large N is

  * `knn_new(pcd, pcd, 40)` x 7 per forward (models/model.py:267-278)      -> `geometry.knn` (tensor-core scores + radix selection)
  * `SA_Layer` global attention x 4 per forward (models/model.py:113-123)  -> `sa_attention` below

`sa_attention` never materialises the N x N energy / attention matrices (400 MB each at N = 10k, 10 GB at 50k): it walks row
chunks with two fp32-equivalent tcgen05 GEMMs (`dvm_linear_act_fwd`, 3xTF32) around one fused row-softmax + transpose kernel
(`dvm_softmax_rows_transposed`, csrc/attention.cu); the column renormalisation `attention / (1e-9 + attention.sum(dim=1))` needs the
column sums of the whole matrix, which come out of the second GEMM through an appended row of ones.

Inference only: with gradients enabled `sa_layer_forward` runs the reference's dense formula (its backward is stock autograd).
"""
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def sa_attention(x_q, x_k, x_v, chunk=2048):
    """x_r = x_v @ A,  A = softmax(x_q @ x_k, dim=-1) / (1e-9 + column sums)   (models/model.py:116-119).

    x_q [B,N,c], x_k [B,c,N], x_v [B,C,N]  ->  [B,C,N] float32."""
    lib = _lib.load()
    B, N, c = x_q.shape
    C = x_v.shape[1]
    dev = x_q.device
    if c % 4:
        raise RuntimeError("sa_attention: the query/key width must be a multiple of 4")
    chunk = max(32, min(chunk, N))
    rpad, npad = (chunk + 3) // 4 * 4, (N + 3) // 4 * 4
    out = torch.empty(B, C, N, dtype=torch.float32, device=dev)
    e = torch.empty(chunk, npad, dtype=torch.float32, device=dev)                # energy rows of one chunk
    pt = torch.empty(N, rpad, dtype=torch.float32, device=dev)                   # their softmax, transposed
    vw = torch.empty(C + 1, rpad, dtype=torch.float32, device=dev)               # [x_v chunk ; 1]: the last row yields the column sums
    uc = torch.empty(N, (C + 4) // 4 * 4, dtype=torch.float32, device=dev)
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
        out[b] = (u[:, :C] / (1e-9 + u[:, C:])).t()
    return out


def sa_layer_forward(self, x):
    """Drop-in for `SA_Layer.forward` (models/model.py:113-123): same modules, same order; the bmm -> softmax -> renormalise ->
    bmm core runs through `sa_attention` when no gradient is required."""
    x_q = self.q_conv(x).permute(0, 2, 1)          # b, n, c
    x_k = self.k_conv(x)                           # b, c, n
    x_v = self.v_conv(x)
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())) or not x.is_cuda:
        energy = torch.bmm(x_q, x_k)
        attention = self.softmax(energy)
        attention = attention / (1e-9 + attention.sum(dim=1, keepdims=True))
        x_r = torch.bmm(x_v, attention)
    else:
        x_r = sa_attention(x_q, x_k, x_v)
    x_r = self.act(self.after_norm(self.trans_conv(x - x_r)))
    return x + x_r


__all__ = ["sa_attention", "sa_layer_forward"]
