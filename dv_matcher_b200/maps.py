"""Hard / soft correspondence maps with the reference's call surface.

Mirrors models/loss.py:91-124 (`knnsearch_t`, `knnsearch_t_grad`, `search_t`), 1339-1347
(`topk_pi`) and the script-local copies in test.py:19-28, test_partial.py:170-179, deform.py:63-90.

The reference materialises Pi as a dense [B,N,M] tensor that is 99.8 % zeros after `topk_pi`.
Here `knnsearch_t_grad` returns a lazy handle and `topk_pi` turns it into a `SparseSoftMap`
(idx [B,N,10], w [B,N,10]) produced by ONE fused kernel launch sequence (dvm_softmap_fwd); every
consumer the reference has -- `torch.matmul(Pi, Y)`, `torch.einsum('bij,bjkm->bikm', Pi, Ynb)` --
is intercepted through `__torch_function__` and runs as a 10-sparse gather.  `.to_dense()` gives the
reference's dense tensor for anything else.
"""
import os

import torch

from . import ops

_PRECISION = os.environ.get("DVM_PREC", "f16")
TOPK = 10


def set_precision(prec):
    """'f16' (default; tcgen05, indices exact, soft weights within 2e-3), 'bf16' (2e-2) or 'fp32' (1e-4)."""
    global _PRECISION
    if prec not in ("f16", "bf16", "fp32"):
        raise ValueError(prec)
    _PRECISION = prec


def get_precision():
    return _PRECISION


# --------------------------------------------------------------------------------------------------
# autograd glue
# --------------------------------------------------------------------------------------------------
class _SoftMapTopK(torch.autograd.Function):
    """(x, y) -> top-k soft-map weights; gradient through the full-row softmax (dvm_softmap_bwd)."""

    @staticmethod
    def forward(ctx, x, y, alpha, topk, prec):
        out = ops.softmap_fwd(x, y, None, alpha=alpha, topk=topk, soft=True, prec=prec)
        ctx.save_for_backward(x, y, out.top_idx, out.top_w, out.top_d, out.row_min, out.row_sum)
        ctx.alpha = float(alpha)
        ctx.prec = os.environ.get("DVM_BWD_PREC", prec)          # the backward follows the forward's precision class
        ctx.mark_non_differentiable(out.top_idx, out.argmin, out.top_d, out.row_min, out.row_sum)
        return out.top_w, out.top_idx, out.argmin, out.top_d, out.row_min, out.row_sum

    @staticmethod
    def backward(ctx, d_w, *_unused):
        x, y, idx, w, d, rmin, rsum = ctx.saved_tensors
        saved = ops.SoftMapOut(None, idx, w, d, rmin, rsum, None, None)
        dx, dy = ops.softmap_bwd(x, y, ctx.alpha, saved, d_w.contiguous(), prec=ctx.prec)
        return dx, dy, None, None, None


class _SparseTransfer(torch.autograd.Function):
    """out = Pi @ Y for a k-sparse Pi given as (idx, w)."""

    @staticmethod
    def forward(ctx, idx, w, y):
        ctx.save_for_backward(idx, w, y)
        return ops.sparse_transfer_fwd(idx, w.contiguous(), y)

    @staticmethod
    def backward(ctx, d_out):
        idx, w, y = ctx.saved_tensors
        dw, dy = ops.sparse_transfer_bwd(idx, w.contiguous(), y, d_out.contiguous(),
                                         need_dw=ctx.needs_input_grad[1], need_dy=ctx.needs_input_grad[2])
        return None, dw, dy


# --------------------------------------------------------------------------------------------------
# tensor-likes
# --------------------------------------------------------------------------------------------------
def _densify(a):
    return a.to_dense() if isinstance(a, (SparseSoftMap, LazySoftMap)) else a


class _MapBase:
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in ("matmul", "bmm") and len(args) == 2 and isinstance(args[0], _MapBase) and torch.is_tensor(args[1]):
            return args[0].sparse().matmul(args[1])
        if name == "einsum" and len(args) == 3 and isinstance(args[1], _MapBase) and torch.is_tensor(args[2]) \
                and args[0].replace(" ", "") == "bij,bjkm->bikm":
            return args[1].sparse().transfer_neighborhood(args[2])
        args = tuple(_densify(a) for a in args)
        return func(*args, **{k: _densify(v) for k, v in kwargs.items()})

    def __matmul__(self, other):
        return self.sparse().matmul(other)

    # minimal tensor protocol used by callers of the reference's Pi
    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 3

    @property
    def is_cuda(self):
        return True

    def transpose(self, a, b):
        return self.to_dense().transpose(a, b)

    def contiguous(self):
        return self


class SparseSoftMap(_MapBase):
    """Top-k soft map: idx int32 [B,N,k] (ascending distance), w fp32 [B,N,k], logical shape [B,N,M]."""

    def __init__(self, idx, w, M, argmin=None, top_d=None, row_min=None, row_sum=None):
        if idx.dtype != torch.int32:
            idx = idx.to(torch.int32)
        if w.dtype != torch.float32:                      # the kernels read float32 weights: never reinterpret another dtype
            w = w.float()
        self.idx, self.w, self.M = idx, w, int(M)
        self.argmin, self.top_d, self.row_min, self.row_sum = argmin, top_d, row_min, row_sum
        self.shape = torch.Size((idx.shape[0], idx.shape[1], self.M))
        self.device, self.dtype = w.device, w.dtype

    def sparse(self):
        return self

    def matmul(self, y):
        """Pi @ Y for Y [B,M,D] (models/loss.py:1408-1409, models/model.py:471)."""
        if y.dim() != 3 or y.shape[1] != self.M:
            return torch.matmul(self.to_dense(), y)
        return _SparseTransfer.apply(self.idx, self.w, y.float().contiguous())

    def transfer_neighborhood(self, y_nb):
        """einsum('bij,bjkm->bikm', Pi, Ynb) for Ynb [B,M,k,m] (models/loss.py:1237)."""
        B, M, k, m = y_nb.shape
        return self.matmul(y_nb.reshape(B, M, k * m)).reshape(B, self.shape[1], k, m)

    def to_dense(self):
        out = torch.zeros(self.shape, dtype=self.w.dtype, device=self.w.device)
        return out.scatter(-1, self.idx.long(), self.w)

    def hard_map(self):
        """arg-min [B,N,1] int64 (== knnsearch_t of the same features), free by-product of the fused kernel."""
        return self.argmin.unsqueeze(-1)

    def __repr__(self):
        return f"SparseSoftMap(shape={tuple(self.shape)}, k={self.idx.shape[-1]}, device={self.device})"


class LazySoftMap(_MapBase):
    """What `knnsearch_t_grad` returns: softmax(-alpha * cdist(x, y)) not yet evaluated.

    `topk_pi` (the only thing the reference ever does with it) fuses softmax + top-k into the sparse
    kernel.  Any other use densifies with the reference's own formula.
    """

    def __init__(self, x, y, alpha):
        self.x, self.y, self.alpha = x, y, float(alpha)
        self.shape = torch.Size((x.shape[0], x.shape[1], y.shape[1]))
        self.device, self.dtype = x.device, torch.float32
        self._sparse = None

    def topk(self, k=TOPK, prec=None):
        if prec is None:
            # training runs the tcgen05 pass too (indices, distances, row statistics are exact fp32 whatever the candidate pass;
            # the full / partial loss tests against the unmodified reference pass with it); DVM_TRAIN_PREC=fp32 selects the
            # CUDA-core candidate pass (1e-4 parity of the soft weights themselves)
            needs_grad = torch.is_grad_enabled() and (self.x.requires_grad or self.y.requires_grad)
            prec = os.environ.get("DVM_TRAIN_PREC", _PRECISION) if needs_grad else _PRECISION
        w, idx, argmin, top_d, rmin, rsum = _SoftMapTopK.apply(self.x.float().contiguous(), self.y.float().contiguous(),
                                                              self.alpha, k, prec)
        return SparseSoftMap(idx, w, self.shape[2], argmin, top_d, rmin, rsum)

    def sparse(self):
        if self._sparse is None:
            self._sparse = self.topk()
        return self._sparse

    def to_dense(self):
        d = torch.cdist(self.x.float(), self.y.float())
        return torch.softmax(-self.alpha * d, dim=-1)


# --------------------------------------------------------------------------------------------------
# the reference's functions
# --------------------------------------------------------------------------------------------------
def knnsearch_t(x, y, prec=None):
    """Hard NN map, int64 [B,N,1], 0-based (models/loss.py:91-95)."""
    out = ops.softmap_fwd(x, y, None, topk=1, soft=False, prec=prec or _PRECISION)
    return out.argmin.unsqueeze(-1)


def search_t(A1, A2):
    """models/loss.py:121-124 (0-based)."""
    return knnsearch_t(A1, A2)


def knnsearch_t_1based(x, y):
    """The script-local variant that returns idx + 1 (test.py:19-23, test_partial.py:170-174, deform.py:86-90)."""
    return knnsearch_t(x, y) + 1


def search_t_1based(A1, A2):
    return knnsearch_t_1based(A1, A2)


def knnsearch_t_grad(x, y, alpha=100):
    """Soft map softmax(-alpha * cdist(x, y)) as a lazy handle (models/loss.py:110-114)."""
    return LazySoftMap(x, y, alpha)


def topk_pi(A, k=TOPK):
    """Keep the 10 largest entries per row, un-renormalised (models/loss.py:1339-1347)."""
    if isinstance(A, LazySoftMap):
        return A.topk(k)
    if isinstance(A, SparseSoftMap):
        return A
    vals, idx = torch.topk(A, k, dim=-1)                      # a dense tensor from elsewhere: reference semantics
    return SparseSoftMap(idx.int().contiguous(), vals.contiguous(), A.shape[-1])


def soft_map(x, y, alpha=100, v=None, k=TOPK, prec=None):
    """One-call inference form: (SparseSoftMap, Pi @ v) from a single fused pass (no autograd)."""
    out = ops.softmap_fwd(x, y, v, alpha=alpha, topk=k, soft=True, prec=prec or _PRECISION)
    return SparseSoftMap(out.top_idx, out.top_w, y.shape[1], out.argmin, out.top_d, out.row_min, out.row_sum), out.piv
