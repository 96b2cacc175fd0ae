"""ctypes binding of libdvm_b200.so (the C ABI declared in include/dvm_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every call passes raw
device pointers.  There is NO fallback: if the library is missing or the device is not sm_100 the
import-time / call-time error is loud.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVM_LIB_PATH: A/B experiments against another build of the same C ABI (tools/); symbols it lacks are skipped
LIB_PATH = os.environ.get("DVM_LIB_PATH") or os.path.join(_HERE, "libdvm_b200.so")

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/dvm_b200.h declares
SIGNATURES = {
    "dvm_version": (c_int, []),
    "dvm_last_error_string": (ctypes.c_char_p, []),
    "dvm_device_check": (c_int, []),
    "dvm_launch_count": (ctypes.c_longlong, []),
    "dvm_profile_enable": (c_int, [c_int]),
    "dvm_profile_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]),
    "dvm_profile_read_channel": (c_int, [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_int)]),
    "dvm_softmap_workspace_bytes": (c_size_t, [c_int] * 5),
    "dvm_softmap_fwd": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_float] + [c_int] * 3 + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p]),
    "dvm_softmap_bwd_workspace_bytes": (c_size_t, [c_int] * 4),
    "dvm_softmap_bwd": (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_float, c_int] + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p]),
    "dvm_softmap_bwd_tc_workspace_bytes": (c_size_t, [c_int] * 4),
    "dvm_softmap_bwd_tc": (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_float, c_int] + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p]),
    "dvm_sparse_transfer_fwd": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p, c_void_p]),
    "dvm_sparse_transfer_bwd": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 3),
    "dvm_knn3_workspace_bytes": (c_size_t, [c_int] * 3),
    "dvm_knn3": (c_int, [c_void_p] * 2 + [c_int] * 5 + [c_void_p] * 4 + [c_void_p, c_size_t, c_void_p]),
    "dvm_chamfer_workspace_bytes": (c_size_t, [c_int] * 3),
    "dvm_chamfer_fwd": (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_void_p] * 4 + [c_void_p, c_size_t, c_void_p]),
    "dvm_chamfer_bwd": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p] * 3),
    "dvm_fps_workspace_bytes": (c_size_t, [c_int] * 2),
    "dvm_fps": (c_int, [c_void_p] + [c_int] * 3 + [c_void_p] * 2 + [c_void_p, c_size_t, c_void_p]),
    "dvm_graph_workspace_bytes": (c_size_t, [c_int] * 3),
    "dvm_graph_weights": (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_void_p] * 5 + [c_void_p, c_size_t, c_void_p]),
    "dvm_rot6d_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "dvm_rot6d_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dvm_skin_fwd": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p] * 2),
    "dvm_skin_bwd": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 3),
    "dvm_arap_workspace_bytes": (c_size_t, [c_int] * 2),
    "dvm_arap_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 2 + [c_void_p, c_size_t, c_void_p]),
    "dvm_arap_bwd": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p] * 3),
    "dvm_node_table": (c_int, [c_void_p] * 4 + [c_int] * 2 + [c_void_p] * 2),
    "dvm_node_table_from_d9": (c_int, [c_void_p] * 3 + [c_int] * 2 + [c_void_p] * 4),
    "dvm_skin_fwd_packed": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 2),
    "dvm_skin_bwd_csr": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p] * 3),
    "dvm_arap_packed_workspace_bytes": (c_size_t, [c_int] * 2),
    "dvm_arap_fwd_packed": (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_void_p] * 2 + [c_void_p, c_size_t, c_void_p]),
    "dvm_gather_conv_fwd": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 2),
    "dvm_gather_conv_bwd": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 4),
    "dvm_topk_select": (c_int, [c_void_p, ctypes.c_longlong, c_int, ctypes.c_longlong, c_int, c_void_p, c_void_p]),
    "dvm_pair_dist_fwd": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dvm_pair_dist_bwd": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p, c_void_p]),
    "dvm_gather_rows_fwd": (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_void_p, c_void_p]),
    "dvm_gather_rows_bwd": (c_int, [c_void_p] * 2 + [c_int] * 4 + [c_void_p, c_void_p]),
    "dvm_softmax_rows_transposed": (c_int, [c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    "dvm_softmax_rows_inplace": (c_int, [c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p]),
    "dvm_attn_softmax_bwd": (c_int, [c_void_p] * 4 + [c_int, c_int, ctypes.c_longlong, c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    "dvm_linear_act_fwd": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
}

PREC = {"fp32": 0, "f16": 1, "bf16": 2}
MODE_HARD, MODE_SOFT = 0, 1

_lib = None
_lock = threading.Lock()
_device_ok = set()


def load():
    """dlopen the in-tree library (built by `python -m dv_matcher_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -m dv_matcher_b200.build` "
                    "(nvcc, sm_100a). dv_matcher_b200 has no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                if os.environ.get("DVM_LIB_PATH") and not hasattr(lib, name):
                    continue
                fn = getattr(lib, name)      # AttributeError here == header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().dvm_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t, dtype=None):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL); `dtype` asserts what the kernel will read."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise RuntimeError("dv_matcher_b200 kernels need contiguous CUDA tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"dv_matcher_b200: expected a {dtype} tensor, got {t.dtype}")
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_device(t):
    """The product path runs on a B200 or not at all."""
    if not t.is_cuda:
        raise RuntimeError("dv_matcher_b200: tensors must live on a CUDA (sm_100) device; there is no CPU path")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev != torch.cuda.current_device():
        # the library launches on the CURRENT device and stream (one process per GPU): refuse instead of launching elsewhere
        raise RuntimeError(f"dv_matcher_b200: tensor lives on cuda:{dev} but the current device is cuda:{torch.cuda.current_device()}; "
                           "call torch.cuda.set_device first (one process per GPU)")
    if dev not in _device_ok:
        with torch.cuda.device(dev):
            check(load().dvm_device_check(), "dvm_device_check")
        _device_ok.add(dev)
    return dev


class _Workspace:
    """Grow-only per-(device, stream) scratch buffers, so the hot loop never allocates."""

    def __init__(self):
        self.buf = {}
        self.keep_retired = False      # set once a CUDA graph has been captured: its kernels hold raw workspace addresses
        self.retired = []

    def get(self, nbytes, device, tag="default"):
        key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None and self.keep_retired:
                self.retired.append(b)
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


workspace = _Workspace()


def f32c(t):
    """float32 + contiguous view/copy (the reference calls .float() on every input)."""
    t = t.float()
    return t if t.is_contiguous() else t.contiguous()
