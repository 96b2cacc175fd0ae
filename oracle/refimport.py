"""Import the *unmodified* reference modules on CPU (build container only).

TEST INFRASTRUCTURE.  Used by `tests/golden/make_golden.py` and by the optional
`reference-live` tests, which skip when /root/reference is absent (it is absent
on the GPU box).  Mechanism: a `sys.meta_path` finder hands out
`MagicMock` modules for the third-party packages the reference imports but the
image lacks; none of them is reached by the hot-path functions we call.
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("DVM_REFERENCE_ROOT", "/root/reference")

_STUBS = (
    "matplotlib", "tensorboardX", "timm", "torch_scatter", "pytorch3d", "potpourri3d", "open3d",
    "psbody", "trimesh", "torchmetrics", "pytorch_lightning", "torch_geometric", "knn_cuda",
    "featup", "ChamferDistancePytorch", "emd", "igl", "robust_laplacian",
)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        if spec.name == "pytorch_lightning":
            import torch.nn as nn
            m.LightningModule = nn.Module  # `class norm(pl.LightningModule)`, models/loss.py:696
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    sys.meta_path.insert(0, _StubFinder())
    sys.path[:0] = [REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "misc")]
    _installed = True


def modules():
    """Returns (models.loss, lib.deformation_graph_point, models.model, lib.deformation_graph)."""
    install()
    import models.loss as ref_loss
    import lib.deformation_graph_point as ref_dg
    import models.model as ref_model
    import lib.deformation_graph as ref_dg_aa
    return ref_loss, ref_dg, ref_model, ref_dg_aa


def load_off_vertices(path):
    """Vertices of an ASCII OFF file (same parse as models/dataset.py:16-27)."""
    import numpy as np
    with open(path, "r") as f:
        f.readline()
        n, _, _ = map(int, f.readline().split())
        pts = [list(map(float, f.readline().split()[:3])) for _ in range(n)]
    return np.asarray(pts, dtype=np.float32)
