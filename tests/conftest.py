import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_maps():
    z = np.load(os.path.join(GOLDEN, "ref_maps.npz"))
    d = {k: z[k] for k in z.files}
    q = float(d["qstep"])
    d["feat1"] = torch.from_numpy(d["feat1_q"].astype(np.float32) * np.float32(q))
    d["feat2"] = torch.from_numpy(d["feat2_q"].astype(np.float32) * np.float32(q))
    return d


@pytest.fixture(scope="session")
def golden_graph():
    z = np.load(os.path.join(GOLDEN, "ref_graph.npz"))
    return {k: z[k] for k in z.files}
