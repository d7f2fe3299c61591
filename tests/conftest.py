import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """npz -> plain dict of numpy arrays."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def golden_lucy_names():
    return sorted(f[len("lucy_"):-4] for f in os.listdir(GOLDEN) if f.startswith("lucy_") and f.endswith(".npz"))


def golden_cfg_kwargs(G):
    kw = {}
    for k, v in G.items():
        if k.startswith("cfg_"):
            v = v.item()
            kw[k[4:]] = v
    return kw


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
