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


def needed_rtol(got, want, atol):
    """Smallest rtol for which |got - want| <= atol + rtol*|want| holds elementwise."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want) - atol
    m = err > 0
    if not m.any():
        return 0.0
    return float((err[m] / np.maximum(np.abs(want[m]), 1e-300)).max())


def record_measure(key, **vals):
    """Append measured errors to gpurun_out/parity_measured.json (best effort; DESIGN.md quotes them)."""
    import json
    path = os.path.join(ROOT, "gpurun_out", "parity_measured.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cur = {}
        if os.path.exists(path):
            with open(path) as f:
                cur = json.load(f)
        cur[key] = {k: float(v) for k, v in vals.items()}
        with open(path, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    except (OSError, ValueError):
        pass


def assert_grad_close(got, want, rtol, atol, key):
    """Elementwise |got-want| <= atol + rtol*|want|; the rtol actually needed is recorded under `key`."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    need = needed_rtol(got, want, atol)
    record_measure(key, needed_rtol=need, atol=atol, bound=rtol, max_abs_want=np.abs(want).max())
    assert need <= rtol, (key, need, rtol)
