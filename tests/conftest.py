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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def load_params(name, dtype=torch.float32, device="cpu"):
    return {k: torch.from_numpy(v).to(dtype).to(device) for k, v in load_golden(name).items()}


@pytest.fixture(scope="session")
def graph_params():
    return load_params("weights_graph_ep960.npz")


@pytest.fixture(scope="session")
def classic_params():
    return load_params("weights_classic_ep990.npz")


def rel_err(a, b):
    a = torch.as_tensor(a).detach().to(torch.float64)
    b = torch.as_tensor(b).detach().to(torch.float64)
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel(a, b, floor=0.1):
    """max |a-b| / max(|b|, floor): elementwise relative error.  The floor (default 0.1, the scale of a live
    cell's state; most of the grid is exactly/near 0) keeps near-zero entries from turning 1e-8 absolute
    rounding noise into a meaningless relative figure."""
    a = torch.as_tensor(a).detach().to(torch.float64)
    b = torch.as_tensor(b).detach().to(torch.float64)
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())
