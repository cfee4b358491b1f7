import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fab():
    """The product package (flashattention.c_b200/) under its importable alias."""
    import flashattention_c_b200

    return flashattention_c_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import fa_oracle

    return fa_oracle


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests must run the native library on a real device; anything else is a failure, not a skip."""
    import torch

    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    major, _ = torch.cuda.get_device_capability(0)
    assert major == 10, "these kernels are built for sm_100a only"
    return torch.device("cuda:0")


def seeded(shape, seed, scale=1.0):
    """Deterministic N(0, scale^2) float32 array shared by the oracle and the CUDA path."""
    import numpy as np

    return (np.random.default_rng(seed).standard_normal(shape, dtype=np.float32) * np.float32(scale)).astype(np.float32)
