"""pytest configuration: registers the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` tests run on the CPU-only build container (oracle vs goldens, host
logic, C-ABI symbol checks); ``-m gpu`` tests are the parity tests proper and call
the CUDA path through the C-ABI on a real B200.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(size: str):
    return np.load(os.path.join(GOLDEN_DIR, f"{size}.npz"))


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint16).astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="session")
def golden():
    return load_golden
