import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_index():
    """One librse handle on cuda:0 for the GPU parity tests (fails, never skips, on a GPU box)."""
    if not _has_gpu():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    from rag_search_engine_b200 import _lib
    idx = _lib.Index(0)
    yield idx
    idx.close()


@pytest.fixture()
def fresh_index():
    if not _has_gpu():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    from rag_search_engine_b200 import _lib
    idx = _lib.Index(0)
    yield idx
    idx.close()
