import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True, scope="module")
def _fresh_default_rng():
    """Seed the default torch generators (CPU and CUDA) at the start of every test module: this torch build seeds them from
    entropy per process, so tests that draw from the default generator would otherwise see different inputs on every run and
    depend on which modules ran before them in the same process."""
    import torch
    torch.manual_seed(67280421310721)
