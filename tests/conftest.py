"""pytest configuration: markers, and the CPU oracle (test infrastructure)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load_oracle():
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    try:
        import _monte_oracle  # noqa: F401
    except ImportError:
        subprocess.check_call(["make", "-C", odir, "-j2"], stdout=subprocess.DEVNULL)
        import _monte_oracle  # noqa: F401
    return sys.modules["_monte_oracle"]


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference path (oracle/monte_oracle.hh)."""
    return _load_oracle()


def have_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
