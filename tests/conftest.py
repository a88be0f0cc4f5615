import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ref():
    """The reference's own kernels (oracle/_ref, built by oracle/build_ref.sh): checker only."""
    from oracle import ref_loader
    mods = ref_loader.load()
    if mods is None:
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    return mods


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
