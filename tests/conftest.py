import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def symad_sd():
    from audiodec_b200 import synthetic as S
    return S.symad_state_dict(seed=0)


@pytest.fixture(scope="session")
def hifigan_sd():
    from audiodec_b200 import synthetic as S
    return S.hifigan_state_dict(seed=1)


@pytest.fixture(params=["tc", "ffma"])
def conv_path(request, monkeypatch):
    """Both conv engines behind the same C ABI: tcgen05 3xTF32 (default) and the CUDA-core FFMA kernels.  The library
    reads ADEC_CONV_PATH when a handle is created."""
    monkeypatch.setenv("ADEC_CONV_PATH", request.param)
    return request.param
