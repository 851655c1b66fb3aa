import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(autouse=True)
def _arithmetic_mode(request):
    """Tests that compare the CUDA path with the CPU oracle run the kernels with ISR_FLAG_SPEC_ARITH (the oracle's
    CPU-reproducible exp / rsqrt stand-ins); every other test runs the product default, the reference's arithmetic."""
    if "oracle" not in request.fixturenames:
        yield
        return
    from instascene_b200 import rasterizer
    with rasterizer.arithmetic("spec"):
        yield
