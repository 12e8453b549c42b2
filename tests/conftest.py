import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with `pytest -m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the product library, the generator and the oracle once per session (CPU only: nvcc cross-compiles)."""
    from draco_sharp_b200 import build as B
    B.build_all()
    B.build_oracle()
    yield


@pytest.fixture(scope="session")
def gpu_decoder():
    import draco_sharp_b200 as D
    dec = D.DracoBatchDecoder()  # raises DCB_ERR_NO_DEVICE without a B200: no fallback
    yield dec
    dec.close()
