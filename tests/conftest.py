import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_lib():
    """libcdfgpu through its C ABI, initialised on cuda:0.  Fails (does not skip) when the library is missing."""
    from cdftools_b200 import lib
    lib.load()
    lib.init(0, 3)
    yield lib
    lib.finalize()
