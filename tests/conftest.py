import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI shared library, built if missing (nvcc cross-compiles without a GPU)."""
    from cmpc_refseg_b200 import build, _lib
    build.build()
    return _lib.lib()
