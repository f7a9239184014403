import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    from cucollections_b200 import _cabi
    return _cabi.native()


@pytest.fixture(scope="session")
def reference_lib():
    """cuco's own build of the shim; absent -> tests that need it are skipped."""
    from cucollections_b200 import _cabi
    try:
        return _cabi.reference()
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libcuco_ref.so not built")
