import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        from c99_vectordb_b200 import _cabi

        return _cabi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.fail("this test needs a CUDA device (run with -m 'not gpu' on CPU boxes)")
    return 0


GOLDEN = ROOT / "tests" / "golden"
