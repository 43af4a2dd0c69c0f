import os
import sys
from pathlib import Path

import pytest

# the opt-in experiment kernels (fvk_set_variant 1-4, 6) stay under parity test: their plans are only built on request
os.environ.setdefault("FVK_EXPERIMENT_PLANS", "1")

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
