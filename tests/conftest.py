import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real sm_100 (B200) GPU; run with `-m gpu` on the GPU box")


def reference_present() -> bool:
    return os.path.isdir("/root/reference/landiff")


needs_reference = pytest.mark.skipif(not reference_present(), reason="/root/reference only exists in the build container")
