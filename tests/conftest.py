import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The CUDA library and the checker libraries are built in-tree (nvcc cross-compiles without a GPU)."""
    from autofrk_b200 import build

    if not build.LIB.exists() or os.environ.get("FRK_REBUILD"):
        build.build_cuda()
    from oracle import build_oracle as ob

    ob.build_oracle()
    ob.build_ref()          # the reference's own translation unit (built where /root/reference exists; travels prebuilt)
    yield


def pytest_sessionfinish(session, exitstatus):
    """observed errors of the parity checks (tests/util.py:within) -> gpurun_out/observed_errors.json"""
    import json

    try:
        from util import OBSERVED
    except Exception:
        return
    if OBSERVED:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / "observed_errors.json").write_text(json.dumps(
            {k: {"observed": v[0], "bound": v[1]} for k, v in sorted(OBSERVED.items())}, indent=1))
