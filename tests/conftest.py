import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


# The suite is a PARITY suite against the reference's CPU path, so a net made without further
# calls is the verification configuration: FP32 SIMT convolutions and the reference's residual
# semantics (H2 / H3). The library's own defaults (tensor-core math, batch-correct residuals) are
# asserted in tests/test_nets_gpu.py::test_library_defaults; tests of the tensor-core path select
# it explicitly.
os.environ.setdefault("BCNN_B200_CONV_MATH", "fp32")
os.environ.setdefault("BCNN_B200_REFERENCE_QUIRKS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 box)")


def _has_gpu() -> bool:
    try:
        from bcnn_b200 import capi
        return capi.b200().bcnn_b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
