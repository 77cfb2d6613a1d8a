"""Development aid: run the kernel-free GPU-marked tests (weight files, config files) against the host stub.
Usage: python tools/hoststub/run_host_tests.py [extra pytest args]"""
import ctypes as C
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from bcnn_b200 import capi  # noqa: E402

stub = C.CDLL(str(Path(__file__).parent / "libbcnn_hoststub.so"), mode=C.RTLD_LOCAL)
capi.bind_bcnn_api(stub, capi.TensorB200)
capi.bind_b200_ext(stub)
capi._B200 = stub
sys.exit(pytest.main([str(ROOT / "tests" / "test_model_io.py"), str(ROOT / "tests" / "test_cfg.py"),
                      "-q", "-m", "gpu", "-k",
                      "layout or save_writes or load_train or load_darknet or error_statuses or "
                      "darknet_dialect_builds or loss_on_the_reference_heads",
                      *sys.argv[1:]]))
