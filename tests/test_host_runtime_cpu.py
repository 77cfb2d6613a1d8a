"""CPU coverage of the C host runtime (bcnn_b200/src/**/*.c) without a GPU.

tools/hoststub/ links the REAL host objects of libbcnn_b200.so against a host-memory stand-in for
csrc/device.cu and csrc/optim.cu ("device" buffers are malloc'd, copies are memcpy, the two
optimizer kernels are restated in scalar C). That library is test infrastructure: it is never
built by build(), never loaded by bcnn_b200/ and cannot run forward or backward (every other
kernel still fails at launch). What it does allow, here and now, is driving the host logic next
to the compiled reference:
  * tools/hoststub/check_host_logic.py -- weight files, optimizer dispatch, config files, yolo
    detections and the host restatement of the yolo loss, each compared with the reference;
  * the kernel-free subset of the GPU-marked tests (tests/test_model_io.py, tests/test_cfg.py)
    with the stub bound in place of the CUDA library.
"""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

from helpers import ROOT, ref_available

STUB_DIR = ROOT / "tools" / "hoststub"
OBJECTS = ROOT / "bcnn_b200" / "_build"

pytestmark = [
    pytest.mark.skipif(not ref_available(), reason="oracle/_ref was not built / did not travel"),
    pytest.mark.skipif(shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists(),
                       reason="no nvcc to link the host stub"),
    pytest.mark.skipif(not any(OBJECTS.glob("*.c.o")), reason="bcnn_b200 objects not built"),
]


@pytest.fixture(scope="module")
def stub():
    proc = subprocess.run(["bash", str(STUB_DIR / "build.sh")], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    return STUB_DIR / "libbcnn_hoststub.so"


def test_host_logic_matches_the_reference(stub):
    proc = subprocess.run([sys.executable, str(STUB_DIR / "check_host_logic.py")],
                          capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    for line in ("save: byte-identical", "PREDICT fold bit-identical", "Darknet files",
                 "update[adam]", "learning-rate policies", "load_net: graphs", "yolo detections", "yolo loss (host loops)", "OK"):
        assert line in proc.stdout, line


def test_kernel_free_gpu_tests_pass_on_the_stub(stub):
    proc = subprocess.run([sys.executable, str(STUB_DIR / "run_host_tests.py")],
                          capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-2000:]
    assert " passed" in proc.stdout and "failed" not in proc.stdout
