"""A C99 program written against bcnn's public API -- the network of the reference's
examples/mnist/mnist_example.c:30-55 call for call -- is compiled with
`gcc -std=gnu99 -DBCNN_USE_CUDA` against include/bcnn/bcnn.h, linked with libbcnn_b200.so and
(on a GPU box) run for three bcnn_train_on_batch steps; it also calls the net-less entry points
through the reference's own prototypes. The CPU suite proves it compiles and links."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "c" / "mnist_example_b200.c"


def _build(out: Path) -> Path:
    exe = out / "mnist_example_b200"
    cmd = ["/usr/bin/gcc", "-std=gnu99", "-O1", "-Wall", "-Werror", "-DBCNN_USE_CUDA",
           f"-I{ROOT / 'include'}", str(SRC), f"-L{ROOT / 'bcnn_b200'}", "-lbcnn_b200", "-lm",
           f"-Wl,-rpath,{ROOT / 'bcnn_b200'}", "-o", str(exe)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_c_program_compiles_and_links_against_the_public_header(tmp_path):
    assert _build(tmp_path).exists()


@pytest.mark.gpu
def test_c_program_trains_three_steps_and_the_reference_prototypes_compute(tmp_path):
    exe = _build(tmp_path)
    proc = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    lines = proc.stdout.strip().splitlines()
    assert lines[-1] == "OK" and sum(line.startswith("step ") for line in lines) == 3, proc.stdout
