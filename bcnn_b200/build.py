"""Build libbcnn_b200.so in-tree: hand-written sm_100a CUDA kernels (csrc/*.cu) plus the
C99 host side (src/**/*.c) that mirrors bcnn's net / node / tensor / layer interface.

    python -m bcnn_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored). No cuBLAS / cuDNN / NCCL is
linked: NCCL is bound at run time with dlopen (src/bcnn_dp.c).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
BUILD = PKG / "_build"
LIB = PKG / "libbcnn_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("BCNN_B200_CC", "/usr/bin/gcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
INCLUDES = [f"-I{ROOT / 'include'}", f"-I{PKG / 'src'}", f"-I{PKG / 'src' / 'layers'}",
            f"-I{PKG / 'csrc'}"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"] + ARCH
C_FLAGS = ["-O2", "-std=gnu99", "-fPIC", "-Wall", "-Wno-unused-function",
           "-DBCNN_USE_CUDA=1", "-I/usr/local/cuda/include"]


def _sources():
    cu = sorted((PKG / "csrc").glob("*.cu"))
    c = sorted((PKG / "src").rglob("*.c"))
    return cu, c


def _digest(path: Path, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    h.update(path.read_bytes())
    # headers are few: hash them all so any header edit rebuilds everything
    for hdr in sorted(list((ROOT / "include").rglob("*.h")) + list(PKG.rglob("*.cuh"))
                      + list((PKG / "src").rglob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()[:16]


def _compile(job):
    src, obj, cmd, stamp, digest, verbose = job
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return src, 0, None
    proc = subprocess.run(cmd, capture_output=True, text=True)
    out = proc.stdout + proc.stderr
    if proc.returncode == 0:
        stamp.write_text(digest)
        (obj.with_suffix(".log")).write_text(out)
    return src, proc.returncode, out


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    cu, c = _sources()
    jobs = []
    for src in cu:
        obj = BUILD / (src.stem + ".cu.o")
        cmd = [NVCC, *NVCC_FLAGS, *INCLUDES, "-c", str(src), "-o", str(obj)]
        jobs.append((src, obj, cmd, obj.with_suffix(".stamp"),
                     _digest(src, NVCC_FLAGS), verbose))
    for src in c:
        obj = BUILD / (src.stem + ".c.o")
        cmd = [CC, *C_FLAGS, *INCLUDES, "-c", str(src), "-o", str(obj)]
        jobs.append((src, obj, cmd, obj.with_suffix(".stamp"), _digest(src, C_FLAGS), verbose))
    if force:
        for j in jobs:
            if j[3].exists():
                j[3].unlink()
    failed = False
    rebuilt = 0
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        for src, rc, out in pool.map(_compile, jobs):
            if rc != 0:
                failed = True
                sys.stderr.write(f"[build] FAILED {src}\n{out}\n")
            elif out is not None:
                rebuilt += 1
                if verbose:
                    sys.stderr.write(f"[build] {src.name}\n{out}\n")
    if failed:
        raise RuntimeError("bcnn_b200 build failed")
    objs = [str(j[1]) for j in jobs]
    if rebuilt or not LIB.exists() or force:
        link = [NVCC, "-shared", *ARCH, "-o", str(LIB), *objs, "-Xcompiler", "-fPIC",
                "--cudart=static", "-Xlinker", "-Bsymbolic", "-ldl", "-lm", "-lpthread", "-lrt"]
        proc = subprocess.run(link, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("bcnn_b200 link failed")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
