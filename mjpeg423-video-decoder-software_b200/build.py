"""In-tree build of the native libraries (no JIT cache: the .so files travel with the snapshot).

  libmjpeg423_b200.so   CUDA kernels for sm_100a + C++ host runtime + the C-ABI (include/mjpeg423_b200.h)
  libmjpeg423_synth.so  host-only from-spec stream producer used by tests and bench

Run as `python mjpeg423-video-decoder-software_b200/build.py [--force]`, or via `__graft_entry__.build()`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_CUDA = os.path.join(PKG_DIR, "libmjpeg423_b200.so")
LIB_SYNTH = os.path.join(PKG_DIR, "libmjpeg423_synth.so")

CUDA_SOURCES = ["entropy.cu", "decode.cu", "idct_colour.cu", "encode.cu", "display.cu", "runtime.cu", "multi.cu", "cabi.cu"]
CUDA_HEADERS = ["common.cuh", "runtime.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--shared", "-cudart", "static",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in CUDA_HEADERS] + [os.path.join(ROOT, "include", "mjpeg423_b200.h")]
    if force or _stale(LIB_CUDA, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_CUDA, *srcs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(PKG_DIR, "build_ptxas.log"), "w") as f:
            f.write(log)
        if res.returncode != 0:
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed building libmjpeg423_b200.so")
        if verbose:
            sys.stderr.write(log)
    return LIB_CUDA


def build_synth(force: bool = False) -> str:
    src = os.path.join(CSRC, "synth_encoder.cpp")
    if force or _stale(LIB_SYNTH, [src]):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", LIB_SYNTH, src]
        subprocess.run(cmd, check=True)
    return LIB_SYNTH


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_synth(force)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB_CUDA, LIB_SYNTH)
