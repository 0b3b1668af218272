"""In-tree build of the sm_100a shared library (explicit nvcc, no JIT cache).

``lib/libfluid_b200.so`` is git-ignored but travels to the GPU box with the
gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libfluid_b200.so")
HARNESS_SRC = os.path.join(REPO_DIR, "harness", "fluid_harness.cpp")
HARNESS_BIN = os.path.join(REPO_DIR, "harness", "fluid_harness")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # the reference's results need FMA contraction off
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(REPO_DIR, "include", "*.h")) + [os.path.abspath(__file__)]
    if force or _stale(LIB_PATH, deps):
        os.makedirs(LIB_DIR, exist_ok=True)
        cmd = ["nvcc", *NVCC_FLAGS, *os.environ.get("FS_NVCC_EXTRA", "").split(), "-o", LIB_PATH, *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True, cwd=REPO_DIR)
    return LIB_PATH


def build_harness(force: bool = False) -> str | None:
    """Host C++ harness (harness/fluid_harness.cpp): dlopens the CUDA library and
    the reference library and swaps implementations per call."""
    if not os.path.exists(HARNESS_SRC):
        return None
    deps = [HARNESS_SRC] + glob.glob(os.path.join(REPO_DIR, "include", "*.h")) + \
        glob.glob(os.path.join(REPO_DIR, "include", "*.hpp"))
    if force or _stale(HARNESS_BIN, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off",
                        "-I", os.path.join(REPO_DIR, "include"), "-o", HARNESS_BIN, HARNESS_SRC,
                        "-ldl"], check=True, cwd=REPO_DIR)
    return HARNESS_BIN


if __name__ == "__main__":
    print(build_cuda(force=True, verbose=True))
    print(build_harness(force=True))
