"""Build libxgating.so (the C-ABI library, include/xgating.h) with nvcc for sm_100a, in-tree.

    python -m controllable_xgating_b200.build [--force]

The .so is git-ignored but travels with gpurun snapshots; nothing is JIT-compiled at import.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libxgating.so")
SOURCES = [os.path.join(CSRC, "xg_abi.cu")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "xgating.h"))
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libxgating.so (there is no CPU fallback)")
    tmp = LIB_PATH + ".tmp"
    extra = os.environ.get("XG_EXTRA_NVCC_FLAGS", "").split()     # diagnostics builds only (e.g. -DPK_FINE_TRACE)
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-o", tmp] + SOURCES
    if verbose:
        print("[xgating build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)
