"""Builds the native library IN-TREE (gam_ngs_b200/libgamx.so) for sm_100a.

The built .so is git-ignored but travels with gpurun snapshots, so the GPU box never
needs to compile.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgamx.so")
SOURCES = [os.path.join(CSRC, "gamx.cu")]
HEADERS = [os.path.join(CSRC, f) for f in
           ("bsw_common.h", "bsw_warp.h", "bsw_warp16.h", "bsw_generic.h", "bsw_traceback.h", "bsw_host.h", "merge_collector.h")] + \
          [os.path.join(os.path.dirname(HERE), "include", "gamx.h")]

NVCC_FLAGS = ["-split-compile", "0", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    few = ["-DGAMX_DEV_FEW_KERNELS"] if os.environ.get("GAMX_BUILD_FEW") else []  # development: a few kernel geometries only
    few += os.environ.get("GAMX_BUILD_DEFS", "").split()  # experiments: extra -D flags ...
    out = os.environ.get("GAMX_BUILD_OUT", LIB)            # ... into another file (load it with GAMX_LIB)
    cmd = [nvcc] + NVCC_FLAGS + few + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
