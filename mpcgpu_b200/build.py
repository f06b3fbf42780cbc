"""In-tree build of libgbdpcg.so (the CUDA kernels + C ABI) with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "gbd_capi.cu")]
DEPS = [os.path.join(ROOT, "include", "gbd", f) for f in os.listdir(os.path.join(ROOT, "include", "gbd"))] + [
    os.path.join(ROOT, "include", "gbd_pcg.h")]
LIB = os.path.join(HERE, "lib", "libgbdpcg.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "186"]


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SRC + DEPS if os.path.exists(p))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRC
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
