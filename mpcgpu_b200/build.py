"""In-tree build of libgbdpcg.so (the CUDA kernels + C ABI) with nvcc for sm_100a.

The kernels are instantiated in several translation units (csrc/variants_*.cu) so that they compile in parallel; each
object is rebuilt only when its source or a header under include/gbd changed."""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SRC = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
DEPS = [os.path.join(ROOT, "include", "gbd", f) for f in os.listdir(os.path.join(ROOT, "include", "gbd"))] + [
    os.path.join(ROOT, "include", "gbd_pcg.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
LIB = os.path.join(HERE, "lib", "libgbdpcg.so")
OBJ = os.path.join(HERE, "lib", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-diag-suppress", "186"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in sources if os.path.exists(p))


def stale() -> bool:
    return _newer(LIB, SRC + DEPS)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not (force or stale()):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for src in SRC:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _newer(obj, [src] + DEPS):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for rc in ex.map(subprocess.call, jobs):
            if rc:
                raise subprocess.CalledProcessError(rc, "nvcc -c")
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in SRC]
    subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build_lib(force=True, verbose=True))
