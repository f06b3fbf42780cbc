"""ctypes binding of oracle/_ref/libqdldl_ref.so -- the reference's own QDLDL (CPU baseline).
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libqdldl_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build(ref: str = "/root/reference") -> bool:
    """Compile the reference's qdldl.c where it lies (only possible where /root/reference exists)."""
    if available():
        return True
    if not os.path.exists(os.path.join(ref, "qdldl", "src", "qdldl.c")):
        return False
    subprocess.check_call(["make", "-s", "-C", _HERE, f"REF={ref}", LIB_PATH])
    return True


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
        _lib.qdldl_ref_nnz.restype = C.c_int
        _lib.qdldl_ref_nnz.argtypes = [C.c_int, C.c_int]
        _lib.qdldl_ref_float_bytes.restype = C.c_int
        _lib.qdldl_ref_pattern.restype = None
        _lib.qdldl_ref_pattern.argtypes = [C.c_int, C.c_int, ip, ip]
        _lib.qdldl_ref_values.restype = None
        _lib.qdldl_ref_values.argtypes = [C.c_int, C.c_int, fp, fp]
        _lib.qdldl_ref_create.restype = C.c_void_p
        _lib.qdldl_ref_create.argtypes = [C.c_int, C.c_int]
        _lib.qdldl_ref_destroy.restype = None
        _lib.qdldl_ref_destroy.argtypes = [C.c_void_p]
        _lib.qdldl_ref_sum_lnz.restype = C.c_int
        _lib.qdldl_ref_sum_lnz.argtypes = [C.c_void_p]
        _lib.qdldl_ref_solve.restype = C.c_int
        _lib.qdldl_ref_solve.argtypes = [C.c_void_p, fp, fp, fp]
        _lib.qdldl_ref_solve_batched.restype = C.c_double
        _lib.qdldl_ref_solve_batched.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp]
        assert _lib.qdldl_ref_float_bytes() == 4
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def nnz(n, N):
    return lib().qdldl_ref_nnz(n, N)


def pattern(n, N):
    cp = np.zeros(n * N + 1, np.int32)
    ri = np.zeros(nnz(n, N), np.int32)
    lib().qdldl_ref_pattern(n, N, cp.ctypes.data_as(C.POINTER(C.c_int)), ri.ctypes.data_as(C.POINTER(C.c_int)))
    return cp, ri


def values(S, n, N):
    """Upper-triangular CSC values of the stored band matrix (csr.cuh:10-36 semantics)."""
    S = np.ascontiguousarray(S, np.float32).reshape(-1, 3 * n * n * N)
    out = np.zeros((S.shape[0], nnz(n, N)), np.float32)
    for i in range(S.shape[0]):
        lib().qdldl_ref_values(n, N, _fp(S[i]), _fp(out[i]))
    return out


def solve(S, gamma, n, N):
    """One factor+solve (qdldl_solve_schur, include/qdldl/sqp.cuh:22-49)."""
    val = values(S, n, N)[0]
    b = np.ascontiguousarray(gamma, np.float32).reshape(-1)
    x = np.zeros_like(b)
    ws = lib().qdldl_ref_create(n, N)
    if not ws:
        raise RuntimeError("QDLDL_etree failed")
    rc = lib().qdldl_ref_solve(ws, _fp(val), _fp(b), _fp(x))
    lib().qdldl_ref_destroy(ws)
    if rc < 0:
        raise RuntimeError("QDLDL_factor failed")
    return x


def time_batched(vals, gammas, n, N, reps=1, nthreads=1):
    """Wall seconds for reps x batch factor+solve pairs over nthreads host threads; returns (seconds, x)."""
    vals = np.ascontiguousarray(vals, np.float32)
    gammas = np.ascontiguousarray(gammas, np.float32)
    batch = vals.shape[0]
    x = np.zeros_like(gammas)
    sec = lib().qdldl_ref_solve_batched(n, N, batch, reps, nthreads, _fp(vals), _fp(gammas), _fp(x))
    if sec < 0:
        raise RuntimeError("qdldl batched solve failed")
    return sec, x
