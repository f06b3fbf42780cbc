"""ctypes binding of the C oracle + fp64 ground truth.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpcg_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile pcg_oracle.c (gcc only; a second or two)."""
    src = os.path.join(_HERE, "pcg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        u32p, u8p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
        _lib.glass_reduce_f32.restype = C.c_float
        _lib.glass_reduce_f32.argtypes = [C.c_uint32, f32p]
        _lib.glass_dot_f32.restype = C.c_float
        _lib.glass_dot_f32.argtypes = [C.c_uint32, f32p, f32p, f32p]
        _lib.glass_reduce_f64.restype = C.c_double
        _lib.glass_reduce_f64.argtypes = [C.c_uint32, f64p]
        _lib.bdmv_f32.restype = None
        _lib.bdmv_f32.argtypes = [C.c_uint32, C.c_uint32, f32p, f32p, f32p, C.c_int]
        _lib.bdmv_f64.restype = None
        _lib.bdmv_f64.argtypes = [C.c_uint32, C.c_uint32, f64p, f64p, f64p, C.c_int]
        _lib.pcg_oracle_f32.restype = C.c_int
        _lib.pcg_oracle_f32.argtypes = [C.c_uint32, C.c_uint32, f32p, f32p, f32p, f32p, C.c_uint32, C.c_float,
                                        C.c_int, u32p, u8p, f32p, f32p, f32p]
        _lib.pcg_oracle_f64.restype = C.c_int
        _lib.pcg_oracle_f64.argtypes = [C.c_uint32, C.c_uint32, f64p, f64p, f64p, f64p, C.c_uint32, C.c_double,
                                        C.c_int, u32p, u8p, f64p, f64p, f64p]
        _lib.pcg_oracle_batched_f32.restype = C.c_int
        _lib.pcg_oracle_batched_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p, f32p, f32p,
                                                C.c_uint32, C.c_float, C.c_int, u32p, u8p]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def glass_reduce(x: np.ndarray):
    """GLASS/src/L1/reduce.cuh:5-33 on a copy of x (fp32 or fp64)."""
    x = np.ascontiguousarray(x).copy()
    if x.dtype == np.float32:
        return np.float32(lib().glass_reduce_f32(x.size, _p(x, C.c_float)))
    x = x.astype(np.float64)
    return np.float64(lib().glass_reduce_f64(x.size, _p(x, C.c_double)))


def glass_dot(x: np.ndarray, y: np.ndarray) -> np.float32:
    """GLASS/src/L1/dot.cuh:52-63 (fp32)."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    s = np.empty_like(x)
    return np.float32(lib().glass_dot_f32(x.size, _p(x, C.c_float), _p(y, C.c_float), _p(s, C.c_float)))


def bdmv(M: np.ndarray, x: np.ndarray, n: int, N: int, contract: bool = True) -> np.ndarray:
    """Band matvec over the [N][3][n][n] column-major-tile layout (utils.cuh:46-85)."""
    dt = M.dtype
    M = np.ascontiguousarray(M).reshape(-1)
    x = np.ascontiguousarray(x, dt).reshape(-1)
    y = np.empty(n * N, dt)
    if dt == np.float32:
        lib().bdmv_f32(n, N, _p(M, C.c_float), _p(x, C.c_float), _p(y, C.c_float), int(contract))
    else:
        lib().bdmv_f64(n, N, _p(M, C.c_double), _p(x, C.c_double), _p(y, C.c_double), int(contract))
    return y


def pcg(S, Pinv, gamma, lambda0, n: int, N: int, max_iter: int, exit_tol: float, contract: bool = True):
    """Reference-order PCG (pcg.cuh:98-217).  Returns dict(lam, iters, max_iter_exit, r, p, eta)."""
    dt = np.asarray(S).dtype
    assert dt in (np.float32, np.float64)
    ct = C.c_float if dt == np.float32 else C.c_double
    S = np.ascontiguousarray(S, dt).reshape(-1)
    Pinv = np.ascontiguousarray(Pinv, dt).reshape(-1)
    gamma = np.ascontiguousarray(gamma, dt).reshape(-1)
    lam = np.ascontiguousarray(lambda0, dt).reshape(-1).copy()
    assert S.size == 3 * n * n * N and Pinv.size == S.size and gamma.size == n * N and lam.size == n * N
    r = np.empty(n * N, dt)
    p = np.empty(n * N, dt)
    iters, flag, eta = C.c_uint32(0), C.c_uint8(0), ct(0)
    fn = lib().pcg_oracle_f32 if dt == np.float32 else lib().pcg_oracle_f64
    rc = fn(n, N, _p(S, ct), _p(Pinv, ct), _p(gamma, ct), _p(lam, ct), max_iter, exit_tol, int(contract),
            C.byref(iters), C.byref(flag), _p(r, ct), _p(p, ct), C.byref(eta))
    if rc:
        raise ValueError(f"pcg_oracle rc={rc}")
    return dict(lam=lam, iters=int(iters.value), max_iter_exit=bool(flag.value), r=r, p=p, eta=float(eta.value))


_FAST_PATH = os.path.join(_HERE, "_build", "libpcg_fast_oracle.so")
_fast = None


def fast_lib():
    """pcg_fast_oracle.c: the tolerance-parity kernels' own operation order, restated on the CPU."""
    global _fast
    if _fast is None:
        src = os.path.join(_HERE, "pcg_fast_oracle.c")
        if not os.path.exists(_FAST_PATH) or os.path.getmtime(_FAST_PATH) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
        _fast = C.CDLL(_FAST_PATH)
        f32p = C.POINTER(C.c_float)
        _fast.pcg_fast_oracle_g_f32.restype = C.c_int
        _fast.pcg_fast_oracle_g_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p, f32p, f32p, C.c_uint32, C.c_float,
                                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), f32p, f32p, f32p]
    return _fast


def pcg_fast(S, Pinv, gamma, lambda0, n: int, N: int, cluster: int, max_iter: int, exit_tol: float, lanes: int = 16):
    """The fast kernels' arithmetic (Chronopoulos-Gear recurrence, per-CTA reductions for a cluster of `cluster` CTAs whose knot
    rows occupy `lanes` lanes each: 16, or n for the packed kernels; 0 = the reduction order of the batch kernel
    include/gbd/gbd_cluster_pcg_fastb.cuh), include/gbd/gbd_cluster_pcg_fast.cuh.
    Same return dict as pcg()."""
    S = np.ascontiguousarray(S, np.float32).reshape(-1)
    Pinv = np.ascontiguousarray(Pinv, np.float32).reshape(-1)
    gamma = np.ascontiguousarray(gamma, np.float32).reshape(-1)
    lam = np.ascontiguousarray(lambda0, np.float32).reshape(-1).copy()
    assert S.size == 3 * n * n * N and Pinv.size == S.size and gamma.size == n * N and lam.size == n * N
    r = np.empty(n * N, np.float32)
    p = np.empty(n * N, np.float32)
    iters, flag, eta = C.c_uint32(0), C.c_uint8(0), C.c_float(0)
    rc = fast_lib().pcg_fast_oracle_g_f32(n, N, cluster, lanes, _p(S, C.c_float), _p(Pinv, C.c_float), _p(gamma, C.c_float),
                                          _p(lam, C.c_float), max_iter, exit_tol, C.byref(iters), C.byref(flag),
                                          _p(r, C.c_float), _p(p, C.c_float), C.byref(eta))
    if rc:
        raise ValueError(f"pcg_fast_oracle rc={rc}")
    return dict(lam=lam, iters=int(iters.value), max_iter_exit=bool(flag.value), r=r, p=p, eta=float(eta.value))


def pcg_batched(S, Pinv, gamma, lambda0, n, N, batch, max_iter, exit_tol, contract=True):
    S = np.ascontiguousarray(S, np.float32).reshape(-1)
    Pinv = np.ascontiguousarray(Pinv, np.float32).reshape(-1)
    gamma = np.ascontiguousarray(gamma, np.float32).reshape(-1)
    lam = np.ascontiguousarray(lambda0, np.float32).reshape(-1).copy()
    iters = np.zeros(batch, np.uint32)
    flags = np.zeros(batch, np.uint8)
    rc = lib().pcg_oracle_batched_f32(n, N, batch, _p(S, C.c_float), _p(Pinv, C.c_float), _p(gamma, C.c_float),
                                      _p(lam, C.c_float), max_iter, exit_tol, int(contract),
                                      _p(iters, C.c_uint32), _p(flags, C.c_uint8))
    if rc:
        raise ValueError(f"pcg_oracle_batched rc={rc}")
    return dict(lam=lam.reshape(batch, N * n), iters=iters, max_iter_exit=flags.astype(bool))


# ---------------------------------------------------------------------------------------------
# fp64 ground truth: direct solve of the block-tridiagonal system held in the [L|D|R] layout.

def band_to_dense(M, n: int, N: int) -> np.ndarray:
    """Assemble the dense (nN x nN) matrix the band layout denotes (pad tiles ignored)."""
    T = np.asarray(M, np.float64).reshape(N, 3, n, n)  # [b][t][c][r]  (column-major tiles)
    A = np.zeros((n * N, n * N))
    for b in range(N):
        for t in range(3):
            bc = b + t - 1
            if 0 <= bc < N:
                A[b * n:(b + 1) * n, bc * n:(bc + 1) * n] = T[b, t].T
    return A


def solve_f64(S, gamma, n: int, N: int) -> np.ndarray:
    """Block-Thomas elimination in fp64 (no pivoting across blocks; S is definite)."""
    T = np.asarray(S, np.float64).reshape(N, 3, n, n)
    L = [T[b, 0].T for b in range(N)]
    D = [T[b, 1].T.copy() for b in range(N)]
    R = [T[b, 2].T for b in range(N)]
    g = np.asarray(gamma, np.float64).reshape(N, n).copy()
    for b in range(1, N):
        W = np.linalg.solve(D[b - 1].T, L[b].T).T  # L_b D_{b-1}^{-1}
        D[b] = D[b] - W @ R[b - 1]
        g[b] = g[b] - W @ g[b - 1]
    x = np.zeros((N, n))
    x[N - 1] = np.linalg.solve(D[N - 1], g[N - 1])
    for b in range(N - 2, -1, -1):
        x[b] = np.linalg.solve(D[b], g[b] - R[b] @ x[b + 1])
    return x.reshape(-1)


def rel_residual(S, gamma, lam, n: int, N: int) -> float:
    """||gamma - S lam||_2 / ||gamma||_2 evaluated in fp64."""
    y = bdmv(np.asarray(S, np.float64), np.asarray(lam, np.float64), n, N, contract=True)
    g = np.asarray(gamma, np.float64).reshape(-1)
    return float(np.linalg.norm(g - y) / max(np.linalg.norm(g), 1e-300))
