"""ctypes binding of oracle/_ref/libref_gbdpcg*.so -- the UNMODIFIED reference pcg<> kernel compiled
for sm_100a.  TEST INFRASTRUCTURE ONLY (GPU A/B parity, golden-vector minting)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB32 = os.path.join(_HERE, "_ref", "libref_gbdpcg.so")
LIB64 = os.path.join(_HERE, "_ref", "libref_gbdpcg_f64.so")
_l32 = _l64 = None

INSTANTIATED = [(2, 3), (6, 12), (14, 8), (14, 16), (14, 32), (14, 64), (14, 128), (14, 256), (14, 512), (64, 256)]


def available() -> bool:
    return os.path.exists(LIB32)


def _lib32():
    global _l32
    if _l32 is None:
        _l32 = C.CDLL(LIB32)
        vp, u32 = C.c_void_p, C.c_uint32
        _l32.ref_gbdpcg_launch_f32.restype = C.c_int
        _l32.ref_gbdpcg_launch_f32.argtypes = [u32, u32] + [vp] * 10 + [u32, C.c_float, C.c_uint, vp]
        _l32.ref_gbdpcg_linsys_window_f32.restype = C.c_double
        _l32.ref_gbdpcg_linsys_window_f32.argtypes = [u32, u32] + [vp] * 10 + [u32, C.c_float, C.c_uint,
                                                                               C.POINTER(u32), C.POINTER(C.c_uint8)]
    return _l32


def _lib64():
    global _l64
    if _l64 is None:
        _l64 = C.CDLL(LIB64)
        vp, u32 = C.c_void_p, C.c_uint32
        _l64.ref_gbdpcg_launch_f64.restype = C.c_int
        _l64.ref_gbdpcg_launch_f64.argtypes = [u32, u32] + [vp] * 10 + [u32, C.c_double, C.c_uint, vp]
    return _l64


class RefWorkspace:
    """Device scratch the reference launch site allocates (include/pcg/sqp.cuh:116-135)."""

    def __init__(self, n, N, dtype=None, device="cuda"):
        import torch
        dtype = dtype or torch.float32
        self.n, self.N = n, N
        self.r = torch.zeros(n * N, dtype=dtype, device=device)
        self.p = torch.zeros(n * N, dtype=dtype, device=device)
        self.v = torch.zeros(max(N, n), dtype=dtype, device=device)
        self.e = torch.zeros(max(N, n), dtype=dtype, device=device)
        self.iters = torch.zeros(1, dtype=torch.int32, device=device)
        self.flag = torch.zeros(1, dtype=torch.uint8, device=device)


def launch(n, N, S, Pinv, gamma, lam, ws: RefWorkspace, max_iter, tol, block=128, stream=0):
    """Asynchronous reference launch on `stream` (cudaStream_t as int)."""
    import torch
    args = [n, N] + [int(t.data_ptr()) for t in (S, Pinv, gamma, lam, ws.r, ws.p, ws.v, ws.e, ws.iters, ws.flag)]
    if S.dtype == torch.float64:
        rc = _lib64().ref_gbdpcg_launch_f64(*args, max_iter, float(tol), block, stream)
    else:
        rc = _lib32().ref_gbdpcg_launch_f32(*args, max_iter, float(tol), block, stream)
    if rc:
        raise RuntimeError(f"reference launch failed rc={rc}")


def solve(n, N, S, Pinv, gamma, lam0, max_iter, tol, block=128):
    """Run the reference kernel; returns dict(lam, iters, max_iter_exit, r, p) as torch tensors / ints."""
    import torch
    ws = RefWorkspace(n, N, S.dtype, S.device)
    lam = lam0.clone()
    launch(n, N, S, Pinv, gamma, lam, ws, max_iter, tol, block, int(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return dict(lam=lam, iters=int(ws.iters.item()), max_iter_exit=bool(ws.flag.item()), r=ws.r, p=ws.p)


def linsys_window(n, N, S, Pinv, gamma, lam, ws: RefWorkspace, max_iter, tol, block=128):
    it, fl = C.c_uint32(0), C.c_uint8(0)
    args = [n, N] + [int(t.data_ptr()) for t in (S, Pinv, gamma, lam, ws.r, ws.p, ws.v, ws.e, ws.iters, ws.flag)]
    us = _lib32().ref_gbdpcg_linsys_window_f32(*args, max_iter, float(tol), block, C.byref(it), C.byref(fl))
    if us < 0:
        raise RuntimeError(f"reference linsys window failed rc={us}")
    return int(it.value), bool(fl.value), float(us)


# ---- the reference's own Schur / preconditioner assembly and dz recovery (oracle/ref_schur_wrapper.cu)
LIB_SCHUR = os.path.join(_HERE, "_ref", "libref_schur.so")
_lschur = None


def schur_available() -> bool:
    return os.path.exists(LIB_SCHUR)


def _libschur():
    global _lschur
    if _lschur is None:
        _lschur = C.CDLL(LIB_SCHUR)
        vp, u32 = C.c_void_p, C.c_uint32
        _lschur.ref_form_schur_system_f32.restype = C.c_int
        _lschur.ref_form_schur_system_f32.argtypes = [u32, u32, u32] + [vp] * 7 + [C.c_float]
        _lschur.ref_compute_dz_f32.restype = C.c_int
        _lschur.ref_compute_dz_f32.argtypes = [u32, u32, u32] + [vp] * 5
    return _lschur


def form_schur_system(n, m, N, G, Cm, g, c, S, Pinv, gamma, rho):
    """Reference form_schur_system<float> on the default stream (cooperative launch); G is overwritten with the inverses."""
    rc = _libschur().ref_form_schur_system_f32(n, m, N, *[int(t.data_ptr()) for t in (G, Cm, g, c, S, Pinv, gamma)], float(rho))
    if rc:
        raise RuntimeError(f"reference form_schur_system failed rc={rc}")


def compute_dz(n, m, N, Ginv, Cm, g, lam, dz):
    rc = _libschur().ref_compute_dz_f32(n, m, N, *[int(t.data_ptr()) for t in (Ginv, Cm, g, lam, dz)])
    if rc:
        raise RuntimeError(f"reference compute_dz failed rc={rc}")
