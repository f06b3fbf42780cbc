"""Host-side mirror of the reference's GBD-PCG call surface, over the C ABI (include/gbd_pcg.h).

Names, argument meaning and outputs follow ``GBD-PCG/include/interface.cuh`` and ``types.cuh``:

* :class:`PcgConfig`      <- ``pcg_config<T>``  (types.cuh:18-35, defaults constants.cuh:14-20)
* :func:`solvePCG`        <- ``solvePCG(h_S, h_gamma, h_lambda, stateSize, knotPoints, config)``
  (interface.cuh:24-89; host buffers)
* :func:`solvePCG_device` <- ``solvePCG(state_size, knot_points, d_S, d_Pinv, ..., config)``
  (interface.cuh:92-144; device buffers, returns the iteration count)
* :func:`pcg_launch`      <- the kernel launch in the SQP loop (include/pcg/sqp.cuh:230)
* :func:`linsys_window`   <- the timed "SQP linsys" window (include/pcg/sqp.cuh:224-241)
* :func:`solve_batched`   <- new: many independent systems per launch

Device buffers are ``torch`` CUDA tensors (torch is only the allocator / stream provider here);
host buffers are numpy arrays.  Nothing in this module computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi


@dataclass
class PcgConfig:
    pcg_exit_tol: float = 1e-6
    pcg_max_iter: int = 25
    pcg_grid: int = 128
    pcg_block: int = 64
    empty_pinv: int = 1


def _ptr(t):
    return 0 if t is None else int(t.data_ptr())


def _stream(stream):
    if stream is None:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)
    return int(getattr(stream, "cuda_stream", stream))


def _check_dev(n, N, batch, d_S, d_Pinv, d_gamma, d_lambda):
    import torch
    dt = d_S.dtype
    if dt not in (torch.float32, torch.float64):
        raise TypeError("GBD-PCG computes in float32 or float64")
    for name, t, cnt in (("d_S", d_S, 3 * n * n * N * batch), ("d_Pinv", d_Pinv, 3 * n * n * N * batch),
                         ("d_gamma", d_gamma, n * N * batch), ("d_lambda", d_lambda, n * N * batch)):
        if not t.is_cuda or not t.is_contiguous() or t.dtype != dt or t.numel() != cnt:
            raise ValueError(f"{name}: expected a contiguous CUDA {dt} tensor of {cnt} elements")
    return dt == torch.float64


def pcg_launch(state_size, knot_points, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_v_temp, d_eta_new_temp,
               d_iters, d_max_iter_exit, max_iter, exit_tol, stream=None):
    """Asynchronous solve; the 12 kernel arguments of pcg<T,n,N> (pcg.cuh:56-68) plus n, N, stream."""
    f64 = _check_dev(state_size, knot_points, 1, d_S, d_Pinv, d_gamma, d_lambda)
    fn = _capi.lib().gbd_pcg_solve_f64 if f64 else _capi.lib().gbd_pcg_solve_f32
    _capi.check(fn(state_size, knot_points, _ptr(d_S), _ptr(d_Pinv), _ptr(d_gamma), _ptr(d_lambda), _ptr(d_r),
                   _ptr(d_p), _ptr(d_v_temp), _ptr(d_eta_new_temp), _ptr(d_iters), _ptr(d_max_iter_exit),
                   int(max_iter), float(exit_tol), _stream(stream)), "gbd_pcg_solve")


def solvePCG_device(state_size, knot_points, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_v_temp, d_eta_new_temp,
                    config: PcgConfig, return_flag: bool = False):
    """interface.cuh:92-144: allocates d_pcg_iters / d_pcg_exit, launches, reads the count back."""
    import torch
    d_iters = torch.zeros(1, dtype=torch.int32, device=d_S.device)
    d_flag = torch.zeros(1, dtype=torch.uint8, device=d_S.device)
    pcg_launch(state_size, knot_points, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_v_temp, d_eta_new_temp,
               d_iters, d_flag, config.pcg_max_iter, config.pcg_exit_tol)
    iters = int(d_iters.item())
    return (iters, bool(d_flag.item())) if return_flag else iters


def linsys_window(state_size, knot_points, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_max_iter_exit,
                  max_iter, exit_tol):
    """sqp.cuh:224-241: sync, launch, two blocking D2H reads, sync.  Returns (iters, flag, microseconds)."""
    if _check_dev(state_size, knot_points, 1, d_S, d_Pinv, d_gamma, d_lambda):
        raise TypeError("linsys_window is float32 only")
    it, fl, us = C.c_uint32(0), C.c_uint8(0), C.c_double(0)
    _capi.check(_capi.lib().gbd_pcg_linsys_f32(state_size, knot_points, _ptr(d_S), _ptr(d_Pinv), _ptr(d_gamma),
                                               _ptr(d_lambda), _ptr(d_r), _ptr(d_p), _ptr(d_iters),
                                               _ptr(d_max_iter_exit), int(max_iter), float(exit_tol),
                                               C.byref(it), C.byref(fl), C.byref(us)), "gbd_pcg_linsys_f32")
    return int(it.value), bool(fl.value), float(us.value)


def solve_batched(state_size, knot_points, batch, d_S, d_Pinv, d_gamma, d_lambda, d_iters, d_max_iter_exit,
                  max_iter, exit_tol, d_r=None, d_p=None, stream=None):
    if _check_dev(state_size, knot_points, batch, d_S, d_Pinv, d_gamma, d_lambda):
        raise TypeError("solve_batched is float32 only")
    _capi.check(_capi.lib().gbd_pcg_solve_batched_f32(state_size, knot_points, batch, _ptr(d_S), _ptr(d_Pinv),
                                                      _ptr(d_gamma), _ptr(d_lambda), _ptr(d_r), _ptr(d_p),
                                                      _ptr(d_iters), _ptr(d_max_iter_exit), int(max_iter),
                                                      float(exit_tol), _stream(stream)), "gbd_pcg_solve_batched_f32")


class HostPlan:
    """Reusable device workspace for host-buffer solves (gbd_pcg_plan_*)."""

    def __init__(self, state_size: int, knot_points: int, batch: int = 1, dtype=np.float32):
        self.n, self.N, self.batch = state_size, knot_points, batch
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("float32 or float64")
        self._h = C.c_void_p()
        self._fn = None
        _capi.check(_capi.lib().gbd_pcg_plan_create(state_size, knot_points, batch, int(self.dtype == np.float64),
                                                    C.byref(self._h)), "gbd_pcg_plan_create")

    def close(self):
        if self._h:
            _capi.lib().gbd_pcg_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, h_S, h_Pinv, h_gamma, h_lambda, max_iter, exit_tol):
        """h_lambda is updated in place.  Returns (iters[batch], max_iter_exit[batch])."""
        n, N, B = self.n, self.N, self.batch
        for name, a, cnt in (("h_S", h_S, 3 * n * n * N * B), ("h_Pinv", h_Pinv, 3 * n * n * N * B),
                             ("h_gamma", h_gamma, n * N * B), ("h_lambda", h_lambda, n * N * B)):
            if not isinstance(a, np.ndarray) or a.dtype != self.dtype or not a.flags.c_contiguous or a.size != cnt:
                raise ValueError(f"{name}: expected a C-contiguous {self.dtype} ndarray of {cnt} elements")
        iters = np.zeros(B, np.uint32)
        flags = np.zeros(B, np.uint8)
        fn = (_capi.lib().gbd_pcg_plan_solve_host_f64 if self.dtype == np.float64
              else _capi.lib().gbd_pcg_plan_solve_host_f32)
        _capi.check(fn(self._h, h_S.ctypes.data, h_Pinv.ctypes.data, h_gamma.ctypes.data, h_lambda.ctypes.data,
                       int(max_iter), float(exit_tol), iters.ctypes.data, flags.ctypes.data), "gbd_pcg_plan_solve_host")
        return iters, flags.astype(bool)


    def solve_raw(self, p_S: int, p_Pinv: int, p_gamma: int, p_lambda: int, max_iter: int, exit_tol: float):
        """Same call with raw host addresses (the caller guarantees dtype / size / contiguity): what a C++ host
        passes.  Returns the iteration count of system 0; per-system results are in self.iters / self.flags."""
        if self._fn is None:
            self._fn = (_capi.lib().gbd_pcg_plan_solve_host_f64 if self.dtype == np.float64
                        else _capi.lib().gbd_pcg_plan_solve_host_f32)
            self.iters = np.zeros(self.batch, np.uint32)
            self.flags = np.zeros(self.batch, np.uint8)
            self._pi, self._pf = self.iters.ctypes.data, self.flags.ctypes.data
        rc = self._fn(self._h, p_S, p_Pinv, p_gamma, p_lambda, max_iter, exit_tol, self._pi, self._pf)
        if rc:
            _capi.check(rc, "gbd_pcg_plan_solve_host")
        return int(self.iters[0])


_plans: dict = {}


def solvePCG(h_S, h_Pinv, h_gamma, h_lambda, stateSize, knotPoints, config: PcgConfig, return_flag: bool = False):
    """interface.cuh:24-89 with the preconditioner made explicit; h_lambda is overwritten with the solution."""
    key = (stateSize, knotPoints, 1, np.dtype(h_S.dtype).str)
    plan = _plans.get(key)
    if plan is None:
        plan = _plans[key] = HostPlan(stateSize, knotPoints, 1, h_S.dtype)
    iters, flags = plan.solve(h_S, h_Pinv, h_gamma, h_lambda, config.pcg_max_iter, config.pcg_exit_tol)
    return (int(iters[0]), bool(flags[0])) if return_flag else int(iters[0])


def form_schur_system(state_size, control_size, knot_points, d_G_dense, d_C_dense, d_g, d_c, d_S, d_Pinv, d_gamma, rho,
                      stream=None):
    """Mirror of form_schur_system<T> (include/pcg/linsys_setup.cuh:621-657, called at include/pcg/sqp.cuh:207): KKT blocks
    -> S, Pinv, gamma in the pcg<> layout; d_G_dense is overwritten with the block inverses.  Asynchronous on `stream`."""
    import torch
    n, m, N = state_size, control_size, knot_points
    want = dict(d_G_dense=(n * n + m * m) * (N - 1) + n * n, d_C_dense=(n * n + n * m) * (N - 1), d_g=(n + m) * (N - 1) + n,
                d_c=n * N, d_S=3 * n * n * N, d_Pinv=3 * n * n * N, d_gamma=n * N)
    for name, t in (("d_G_dense", d_G_dense), ("d_C_dense", d_C_dense), ("d_g", d_g), ("d_c", d_c), ("d_S", d_S),
                    ("d_Pinv", d_Pinv), ("d_gamma", d_gamma)):
        if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or t.numel() != want[name]:
            raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor of {want[name]} elements")
    _capi.check(_capi.lib().gbd_form_schur_system_f32(n, m, N, _ptr(d_G_dense), _ptr(d_C_dense), _ptr(d_g), _ptr(d_c), _ptr(d_S),
                                                      _ptr(d_Pinv), _ptr(d_gamma), float(rho), _stream(stream)),
                "gbd_form_schur_system_f32")


def compute_dz(state_size, control_size, knot_points, d_G_dense, d_C_dense, d_g_val, d_lambda, d_dz, stream=None):
    """Mirror of compute_dz<T> (include/common/dz.cuh:125-136, called at include/pcg/sqp.cuh:250); d_G_dense holds the block
    inverses form_schur_system left there.  Asynchronous on `stream`."""
    import torch
    n, m, N = state_size, control_size, knot_points
    want = dict(d_G_dense=(n * n + m * m) * (N - 1) + n * n, d_C_dense=(n * n + n * m) * (N - 1), d_g_val=(n + m) * (N - 1) + n,
                d_lambda=n * N, d_dz=(n + m) * (N - 1) + n)
    for name, t in (("d_G_dense", d_G_dense), ("d_C_dense", d_C_dense), ("d_g_val", d_g_val), ("d_lambda", d_lambda),
                    ("d_dz", d_dz)):
        if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or t.numel() != want[name]:
            raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor of {want[name]} elements")
    _capi.check(_capi.lib().gbd_compute_dz_f32(n, m, N, _ptr(d_G_dense), _ptr(d_C_dense), _ptr(d_g_val), _ptr(d_lambda),
                                               _ptr(d_dz), _stream(stream)), "gbd_compute_dz_f32")


class StepPlan:
    """One SQP linear-system step for `batch` trajectories (gbd_step_*): form_schur_system -> pcg (warm-started) -> compute_dz
    enqueued on one stream without a host round trip; mirrors include/pcg/sqp.cuh:207-258 per SQP iteration.  The plan owns
    S, Pinv, gamma and the result slots."""

    def __init__(self, state_size: int, control_size: int, knot_points: int, batch: int = 1):
        import ctypes as C
        self.n, self.m, self.N, self.batch = state_size, control_size, knot_points, batch
        h = C.c_void_p()
        _capi.check(_capi.lib().gbd_step_plan_create(state_size, control_size, knot_points, batch, C.byref(h)), "gbd_step_plan_create")
        self._h = h

    def sizes(self):
        n, m, N, B = self.n, self.m, self.N, self.batch
        return dict(G=B * ((n * n + m * m) * (N - 1) + n * n), C=B * (n * n + n * m) * (N - 1), g=B * ((n + m) * (N - 1) + n),
                    c=B * n * N, lam=B * n * N, dz=B * ((n + m) * (N - 1) + n))

    def run(self, d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol, stream=None, direct_fallback=False):
        """Asynchronous.  d_G is overwritten with the block inverses; d_lambda is in/out; d_dz is written.
        direct_fallback: trajectories whose PCG solve hits max_iter are re-solved directly (block cyclic reduction)."""
        import torch
        sz = self.sizes()
        for name, t, key in (("d_G", d_G, "G"), ("d_C", d_C, "C"), ("d_g", d_g, "g"), ("d_c", d_c, "c"), ("d_lambda", d_lambda, "lam"),
                             ("d_dz", d_dz, "dz")):
            if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or t.numel() != sz[key]:
                raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor of {sz[key]} elements")
        fn = _capi.lib().gbd_step_run_fallback_f32 if direct_fallback else _capi.lib().gbd_step_run_f32
        _capi.check(fn(self._h, _ptr(d_G), _ptr(d_C), _ptr(d_g), _ptr(d_c), float(rho), _ptr(d_lambda), _ptr(d_dz), int(max_iter),
                       float(exit_tol), _stream(stream)), "gbd_step_run_f32")

    def results(self, stream=None):
        """Blocks on the stream; returns (iters[batch] uint32, max_iter_exit[batch] uint8) of the last run."""
        import numpy as np
        it = np.zeros(self.batch, np.uint32)
        fl = np.zeros(self.batch, np.uint8)
        _capi.check(_capi.lib().gbd_step_results(self._h, it.ctypes.data, fl.ctypes.data, _stream(stream)), "gbd_step_results")
        return it, fl

    def device_flags(self):
        """[batch] uint8 CUDA tensor view of the max_iter_exit flags of the last run (what the multi-GPU driver all-gathers)."""
        import torch
        ptr = _capi.lib().gbd_step_device_flags(self._h)

        class _Arr:
            __cuda_array_interface__ = {"shape": (self.batch,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    def close(self):
        if self._h:
            _capi.lib().gbd_step_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_direct(state_size, knot_points, d_S, d_gamma, d_lambda, batch: int = 1, stream=None):
    """Direct solve of S lambda = gamma by block cyclic reduction (gbd_bcr_solve_batched_f32): the GPU alternative to the
    reference's CPU QDLDL path (include/qdldl/sqp.cuh:22-49).  Same band layout as pcg_launch; no preconditioner, no initial
    guess.  A different algorithm from the reference's solvers: agreement is to fp32 solver tolerance.  Asynchronous."""
    import torch
    n, N = state_size, knot_points
    for name, t, cnt in (("d_S", d_S, 3 * n * n * N * batch), ("d_gamma", d_gamma, n * N * batch), ("d_lambda", d_lambda, n * N * batch)):
        if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or t.numel() != cnt:
            raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor of {cnt} elements")
    _capi.check(_capi.lib().gbd_bcr_solve_batched_f32(n, N, batch, _ptr(d_S), _ptr(d_gamma), _ptr(d_lambda), _stream(stream)),
                "gbd_bcr_solve_batched_f32")
