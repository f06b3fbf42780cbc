"""ctypes binding of libgbdpcg.so (include/gbd_pcg.h).  No CPU fallback: if the library is missing
this module raises, and every compute entry returns an error without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgbdpcg.so")

OK, ERR_UNSUPPORTED, ERR_BADARG, ERR_CUDA, ERR_NODEVICE, ERR_DEVICE = 0, -1, -2, -3, -4, -5
NUMERICS_BITEXACT, NUMERICS_FAST = 0, 1
FAST_MODES = tuple(range(20, 32))              # gbd_variants.h: mode >= 20 is the tolerance-parity family

# every symbol include/gbd_pcg.h declares (tests check the library exports all of them)
SYMBOLS = [
    "gbd_pcg_abi_version", "gbd_pcg_strerror", "gbd_pcg_last_cuda_error", "gbd_pcg_supported",
    "gbd_pcg_num_variants", "gbd_pcg_variant_at", "gbd_pcg_set_tuning", "gbd_pcg_solve_f32",
    "gbd_pcg_solve_f64", "gbd_pcg_linsys_f32", "gbd_pcg_solve_batched_f32", "gbd_pcg_plan_create",
    "gbd_pcg_plan_destroy", "gbd_pcg_plan_solve_host_f32", "gbd_pcg_plan_solve_host_f64",
    "gbd_pcg_launch_count", "gbd_pcg_set_debug_buffer", "gbd_schur_supported", "gbd_form_schur_system_f32",
    "gbd_compute_dz_f32", "gbd_step_plan_create", "gbd_step_plan_destroy", "gbd_step_run_f32", "gbd_step_results",
    "gbd_step_device_flags", "gbd_schur_csr_nnz", "gbd_schur_csr_pattern_i32", "gbd_schur_csr_values_f32",
    "gbd_bcr_supported", "gbd_bcr_solve_f32", "gbd_bcr_solve_batched_f32", "gbd_bcr_solve_flagged_f32",
    "gbd_step_run_fallback_f32", "gbd_pcg_set_numerics", "gbd_pcg_get_numerics", "gbd_pcg_resolved_variant",
    "gbd_pcg_plan_invalidate", "gbd_schur_set_team",
]

_lib = None


class GbdPcgError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib().gbd_pcg_strerror(status).decode()
        cuda = lib().gbd_pcg_last_cuda_error() if status in (ERR_CUDA, ERR_NODEVICE) else 0
        super().__init__(f"{where}: {msg} (status {status}" + (f", cudaError {cuda})" if cuda else ")"))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mpcgpu_b200.build` (or __graft_entry__.build()). "
            "mpcgpu_b200 has no CPU or PyTorch fallback for the PCG solve.")
    L = C.CDLL(LIB_PATH)
    vp, u32, f32, f64 = C.c_void_p, C.c_uint32, C.c_float, C.c_double
    L.gbd_pcg_abi_version.restype = C.c_int
    L.gbd_pcg_strerror.restype = C.c_char_p
    L.gbd_pcg_strerror.argtypes = [C.c_int]
    L.gbd_pcg_last_cuda_error.restype = C.c_int
    L.gbd_pcg_supported.restype = C.c_int
    L.gbd_pcg_supported.argtypes = [u32, u32, C.c_int]
    L.gbd_pcg_num_variants.restype = C.c_int
    L.gbd_pcg_variant_at.restype = C.c_int
    L.gbd_pcg_variant_at.argtypes = [C.c_int, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(u32), C.POINTER(C.c_size_t)]
    L.gbd_pcg_set_tuning.restype = C.c_int
    L.gbd_pcg_set_tuning.argtypes = [u32, u32, C.c_int, u32, C.c_int]
    L.gbd_pcg_solve_f32.restype = C.c_int
    L.gbd_pcg_solve_f32.argtypes = [u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, f32, vp]
    L.gbd_pcg_solve_f64.restype = C.c_int
    L.gbd_pcg_solve_f64.argtypes = [u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32, f64, vp]
    L.gbd_pcg_linsys_f32.restype = C.c_int
    L.gbd_pcg_linsys_f32.argtypes = [u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, u32, f32, C.POINTER(u32),
                                     C.POINTER(C.c_uint8), C.POINTER(f64)]
    L.gbd_pcg_solve_batched_f32.restype = C.c_int
    L.gbd_pcg_solve_batched_f32.argtypes = [u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, u32, f32, vp]
    L.gbd_pcg_plan_create.restype = C.c_int
    L.gbd_pcg_plan_create.argtypes = [u32, u32, u32, C.c_int, C.POINTER(vp)]
    L.gbd_pcg_plan_destroy.restype = C.c_int
    L.gbd_pcg_plan_destroy.argtypes = [vp]
    L.gbd_pcg_plan_solve_host_f32.restype = C.c_int
    L.gbd_pcg_plan_solve_host_f32.argtypes = [vp, vp, vp, vp, vp, u32, f32, vp, vp]
    L.gbd_pcg_plan_solve_host_f64.restype = C.c_int
    L.gbd_pcg_plan_solve_host_f64.argtypes = [vp, vp, vp, vp, vp, u32, f64, vp, vp]
    L.gbd_pcg_launch_count.restype = C.c_uint64
    L.gbd_schur_set_team.restype = C.c_int
    L.gbd_schur_set_team.argtypes = [C.c_int]
    L.gbd_schur_supported.restype = C.c_int
    L.gbd_schur_supported.argtypes = [u32, u32]
    L.gbd_form_schur_system_f32.restype = C.c_int
    L.gbd_form_schur_system_f32.argtypes = [u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    L.gbd_compute_dz_f32.restype = C.c_int
    L.gbd_compute_dz_f32.argtypes = [u32, u32, u32, vp, vp, vp, vp, vp, vp]
    L.gbd_step_plan_create.restype = C.c_int
    L.gbd_step_plan_create.argtypes = [u32, u32, u32, u32, C.POINTER(vp)]
    L.gbd_step_plan_destroy.restype = C.c_int
    L.gbd_step_plan_destroy.argtypes = [vp]
    L.gbd_step_run_f32.restype = C.c_int
    L.gbd_step_run_f32.argtypes = [vp, vp, vp, vp, vp, f32, vp, vp, u32, f32, vp]
    L.gbd_step_results.restype = C.c_int
    L.gbd_step_results.argtypes = [vp, vp, vp, vp]
    L.gbd_step_device_flags.restype = vp
    L.gbd_step_device_flags.argtypes = [vp]
    L.gbd_schur_csr_nnz.restype = u32
    L.gbd_schur_csr_nnz.argtypes = [u32, u32]
    L.gbd_schur_csr_pattern_i32.restype = C.c_int
    L.gbd_schur_csr_pattern_i32.argtypes = [u32, u32, vp, vp, vp]
    L.gbd_schur_csr_values_f32.restype = C.c_int
    L.gbd_schur_csr_values_f32.argtypes = [u32, u32, vp, vp, vp]
    L.gbd_bcr_supported.restype = C.c_int
    L.gbd_bcr_supported.argtypes = [u32, u32]
    L.gbd_bcr_solve_f32.restype = C.c_int
    L.gbd_bcr_solve_f32.argtypes = [u32, u32, vp, vp, vp, vp]
    L.gbd_bcr_solve_batched_f32.restype = C.c_int
    L.gbd_bcr_solve_batched_f32.argtypes = [u32, u32, u32, vp, vp, vp, vp]
    L.gbd_bcr_solve_flagged_f32.restype = C.c_int
    L.gbd_bcr_solve_flagged_f32.argtypes = [u32, u32, u32, vp, vp, vp, vp, vp]
    L.gbd_step_run_fallback_f32.restype = C.c_int
    L.gbd_step_run_fallback_f32.argtypes = [vp, vp, vp, vp, vp, f32, vp, vp, u32, f32, vp]
    L.gbd_pcg_set_numerics.restype = C.c_int
    L.gbd_pcg_set_numerics.argtypes = [C.c_int]
    L.gbd_pcg_get_numerics.restype = C.c_int
    L.gbd_pcg_resolved_variant.restype = C.c_int
    L.gbd_pcg_resolved_variant.argtypes = [u32, u32, C.c_int, C.c_int, C.POINTER(u32), C.POINTER(C.c_int), C.POINTER(u32),
                                           C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
    L.gbd_pcg_plan_invalidate.restype = C.c_int
    L.gbd_pcg_plan_invalidate.argtypes = [vp]
    L.gbd_pcg_set_debug_buffer.restype = None
    L.gbd_pcg_set_debug_buffer.argtypes = [vp]
    _lib = L
    return L


def check(status: int, where: str):
    if status != OK:
        raise GbdPcgError(status, where)


def variants():
    L = lib()
    out = []
    for i in range(L.gbd_pcg_num_variants()):
        n, N, c, t = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        regs, f64, smem = C.c_int(), C.c_int(), C.c_size_t()
        check(L.gbd_pcg_variant_at(i, C.byref(n), C.byref(N), C.byref(c), C.byref(regs), C.byref(f64), C.byref(t),
                                   C.byref(smem)), "gbd_pcg_variant_at")
        out.append(dict(n=n.value, N=N.value, cluster=c.value, mode=regs.value, f64=bool(f64.value),
                        threads=t.value, smem=smem.value, fast=regs.value in FAST_MODES))
    return out


def set_numerics(numerics: int):
    """GBD_PCG_NUMERICS_FAST (tolerance parity, the default) or GBD_PCG_NUMERICS_BITEXACT; returns the previous setting."""
    L = lib()
    prev = L.gbd_pcg_get_numerics()
    check(L.gbd_pcg_set_numerics(int(numerics)), "gbd_pcg_set_numerics")
    return prev


def resolved_variant(n: int, N: int, f64: bool = False, batched: bool = False):
    """The kernel a launch of this shape would run now: dict(cluster, mode, threads, smem, kernel, fast)."""
    c, t, mode, smem = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_size_t()
    name = C.create_string_buffer(96)
    check(lib().gbd_pcg_resolved_variant(n, N, int(f64), int(batched), C.byref(c), C.byref(mode), C.byref(t), C.byref(smem),
                                         name, 96), "gbd_pcg_resolved_variant")
    return dict(cluster=c.value, mode=mode.value, threads=t.value, smem=smem.value, kernel=name.value.decode(),
                fast=mode.value in FAST_MODES)
