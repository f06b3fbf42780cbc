"""mpcgpu_b200 -- B200-native (sm_100a) GBD-PCG: the block-tridiagonal preconditioned conjugate
gradient solve of MPCGPU's SQP loop, behind the reference's own call surface.

The product is ``lib/libgbdpcg.so`` (CUDA kernels + C ABI, ``include/gbd_pcg.h``) and the C++
drop-in headers under ``include/``; this Python package is the thin host mirror used by the tests
and the benchmark.  There is no CPU / PyTorch fallback: a missing library or GPU raises.
"""
from .solver import (HostPlan, PcgConfig, compute_dz, form_schur_system, linsys_window, pcg_launch,  # noqa: F401
                     StepPlan, solve_batched, solve_direct, solvePCG, solvePCG_device)
from . import synth  # noqa: F401

__all__ = ["HostPlan", "PcgConfig", "StepPlan", "solve_direct", "compute_dz", "form_schur_system", "linsys_window", "pcg_launch", "solve_batched", "solvePCG",
           "solvePCG_device", "synth"]
