// gpu_pcg.cuh -- header-only DROP-IN for the reference's GBD-PCG include directory.
//
// Build the reference's examples/track_iiwa_pcg.cu (or anything that includes "gpu_pcg.cuh",
// "pcg.cuh", "interface.cuh", "types.cuh", "constants.cuh", "utils.cuh", "gpuassert.cuh") with
//     -I<this repo>/include/gbd_dropin -I<this repo>/include        instead of  -IGBD-PCG/include
// and nothing else changes: the same names, template parameters, argument lists and launch
// convention are exported (SURVEY.md section 8b lists them with the reference lines they replace).
//
//   pcg<T, n, N>                  GBD-PCG/include/pcg.cuh:54-68    12-argument __global__, cooperative launch,
//                                                                 grid = N CTAs, any block size >= 32 (<= 256)
//   pcgSharedMemSize<T>(n, N)     pcg.cuh:13-20                   dynamic smem the caller must pass
//   checkPcgOccupancy<T>(...)     pcg.cuh:23-49                   co-residency check (returns false instead of exit())
//   pcg_config<T>, csr_t<T>       types.cuh:7-35
//   pcg_constants::*              constants.cuh:14-20, STATE_SIZE / KNOT_POINTS defaults :5-11
//   solvePCG<T> x3                interface.cuh:8-20, 24-89, 92-144
//   gpuErrchk / gpuAssert         gpuassert.cuh:5-14
//   loadbdVec, bdmv, gato_memcpy, load_block_bd, store_block_bd     utils.cuh:9-161 (the last three are used by
//                                                                 the reference's Schur assembly, which stays as is)
//
// The kernel behind pcg<> is gbd::pcg_grid_body (include/gbd/gbd_grid_pcg.cuh): bit-identical results,
// 2 packet exchanges per iteration instead of 4 grid.sync().  It keeps a small packet workspace in a
// __device__ variable per instantiation, so two launches of the SAME instantiation must not overlap
// in time (the reference is not re-entrant on its scratch buffers either).  For the faster
// cluster-resident kernels link libgbdpcg.so and call gbd_pcg_solve_f32 (include/gbd_pcg.h).
#pragma once
#include "interface.cuh"   // -> pcg.cuh -> types.cuh / gpuassert.cuh / utils.cuh, exactly like GBD-PCG/include/gpu_pcg.cuh:3
