// constants.cuh -- part of the header-only DROP-IN for the reference's GBD-PCG include directory
// (see gpu_pcg.cuh for the overview).  Replaces GBD-PCG/include/constants.cuh: STATE_SIZE / KNOT_POINTS defaults (:5-11), pcg_constants (:14-20).
// The split into files and what each one defines mirrors the reference, because the reference's other headers
// include these files individually (include/mpcsim.cuh:19 and include/pcg/linsys_setup.cuh:3 take only
// "gpuassert.cuh"; include/utils/matrix.cuh:4 takes "utils.cuh") and rely on WHEN the STATE_SIZE / KNOT_POINTS
// defaults of constants.cuh become visible relative to include/common/settings.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef STATE_SIZE
#define STATE_SIZE 3
#endif
#ifndef KNOT_POINTS
#define KNOT_POINTS 3
#endif

namespace pcg_constants {
inline uint32_t DEFAULT_MAX_PCG_ITER = 25;
template <typename T>
inline T DEFAULT_EPSILON = static_cast<T>(1e-6);
inline dim3 DEFAULT_GRID(128);
inline dim3 DEFAULT_BLOCK(64);
}  // namespace pcg_constants
