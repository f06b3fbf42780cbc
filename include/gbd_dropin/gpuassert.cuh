// gpuassert.cuh -- part of the header-only DROP-IN for the reference's GBD-PCG include directory
// (see gpu_pcg.cuh for the overview).  Replaces GBD-PCG/include/gpuassert.cuh: gpuAssert / gpuErrchk (:5-14), nothing else.
// The split into files and what each one defines mirrors the reference, because the reference's other headers
// include these files individually (include/mpcsim.cuh:19 and include/pcg/linsys_setup.cuh:3 take only
// "gpuassert.cuh"; include/utils/matrix.cuh:4 takes "utils.cuh") and rely on WHEN the STATE_SIZE / KNOT_POINTS
// defaults of constants.cuh become visible relative to include/common/settings.cuh.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

inline void gpuAssert(cudaError_t code, const char *file, int line, bool abort = true)
{
    if (code == cudaSuccess) return;
    fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
    if (abort) exit(code);
}
#define gpuErrchk(ans) { gpuAssert((ans), __FILE__, __LINE__); }
