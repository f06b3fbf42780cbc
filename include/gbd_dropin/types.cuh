// types.cuh -- part of the header-only DROP-IN for the reference's GBD-PCG include directory
// (see gpu_pcg.cuh for the overview).  Replaces GBD-PCG/include/types.cuh: csr_t (:7-15), pcg_config (:18-35).
// The split into files and what each one defines mirrors the reference, because the reference's other headers
// include these files individually (include/mpcsim.cuh:19 and include/pcg/linsys_setup.cuh:3 take only
// "gpuassert.cuh"; include/utils/matrix.cuh:4 takes "utils.cuh") and rely on WHEN the STATE_SIZE / KNOT_POINTS
// defaults of constants.cuh become visible relative to include/common/settings.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "constants.cuh"

template <typename T>
struct csr_t {
    uint32_t *row_ptr;
    uint32_t *col_ind;
    T *val;
    uint32_t rows, cols, nnz;
};

template <typename T>
struct pcg_config {
    T pcg_exit_tol;
    uint32_t pcg_max_iter;
    dim3 pcg_grid;
    dim3 pcg_block;
    int empty_pinv;
    pcg_config(T exit_tol = pcg_constants::DEFAULT_EPSILON<T>, uint32_t max_iter = pcg_constants::DEFAULT_MAX_PCG_ITER,
               dim3 grid = pcg_constants::DEFAULT_GRID, dim3 block = pcg_constants::DEFAULT_BLOCK, int empty_pinv_ = 1)
        : pcg_exit_tol(exit_tol), pcg_max_iter(max_iter), pcg_grid(grid), pcg_block(block), empty_pinv(empty_pinv_)
    {
    }
};
