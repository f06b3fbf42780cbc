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

// Sparse-matrix handle of the reference's QDLDL path; only the member names are contract.
template <typename T>
struct csr_t {
    uint32_t *row_ptr, *col_ind;
    T *val;
    uint32_t rows, cols, nnz;
};

// Solver settings.  Member names, their order and the positional constructor are contract (include/mpcsim.cuh:212-233
// assigns pcg_block / pcg_exit_tol / pcg_max_iter by name); the defaults come from constants.cuh.
template <typename T>
struct pcg_config {
    T pcg_exit_tol = pcg_constants::DEFAULT_EPSILON<T>;
    uint32_t pcg_max_iter = pcg_constants::DEFAULT_MAX_PCG_ITER;
    dim3 pcg_grid = pcg_constants::DEFAULT_GRID;
    dim3 pcg_block = pcg_constants::DEFAULT_BLOCK;
    int empty_pinv = 1;

    pcg_config() = default;
    pcg_config(T tol, uint32_t iters = pcg_constants::DEFAULT_MAX_PCG_ITER, dim3 grid = pcg_constants::DEFAULT_GRID,
               dim3 block = pcg_constants::DEFAULT_BLOCK, int no_pinv = 1)
    {
        pcg_exit_tol = tol;
        pcg_max_iter = iters;
        pcg_grid = grid;
        pcg_block = block;
        empty_pinv = no_pinv;
    }
};
