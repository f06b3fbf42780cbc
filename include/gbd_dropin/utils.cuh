// utils.cuh -- part of the header-only DROP-IN for the reference's GBD-PCG include directory
// (see gpu_pcg.cuh for the overview).  Replaces GBD-PCG/include/utils.cuh: loadbdVec (:9-40), bdmv (:46-85), gato_memcpy (:87-94), load_block_bd (:96-130), store_block_bd (:132-161).
// The split into files and what each one defines mirrors the reference, because the reference's other headers
// include these files individually (include/mpcsim.cuh:19 and include/pcg/linsys_setup.cuh:3 take only
// "gpuassert.cuh"; include/utils/matrix.cuh:4 takes "utils.cuh") and rely on WHEN the STATE_SIZE / KNOT_POINTS
// defaults of constants.cuh become visible relative to include/common/settings.cuh.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cooperative_groups.h>
#include "types.cuh"
#if __has_include("glass.cuh")
#include "glass.cuh"                    // the reference's utils.cuh:5 pulls GLASS in; its other headers rely on that
#endif
#include "gbd/gbd_device.cuh"

namespace cgrps = cooperative_groups;   // the reference's other headers rely on this alias (utils.cuh:7, pcg.cuh:11)

// [x_{b-1}; x_b; x_{b+1}] window of a global N*n vector into 3n shared values (edge blocks load 2n)
template <typename T, uint32_t block_dim, uint32_t max_block_id>
__device__ void loadbdVec(T *s_var, const uint32_t block_id, T *d_var_b)
{
    const int lo = block_id == 0 ? (int)block_dim : 0, hi = block_id == max_block_id ? 2 * (int)block_dim : 3 * (int)block_dim;
    for (int i = lo + (int)threadIdx.x; i < hi; i += (int)blockDim.x) s_var[i] = d_var_b[i - (int)block_dim];
}

// one block row of the band matvec: dst[r] = sum_c mat[n*c + r] * vec[c], c ascending, edge rows use 2n columns
template <typename T>
__device__ void bdmv(T *s_dst, T *s_mat, T *s_vec, uint32_t b_dim, uint32_t max_block_id, uint32_t block_id)
{
    const uint32_t c0 = block_id == 0 ? b_dim : 0, c1 = block_id == max_block_id && block_id != 0 ? 2 * b_dim : 3 * b_dim;
    for (uint32_t r = threadIdx.x; r < b_dim; r += blockDim.x) {
        T acc = static_cast<T>(0);
        for (uint32_t c = c0; c < c1; ++c) acc = gbd::fma_rn(s_mat[b_dim * c + r], s_vec[c], acc);
        s_dst[r] = acc;
    }
}

template <typename T>
__device__ void gato_memcpy(T *dst, T *src, unsigned size_Ts)
{
    for (unsigned i = threadIdx.x; i < size_Ts; i += blockDim.x) dst[i] = src[i];
}

// tile (brow, bcol) of a [N][3][n][n] band matrix -> dst (optionally transposed)
template <typename T>
__device__ void load_block_bd(uint32_t b_dim, uint32_t m_dim, T *src, T *dst, unsigned bcol, unsigned brow,
                              bool transpose = false,
                              cooperative_groups::thread_group g = cooperative_groups::this_thread_block())
{
    if (bcol > 2 || brow > m_dim - 1) {
        printf("doing somehting wrong in load_block_bd\n");
        return;
    }
    const T *tile = src + (size_t)brow * 3 * b_dim * b_dim + (size_t)bcol * b_dim * b_dim;
    for (unsigned i = threadIdx.x; i < b_dim * b_dim; i += blockDim.x) {
        if (!transpose) dst[i] = tile[i];
        else dst[(i % b_dim) * b_dim + i / b_dim] = tile[i];
    }
}

// src -> tile (BLOCKNO, col), scaled by an integer multiplier (the reference stores with -1)
template <typename T>
__device__ void store_block_bd(uint32_t b_dim, uint32_t m_dim, T *src, T *dst, unsigned col, unsigned BLOCKNO,
                               int multiplier = 1,
                               cooperative_groups::thread_group g = cooperative_groups::this_thread_block())
{
    T *tile = dst + (size_t)BLOCKNO * 3 * b_dim * b_dim + (size_t)col * b_dim * b_dim;
    if (multiplier == 1) {
        for (unsigned i = threadIdx.x; i < b_dim * b_dim; i += blockDim.x) tile[i] = src[i];
    } else {
        for (unsigned i = g.thread_rank(); i < b_dim * b_dim; i += g.size()) tile[i] = src[i] * multiplier;
    }
}
