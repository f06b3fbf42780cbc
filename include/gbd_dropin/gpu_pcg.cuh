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
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include "gbd/gbd_grid_pcg.cuh"

namespace cgrps = cooperative_groups;   // the reference's other headers rely on this alias

#ifndef STATE_SIZE
#define STATE_SIZE 3
#endif
#ifndef KNOT_POINTS
#define KNOT_POINTS 3
#endif
#ifndef GBD_PCG_MAX_BLOCK
#define GBD_PCG_MAX_BLOCK 256           // upper bound on the caller's block size (register budget of pcg<>)
#endif

// ------------------------------------------------------------------------------------ error helper
inline void gpuAssert(cudaError_t code, const char *file, int line, bool abort = true)
{
    if (code == cudaSuccess) return;
    fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
    if (abort) exit(code);
}
#define gpuErrchk(ans) { gpuAssert((ans), __FILE__, __LINE__); }

// ------------------------------------------------------------------------------------ types / defaults
namespace pcg_constants {
inline uint32_t DEFAULT_MAX_PCG_ITER = 25;
template <typename T>
inline T DEFAULT_EPSILON = static_cast<T>(1e-6);
inline dim3 DEFAULT_GRID(128);
inline dim3 DEFAULT_BLOCK(64);
}  // namespace pcg_constants

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

// ------------------------------------------------------------------------------------ device helpers kept for callers
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

// ------------------------------------------------------------------------------------ the kernel
namespace gbd_dropin {
// packet workspace + epoch counter of one instantiation (zero-initialised by the loader)
template <typename T, uint32_t n, uint32_t N>
__device__ unsigned long long g_ws[gbd::GridPcg<T, n, N, 1>::WS_WORDS];
template <typename T, uint32_t n, uint32_t N>
__device__ uint32_t g_epoch;
}  // namespace gbd_dropin

template <typename T, uint32_t state_size, uint32_t knot_points>
__global__ void __launch_bounds__(GBD_PCG_MAX_BLOCK)
pcg(T *d_S, T *d_Pinv, T *d_gamma, T *d_lambda, T *d_r, T *d_p, T *d_v_temp, T *d_eta_new_temp, uint32_t *d_iters,
    bool *d_max_iter_exit, uint32_t max_iter, T exit_tol)
{
    extern __shared__ __align__(16) unsigned char gbd_dropin_smem[];
    (void)d_v_temp; (void)d_eta_new_temp;          // reference scratch for its smem trees; not needed here
    const uint32_t base = gbd_dropin::g_epoch<T, state_size, knot_points>;
    const bool tma = ((((uintptr_t)d_S) | ((uintptr_t)d_Pinv)) & 15u) == 0;
    const uint32_t last = gbd::pcg_grid_body<T, state_size, knot_points, 1>(
        d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, reinterpret_cast<uint8_t *>(d_max_iter_exit), max_iter, exit_tol,
        gbd_dropin::g_ws<T, state_size, knot_points>, base, tma, gbd_dropin_smem);
    // every CTA has read `base` before CTA 0 can get here (it needed all of their packets)
    if (blockIdx.x == 0 && threadIdx.x == 0) gbd_dropin::g_epoch<T, state_size, knot_points> = last;
}

template <typename T>
size_t pcgSharedMemSize(uint32_t state_size, uint32_t knot_points)
{
    // mirrors gbd::GridPcg<T,n,N,1>::SMEM_BYTES for run-time (n, N): 2 tiles + 2 windows + halos + products + partials
    auto a16 = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t n = state_size, N = knot_points, xs = (n + 3) / 4 * 4;
    const size_t g = n <= 16 ? 16 : (n <= 32 ? 32 : (n + 31) / 32 * 32);
    return 16 + 2 * a16(sizeof(T) * 3 * n * n) + 2 * a16(sizeof(T) * 3 * xs) + 2 * a16(sizeof(T) * 2 * xs) +
           a16(sizeof(T) * g) + a16(sizeof(T) * N);
}

template <typename T>
bool checkPcgOccupancy(void *kernel, dim3 block, uint32_t state_size, uint32_t knot_points)
{
    const size_t smem = pcgSharedMemSize<T>(state_size, knot_points);
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    gpuErrchk(cudaGetDevice(&dev));
    gpuErrchk(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    gpuErrchk(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!coop) {
        printf("[Error] Device does not support Cooperative Threads\n");
        return false;
    }
    if (smem > 48 * 1024) gpuErrchk(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gpuErrchk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)(block.x * block.y * block.z), smem));
    if ((int)knot_points > sms * per_sm) {
        printf("Too many knot points ([%d]). Device supports [%d] active blocks, over [%d] SMs.\n", knot_points,
               sms * per_sm, sms);
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------ host wrappers
template <typename T>
uint32_t solvePCG(csr_t<T> *, csr_t<T> *, T *, T *, unsigned, unsigned, struct pcg_config<T> *)
{
    std::cout << "NOT IMPLEMENTED" << std::endl;   // same as the reference (interface.cuh:8-20)
    exit(12);
}

// device-buffer overload (interface.cuh:92-144).  Like the reference it instantiates the kernel from the
// STATE_SIZE / KNOT_POINTS macros; unlike it, it honours config->pcg_block and checks that they match.
template <typename T>
uint32_t solvePCG(const uint32_t state_size, const uint32_t knot_points, T *d_S, T *d_Pinv, T *d_gamma, T *d_lambda,
                  T *d_r, T *d_p, T *d_v_temp, T *d_eta_new_temp, struct pcg_config<T> *config)
{
    if (state_size != STATE_SIZE || knot_points != KNOT_POINTS) {
        fprintf(stderr, "solvePCG: built for STATE_SIZE=%d KNOT_POINTS=%d, called with %u %u\n", STATE_SIZE, KNOT_POINTS,
                state_size, knot_points);
        exit(13);
    }
    uint32_t *d_pcg_iters;
    bool *d_pcg_exit;
    gpuErrchk(cudaMalloc(&d_pcg_iters, sizeof(uint32_t)));
    gpuErrchk(cudaMalloc(&d_pcg_exit, sizeof(bool)));
    void *kernel = (void *)pcg<T, STATE_SIZE, KNOT_POINTS>;
    void *args[] = {&d_S, &d_Pinv, &d_gamma, &d_lambda, &d_r, &d_p, &d_v_temp, &d_eta_new_temp, &d_pcg_iters, &d_pcg_exit,
                    &config->pcg_max_iter, &config->pcg_exit_tol};
    const size_t smem = pcgSharedMemSize<T>(state_size, knot_points);
    if (smem > 48 * 1024) gpuErrchk(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned block = config->pcg_block.x < 32 ? 32 : config->pcg_block.x;
    gpuErrchk(cudaLaunchCooperativeKernel(kernel, knot_points, block, args, smem));
    gpuErrchk(cudaPeekAtLastError());
    uint32_t h_iters = 0;
    gpuErrchk(cudaMemcpy(&h_iters, d_pcg_iters, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaFree(d_pcg_iters));
    gpuErrchk(cudaFree(d_pcg_exit));
    return h_iters;
}

// host-buffer overload (interface.cuh:24-89).  The reference allocates d_Pinv and never fills it; here
// "no preconditioner" (config->empty_pinv, the only mode that overload admits) means Pinv = identity tiles.
template <typename T>
uint32_t solvePCG(T *h_S, T *h_gamma, T *h_lambda, unsigned stateSize, unsigned knotPoints, struct pcg_config<T> *config)
{
    if (!config->empty_pinv) printf("This api can only be called with no preconditioner\n");
    const size_t nn = (size_t)stateSize * stateSize, mat = 3 * nn * knotPoints, vec = (size_t)stateSize * knotPoints;
    T *d_S, *d_Pinv, *d_gamma, *d_lambda, *d_r, *d_p, *d_v, *d_e;
    gpuErrchk(cudaMalloc(&d_S, mat * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_Pinv, mat * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_gamma, vec * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_lambda, vec * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_r, vec * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_p, vec * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_v, knotPoints * sizeof(T)));
    gpuErrchk(cudaMalloc(&d_e, knotPoints * sizeof(T)));
    T *h_P = (T *)calloc(mat, sizeof(T));
    for (size_t b = 0; b < knotPoints; ++b)
        for (size_t d = 0; d < stateSize; ++d) h_P[b * 3 * nn + nn + d * stateSize + d] = static_cast<T>(1);
    gpuErrchk(cudaMemcpy(d_S, h_S, mat * sizeof(T), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_Pinv, h_P, mat * sizeof(T), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_gamma, h_gamma, vec * sizeof(T), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_lambda, h_lambda, vec * sizeof(T), cudaMemcpyHostToDevice));
    free(h_P);
    const uint32_t iters = solvePCG<T>(stateSize, knotPoints, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_v, d_e, config);
    gpuErrchk(cudaMemcpy(h_lambda, d_lambda, vec * sizeof(T), cudaMemcpyDeviceToHost));
    cudaFree(d_S); cudaFree(d_Pinv); cudaFree(d_gamma); cudaFree(d_lambda);
    cudaFree(d_r); cudaFree(d_p); cudaFree(d_v); cudaFree(d_e);
    return iters;                                   // the reference returns the constant 1 here (interface.cuh:88)
}
