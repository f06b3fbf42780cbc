// pcg.cuh -- part of the header-only DROP-IN for the reference's GBD-PCG include directory
// (see gpu_pcg.cuh for the overview).  Replaces GBD-PCG/include/pcg.cuh: pcgSharedMemSize (:13-20), checkPcgOccupancy (:23-49), pcg<T,n,N> (:54-218).
// The split into files and what each one defines mirrors the reference, because the reference's other headers
// include these files individually (include/mpcsim.cuh:19 and include/pcg/linsys_setup.cuh:3 take only
// "gpuassert.cuh"; include/utils/matrix.cuh:4 takes "utils.cuh") and rely on WHEN the STATE_SIZE / KNOT_POINTS
// defaults of constants.cuh become visible relative to include/common/settings.cuh.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "types.cuh"
#include "gpuassert.cuh"
#include "utils.cuh"
#include "gbd/gbd_grid_pcg.cuh"
#include "gbd/gbd_cluster_pcg_v3.cuh"
#include "gbd/gbd_cluster_pcg_v4.cuh"
#include "gbd/gbd_cluster_pcg_v5.cuh"
#include "gbd/gbd_cluster_pcg_fast.cuh"
#include "gbd/gbd_bcr.cuh"

// Numerics of the drop-in pcg<T,n,N>.  Default (0): bit-identical to the reference kernel.  -DGBD_DROPIN_FAST=1: the tolerance-parity
// kernel of gbd_cluster_pcg_fast.cuh (same contract, iteration count within +-2, lambda within 1e-3 relative: include/gbd_pcg.h)
// for fp32 IIWA shapes N = 32 / 64 / 128 (clusters of N/8 CTAs, 8 knot rows each), when the caller's block has at least 160
// threads -- build with -DPCG_NUM_THREADS=160 (GBD_PCG_MAX_BLOCK then defaults to 160); smaller blocks run the bit-exact body.
// N = 128 uses 16-CTA clusters, a non-portable size: call checkPcgOccupancy first (examples/track_iiwa_pcg.cu:24 does), it
// sets cudaFuncAttributeNonPortableClusterSizeAllowed on the kernel.
#ifndef GBD_DROPIN_FAST
#define GBD_DROPIN_FAST 0
#endif

// -DGBD_DROPIN_DIRECT=1: pcg<T,n,N> does not iterate at all -- the block-tridiagonal system is solved DIRECTLY by block cyclic reduction
// in one thread-block cluster (include/gbd/gbd_bcr.cuh, the GPU counterpart of the reference's CPU QDLDL path; no preconditioner, no
// cap), for fp32, n <= 15, power-of-two N = 32 / 64 / 128 and the reference's default block of exactly 128 threads.  *d_iters = 0 and
// *d_max_iter_exit = false on return.  This is an experiment switch for the closed loop (profiles/r02_closed_loop.json): on the
// reference's IIWA systems PCG often stops at its iteration cap, and this shows what exact solves do to tracking.
#ifndef GBD_DROPIN_DIRECT
#define GBD_DROPIN_DIRECT 0
#endif

#ifndef GBD_PCG_MAX_BLOCK
#if GBD_DROPIN_FAST
#define GBD_PCG_MAX_BLOCK 160           // the tolerance-parity body wants its registers: 160 threads, two blocks per SM
#else
#define GBD_PCG_MAX_BLOCK 256           // upper bound on the caller's block size (register budget of pcg<>)
#endif
#endif

namespace gbd_dropin {
// shared-memory bytes of gbd::GridPcg<T,n,N,R> / gbd::ClusterPcg3<n,N,N/16,false> as plain functions of run-time sizes
constexpr size_t a16(size_t x) { return (x + 15) / 16 * 16; }
constexpr size_t lanes_per_row(size_t n) { return n <= 16 ? 16 : (n <= 32 ? 32 : (n + 31) / 32 * 32); }
constexpr size_t grid_smem(size_t n, size_t N, size_t R, size_t e)
{
    const size_t xs = (n + 3) / 4 * 4;
    return 16 + 2 * a16(e * R * 3 * n * n) + 2 * a16(e * (R + 2) * xs) + 2 * a16(e * 2 * xs) + a16(e * R * lanes_per_row(n)) + a16(e * N);
}
constexpr size_t fast_smem(size_t n, size_t N)
{
    const size_t xs = (n + 3) / 4 * 4;
    return 32 + 2 * a16(4 * 18 * xs) + 3 * a16(4 * 2 * xs) + 2 * a16(4 * N);
}
// v5 body (v3's mapping + packet exchange, tiles loaded straight from global): the FAST shapes with a power-of-two N >= 32
constexpr bool fast5_shape(size_t N) { return N >= 32 && (N & (N - 1)) == 0; }
constexpr size_t fast5_smem(size_t n, size_t N)
{
    const size_t xs = (n + 3) / 4 * 4;
    return 16 + 8 * (3 * N + 8 * xs) + 2 * a16(4 * 18 * xs);
}
// v4 (packet) body: fp32, n <= 16, N = 32 or 64 -> clusters of N/8 CTAs x 128 threads (8 knot rows per CTA)
constexpr bool fast4_shape(size_t n, size_t N, size_t e)
{
#ifdef GBD_DROPIN_NO_CLUSTER
    return false;
#else
    return e == 4 && n >= 2 && n <= 16 && (N == 32 || N == 64);
#endif
}
constexpr size_t fast4_smem(size_t n, size_t N)
{
    const size_t xs = (n + 3) / 4 * 4;
    return 16 + 8 * (3 * N + 8 * xs) + 2 * a16(4 * 10 * xs) + 2 * a16(4 * 8 * 3 * n * n);
}
// tolerance-parity body (GBD_DROPIN_FAST): CTAs per system for the shapes it is built for, 0 = not available
constexpr size_t fastcg_cluster(size_t n, size_t N, size_t e)
{
#ifdef GBD_DROPIN_NO_CLUSTER
    return 0;
#else
    return (GBD_DROPIN_FAST && e == 4 && n % 2 == 0 && n <= 16 && (N == 32 || N == 64 || N == 128)) ? N / 8 : 0;
#endif
}
constexpr size_t fastcg_threads(size_t n, size_t N, size_t e) { return fastcg_cluster(n, N, e) ? (N / fastcg_cluster(n, N, e) + 2) * 16 : 0; }
constexpr size_t fastcg_smem(size_t n, size_t N, size_t e)
{
    const size_t C = fastcg_cluster(n, N, e), R = C ? N / C : 0;
    return C ? 32 + 2 * C * 16 + R * 16 * 8 + 2 * (2 * 2 * 16) * 8 + 4 * ((R + 6) + (R + 4) + (R + 2)) * 16 + 4 * ((R + 4) + (R + 2)) * 3 * n * n : 0;
}
// direct body (GBD_DROPIN_DIRECT): CTAs per system, 0 = not available
constexpr size_t direct_cluster(size_t n, size_t N, size_t e)
{
#ifdef GBD_DROPIN_NO_CLUSTER
    return 0;
#else
    return (GBD_DROPIN_DIRECT && e == 4 && n + 1 <= 16 && (N == 32 || N == 64 || N == 128)) ? N / 8 : 0;
#endif
}
constexpr size_t direct_smem(size_t n, size_t N, size_t e)
{
    const size_t C = direct_cluster(n, N, e), R = C ? N / C : 0, p4 = 3;
    const size_t wf = (n * (2 * n + 1) + p4) / 4 * 4, rowf = (3 * n * n + n + p4) / 4 * 4 + wf + (n + p4) / 4 * 4;
    return C ? 4 * (R * rowf + 4 * (2 * wf + 32)) : 0;
}
constexpr bool fast_shape(size_t n, size_t N, size_t e)
{
#ifdef GBD_DROPIN_NO_CLUSTER
    return false;
#else
    return e == 4 && n % 2 == 0 && n <= 16 && N % 16 == 0 && N / 16 <= 8 && !fast4_shape(n, N, e);
#endif
}
constexpr size_t grid_rows(size_t n, size_t N, size_t e)     // knot rows per CTA of the packet kernel when the block allows
{
    return (!fast_shape(n, N, e) && !fast4_shape(n, N, e) && N % 8 == 0 && N / 8 >= 2 && 8 * lanes_per_row(n) <= GBD_PCG_MAX_BLOCK &&
            grid_smem(n, N, 8, e) <= 48 * 1024) ? 8 : 1;
}
// How pcg<T,n,N> is carried out under the reference's launch (cooperative, grid = N CTAs, caller's block size):
//   FAST  fp32, even n <= 16, N a multiple of 16 with N/16 <= 8 (IIWA: N = 16 .. 128), block >= 128 threads:
//         the kernel carries compile-time cluster dimensions C = N/16; cluster 0 of the grid runs the
//         cluster-resident solver (gbd_cluster_pcg_v3.cuh, DSMEM + mbarrier exchange), all other CTAs return.
//         (power-of-two N >= 32, i.e. N = 128 for IIWA: the packet-exchange body gbd_cluster_pcg_v5.cuh instead.)
//   FAST4 fp32, n <= 16, N = 32 / 64, block >= 128 threads: same scheme with clusters of N/8 CTAs running the packet
//         solver (gbd_cluster_pcg_v4.cuh: {value, epoch} packets polled in shared memory, tiles staged by TMA).
//   GRID  everything else: one CTA per RG knot rows exchanges through L2 packets (gbd_grid_pcg.cuh);
//         RG = 8 when the block has >= 8 row groups of threads, else 1.  CTAs beyond N/RG return.
template <typename T, uint32_t n, uint32_t N>
struct Shape {
    static constexpr uint32_t CDIR = (uint32_t)direct_cluster(n, N, sizeof(T));
    static constexpr bool DIRECT = CDIR != 0;
    using Direct = gbd::BcrShape<DIRECT ? n : 2, DIRECT ? N : 4, DIRECT ? CDIR : 1>;
    static_assert(!DIRECT || Direct::SMEM_BYTES == direct_smem(n, N, sizeof(T)), "run-time smem formula out of sync");
    static constexpr uint32_t CCG = DIRECT ? 0u : (uint32_t)fastcg_cluster(n, N, sizeof(T));
    static constexpr bool FASTCG = CCG != 0;
    static constexpr bool FAST = !FASTCG && !DIRECT && fast_shape(n, N, sizeof(T));
    static constexpr bool FAST4 = !FASTCG && !DIRECT && fast4_shape(n, N, sizeof(T));
    static constexpr uint32_t C = DIRECT ? CDIR : (FASTCG ? CCG : (FAST4 ? N / 8 : (FAST ? N / 16 : 1)));
    using FastCg = gbd::ClusterPcgFast<FASTCG ? n : 2, FASTCG ? N : 8, FASTCG ? CCG : 2>;
    static constexpr uint32_t NT_FASTCG = FASTCG ? FastCg::NT : 0;
    static_assert(!FASTCG || (FastCg::SMEM_BYTES == fastcg_smem(n, N, sizeof(T)) && FastCg::NT == fastcg_threads(n, N, sizeof(T))), "run-time smem formula out of sync");
    // FASTCG with a block that is too small: the grid body serves it (it needs no cluster and ignores the cluster dimensions)
    static constexpr uint32_t NT_FAST = 128;
    // co-residency of all N CTAs (cooperative launch): beyond 2 x 148 CTAs the register budget must allow 4 blocks of
    // 128 threads (= 2 of GBD_PCG_MAX_BLOCK) per SM
    // (tolerance-parity body: 16-CTA clusters of 160 threads; two blocks per SM make all N = 128 CTAs co-resident)
    static constexpr uint32_t MIN_BLOCKS = (N > 296 || FASTCG) ? 2 : 1;
    static constexpr uint32_t G = n <= 16 ? 16 : (n <= 32 ? 32 : (n + 31) / 32 * 32);
    static constexpr uint32_t RG = (uint32_t)grid_rows(n, N, sizeof(T));
    using Fast = gbd::ClusterPcg3<FAST ? n : 2, FAST ? N : 16, FAST ? C : 1, false>;
    static constexpr bool FAST5 = FAST && fast5_shape(N);
    using Fast5 = gbd::ClusterPcg5<FAST5 ? n : 2, FAST5 ? N : 32, FAST5 ? C : 2, false>;
    static_assert(!FAST5 || (Fast5::SMEM_BYTES == fast5_smem(n, N) && Fast5::NT == 128), "run-time smem formula out of sync");
    using Fast4 = gbd::ClusterPcg4<FAST4 ? n : 2, FAST4 ? N : 32, FAST4 ? C : 4>;
    static_assert(!FAST4 || (Fast4::SMEM_BYTES == fast4_smem(n, N) && Fast4::NT == 128), "run-time smem formula out of sync");
    static constexpr size_t SMEM_FAST = FAST4 ? Fast4::SMEM_BYTES : (FAST5 ? Fast5::SMEM_BYTES : (FAST ? Fast::SMEM_BYTES : 0));
    static_assert(!FAST || Fast::SMEM_BYTES == fast_smem(n, N), "run-time smem formula out of sync");
    static_assert(gbd::GridPcg<T, n, N, 1>::SMEM_BYTES == grid_smem(n, N, 1, sizeof(T)), "run-time smem formula out of sync");
    static constexpr size_t SMEM_G1 = gbd::GridPcg<T, n, N, 1>::SMEM_BYTES;
    static constexpr size_t SMEM_GR = gbd::GridPcg<T, n, N, RG>::SMEM_BYTES;
    static constexpr size_t SMEM_BASE = SMEM_FAST > (SMEM_G1 > SMEM_GR ? SMEM_G1 : SMEM_GR) ? SMEM_FAST : (SMEM_G1 > SMEM_GR ? SMEM_G1 : SMEM_GR);
    static constexpr size_t SMEM_CG = (FASTCG && FastCg::SMEM_BYTES > SMEM_BASE) ? FastCg::SMEM_BYTES : SMEM_BASE;
    static constexpr size_t SMEM_BYTES = (DIRECT && Direct::SMEM_BYTES > SMEM_CG) ? Direct::SMEM_BYTES : SMEM_CG;
};
// packet workspace + epoch counter of one instantiation (zero-initialised by the loader); R = 1 is the largest layout
template <typename T, uint32_t n, uint32_t N>
__device__ unsigned long long g_ws[gbd::GridPcg<T, n, N, 1>::WS_WORDS];
template <typename T, uint32_t n, uint32_t N>
__device__ uint32_t g_epoch;
}  // namespace gbd_dropin

template <typename T, uint32_t state_size, uint32_t knot_points>
__global__ void __cluster_dims__(gbd_dropin::Shape<T, state_size, knot_points>::C, 1, 1)
__launch_bounds__(GBD_PCG_MAX_BLOCK, gbd_dropin::Shape<T, state_size, knot_points>::MIN_BLOCKS)
pcg(T *d_S, T *d_Pinv, T *d_gamma, T *d_lambda, T *d_r, T *d_p, T *d_v_temp, T *d_eta_new_temp, uint32_t *d_iters,
    bool *d_max_iter_exit, uint32_t max_iter, T exit_tol)
{
    using SH = gbd_dropin::Shape<T, state_size, knot_points>;
    extern __shared__ __align__(16) unsigned char gbd_dropin_smem[];
    (void)d_v_temp; (void)d_eta_new_temp;          // reference scratch for its smem trees; not needed here
    uint8_t *d_flag = reinterpret_cast<uint8_t *>(d_max_iter_exit);
    if constexpr (SH::DIRECT) {
        if (blockDim.x == SH::Direct::NT) {
            if (gbd::cluster_idx() != 0) return;                  // whole clusters leave together
            const gbd::BcrArgs ba{d_S, d_gamma, d_lambda, 1u, nullptr, nullptr};
            gbd::bcr_cluster_body<state_size, knot_points, SH::CDIR>(ba, reinterpret_cast<float *>(gbd_dropin_smem), 0u, 1u);
            if (gbd::cluster_ctarank() == 0 && threadIdx.x == 0) {
                *d_iters = 0u;
                *d_flag = 0;
            }
            return;
        }
    }
    if constexpr (SH::FASTCG) {
        if (blockDim.x >= SH::NT_FASTCG) {
            if (gbd::cluster_idx() != 0) return;                  // whole clusters leave together
            const uint32_t tma = ((((uintptr_t)d_S) | ((uintptr_t)d_Pinv)) & 15u) == 0 ? 1u : 0u;
            gbd::PcgArgs<float> a{d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_flag, 1u, max_iter, exit_tol, tma};
            gbd::pcg_cluster_fast_init<state_size, knot_points, SH::CCG>(gbd_dropin_smem);
            __syncthreads();
            gbd::cluster_sync();
            if (threadIdx.x < SH::NT_FASTCG) gbd::pcg_cluster_fast_run<state_size, knot_points, SH::CCG, false, false>(a, gbd_dropin_smem, 0u, 1u);
            gbd::cluster_sync();
            return;
        }
    }
    if constexpr (SH::FAST4) {
        if (blockDim.x >= SH::NT_FAST) {
            if (gbd::cluster_idx() != 0) return;                  // whole clusters leave together
            const uint32_t tma4 = ((((uintptr_t)d_S) | ((uintptr_t)d_Pinv)) & 15u) == 0 ? 1u : 0u;
            gbd::PcgArgs<float> a{d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_flag, 1u, max_iter, exit_tol, tma4};
            gbd::pcg_cluster_v4_init<state_size, knot_points, SH::C>(gbd_dropin_smem);
            __syncthreads();
            gbd::cluster_sync();
            if (threadIdx.x < SH::NT_FAST) gbd::pcg_cluster_v4_run<state_size, knot_points, SH::C>(a, gbd_dropin_smem, 0u, 1u);
            gbd::cluster_sync();
            return;
        }
    }
    if constexpr (SH::FAST) {
        if (blockDim.x >= SH::NT_FAST) {
            if (gbd::cluster_idx() != 0) return;                  // whole clusters leave together
            gbd::PcgArgs<float> a{d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_flag, 1u, max_iter, exit_tol, 0u};
            if constexpr (SH::FAST5) {
                gbd::pcg_cluster_v5_init<state_size, knot_points, SH::C, false>(gbd_dropin_smem);
                __syncthreads();
                gbd::cluster_sync();
                if (threadIdx.x < SH::NT_FAST)
                    gbd::pcg_cluster_v5_run<state_size, knot_points, SH::C, false, false>(a, gbd_dropin_smem, 0u, 1u);
                gbd::cluster_sync();
                return;
            }
            gbd::pcg_cluster_v3_init<state_size, knot_points, SH::C, false>(gbd_dropin_smem);
            __syncthreads();
            gbd::cluster_sync();
            if (threadIdx.x < SH::NT_FAST) gbd::pcg_cluster_v3_run<state_size, knot_points, SH::C, false>(a, gbd_dropin_smem, 0u, 1u);
            gbd::cluster_sync();
            return;
        }
    }
    const uint32_t base = gbd_dropin::g_epoch<T, state_size, knot_points>;
    const bool tma = ((((uintptr_t)d_S) | ((uintptr_t)d_Pinv)) & 15u) == 0;
    unsigned long long *ws = gbd_dropin::g_ws<T, state_size, knot_points>;
    uint32_t last;
    if (SH::RG > 1 && blockDim.x >= SH::RG * SH::G) {
        if (blockIdx.x >= knot_points / SH::RG) return;
        last = gbd::pcg_grid_body<T, state_size, knot_points, SH::RG>(d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_flag, max_iter,
                                                                     exit_tol, ws, base, tma, gbd_dropin_smem);
    } else {
        last = gbd::pcg_grid_body<T, state_size, knot_points, 1>(d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_flag, max_iter,
                                                                exit_tol, ws, base, tma, gbd_dropin_smem);
    }
    // every working CTA has read `base` before CTA 0 can get here (it needed all of their packets)
    if (blockIdx.x == 0 && threadIdx.x == 0) gbd_dropin::g_epoch<T, state_size, knot_points> = last;
}

// Dynamic shared memory the launch site must pass (pcg.cuh:13-20 in the reference).  Run-time mirror of
// gbd_dropin::Shape<T,n,N>::SMEM_BYTES: the largest of the three bodies' layouts; always < 48 KB for n <= 32.
template <typename T>
size_t pcgSharedMemSize(uint32_t state_size, uint32_t knot_points)
{
    const size_t n = state_size, N = knot_points, e = sizeof(T);
    size_t need = gbd_dropin::grid_smem(n, N, 1, e);
    if (gbd_dropin::fast_shape(n, N, e) && gbd_dropin::fast_smem(n, N) > need) need = gbd_dropin::fast_smem(n, N);
    if (gbd_dropin::fast4_shape(n, N, e) && gbd_dropin::fast4_smem(n, N) > need) need = gbd_dropin::fast4_smem(n, N);
    if (gbd_dropin::fast_shape(n, N, e) && gbd_dropin::fast5_shape(N) && gbd_dropin::fast5_smem(n, N) > need) need = gbd_dropin::fast5_smem(n, N);
    const size_t rg = gbd_dropin::grid_rows(n, N, e);
    if (rg > 1 && gbd_dropin::grid_smem(n, N, rg, e) > need) need = gbd_dropin::grid_smem(n, N, rg, e);
    if (gbd_dropin::fastcg_smem(n, N, e) > need) need = gbd_dropin::fastcg_smem(n, N, e);
    if (gbd_dropin::direct_smem(n, N, e) > need) need = gbd_dropin::direct_smem(n, N, e);
    return need;
}

// Replaces checkPcgOccupancy (pcg.cuh:23-49), same contract: prints the reference's messages and exits with its codes (5: no
// cooperative launch, 6: the N CTAs of the reference's launch cannot be co-resident), returns true otherwise.  Also opts the kernel
// in to what its bodies need (shared memory beyond 48 KB, 16-CTA clusters), so the reference's launch site needs no change.
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
        exit(5);
    }
    if (smem > 48 * 1024) gpuErrchk(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) (void)cudaGetLastError();
    const int threads = (int)(block.x * block.y * block.z);
    gpuErrchk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    int capacity = sms * per_sm;
    // kernels with compile-time cluster dimensions: whole clusters must be placed, which can be fewer CTAs than sms * per_sm
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(knot_points);
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    int clusters = 0;
    cudaFuncAttributes fa;
    if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) == cudaSuccess && cudaFuncGetAttributes(&fa, kernel) == cudaSuccess &&
        fa.clusterDimMustBeSet == 0 && fa.requiredClusterWidth > 0) {
        const int by_cluster = clusters * fa.requiredClusterWidth;
        if (by_cluster < capacity) capacity = by_cluster;
    } else {
        (void)cudaGetLastError();
    }
    if ((int)knot_points > capacity) {
        printf("Too many knot points ([%d]). Device supports [%d] active blocks, over [%d] SMs.\n", knot_points, capacity, sms);
        exit(6);
    }
    return true;
}
