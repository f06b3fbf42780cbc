// chain_bench.cu -- microbenchmark of the 4-rows-per-thread band-row chain (gbd_cluster_pcg_fastb.cuh) in isolation:
// 256 threads, each with 84 register pairs, windows in shared memory; cycles per chain call with W active warps.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../include -o chain_bench chain_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "gbd/gbd_cluster_pcg_fastb.cuh"

using namespace gbd;
constexpr uint32_t n = 14, H = 7, XS = 16, RPT = 4;

// scalar-FMA restatement of chain_pairs_multi: 24 independent chains of 7, same operations
__device__ __forceinline__ void chain_scalar_multi(const float (&ml)[RPT * 3 * H], const float (&mh)[RPT * 3 * H], const float *xw, float (&out)[RPT])
{
    float lo[RPT][3], hi[RPT][3];
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        float x[16];
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const float4 f = reinterpret_cast<const float4 *>(xw + blk * XS)[q];
            x[4 * q] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
        }
#pragma unroll
        for (uint32_t k = 0; k < RPT; ++k) {
            float sl = __fmul_rn(ml[(k * 3 + blk) * H], x[0]), sh = __fmul_rn(mh[(k * 3 + blk) * H], x[1]);
#pragma unroll
            for (uint32_t c = 1; c < H; ++c) {
                sl = __fmaf_rn(ml[(k * 3 + blk) * H + c], x[2 * c], sl);
                sh = __fmaf_rn(mh[(k * 3 + blk) * H + c], x[2 * c + 1], sh);
            }
            lo[k][blk] = sl; hi[k][blk] = sh;
        }
    }
#pragma unroll
    for (uint32_t k = 0; k < RPT; ++k)
        out[k] = __fadd_rn(__fadd_rn(__fadd_rn(lo[k][0], lo[k][1]), lo[k][2]), __fadd_rn(__fadd_rn(hi[k][0], hi[k][1]), hi[k][2]));
}

template <int MODE, int UNROLL = 1>
__global__ void __launch_bounds__(256, 1) bench(uint32_t *out, const float *src, int iters, int active_warps)
{
    __shared__ __align__(16) float win[40 * XS];
    for (uint32_t i = threadIdx.x; i < 40 * XS; i += blockDim.x) win[i] = src[i];
    const uint32_t t = threadIdx.x, g = (t % 128) / 4;
    f32x2 mm[RPT * 3 * H];
    float ml[RPT * 3 * H], mh[RPT * 3 * H];
#pragma unroll
    for (uint32_t c = 0; c < RPT * 3 * H; ++c) {
        ml[c] = src[1000 + (t * 7 + c) % 3000];
        mh[c] = src[1000 + (t * 11 + c) % 3000];
        mm[c] = pack2(ml[c], mh[c]);
    }
    __syncthreads();
    float o[RPT] = {0, 0, 0, 0};
    uint32_t t0 = 0, t1 = 0;
    if ((int)(t >> 5) < active_warps) {
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t0)::"memory");
        for (int it = 0; it < iters; it += UNROLL) {
#pragma unroll
            for (int uu = 0; uu < UNROLL; ++uu) {
                float r[RPT];
                if constexpr (MODE == 0) chain_pairs_multi<n, XS, RPT>(mm, win + g * XS, r);
                else chain_scalar_multi(ml, mh, win + g * XS, r);
#pragma unroll
                for (uint32_t k = 0; k < RPT; ++k) o[k] += r[k];
                win[(g + 1) * XS + (t & 3)] = o[0] * 1e-30f;      // keeps the loads inside the loop
            }
        }
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t1)::"memory");
    }
    if ((t & 31) == 0) out[t >> 5] = t1 - t0;
    if (o[0] + o[1] + o[2] + o[3] == 0.1234f) out[200] = 1;
}

int main()
{
    uint32_t *d;
    float *src;
    cudaMalloc(&d, 2048);
    cudaMalloc(&src, 5000 * 4);
    float h[5000];
    for (int i = 0; i < 5000; ++i) h[i] = 1e-3f * (i % 97);
    cudaMemcpy(src, h, sizeof h, cudaMemcpyHostToDevice);
    const int iters = 100;
    for (int mode = 0; mode < 2; ++mode)
        for (int aw : {1, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) bench<0><<<1, 256>>>(d, src, iters, aw);
                else bench<1><<<1, 256>>>(d, src, iters, aw);
            }
            cudaError_t e = cudaDeviceSynchronize();
            uint32_t hc[8];
            cudaMemcpy(hc, d, sizeof hc, cudaMemcpyDeviceToHost);
            uint32_t mx = 0;
            for (int w = 0; w < aw; ++w) mx = hc[w] > mx ? hc[w] : mx;
            printf("%s 4-row chain, %d active warps: %.1f cycles per call (%s)\n", mode == 0 ? "FFMA2 " : "scalar", aw, mx / (double)iters, cudaGetErrorString(e));
        }
    // instruction-footprint test: the same chain, the loop body unrolled U times (U x ~2 KB of straight-line code)
    auto go = [&](auto kern, int U) {
        for (int aw : {1, 4}) {
            for (int rep = 0; rep < 2; ++rep) kern<<<1, 256>>>(d, src, 128, aw);
            cudaError_t e = cudaDeviceSynchronize();
            uint32_t hc[8];
            cudaMemcpy(hc, d, sizeof hc, cudaMemcpyDeviceToHost);
            uint32_t mx = 0;
            for (int w = 0; w < aw; ++w) mx = hc[w] > mx ? hc[w] : mx;
            printf("FFMA2 4-row chain unrolled x%-2d, %d active warps: %.1f cycles per call (%s)\n", U, aw, mx / 128.0, cudaGetErrorString(e));
        }
    };
    go(bench<0, 2>, 2);
    go(bench<0, 4>, 4);
    go(bench<0, 8>, 8);
    go(bench<0, 16>, 16);
    go(bench<0, 32>, 32);
    go(bench<0, 64>, 64);
    return 0;
}
