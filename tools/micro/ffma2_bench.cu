// ffma2_bench.cu -- microbenchmark: latency and issue interval of FFMA (fma.rn.f32) and FFMA2 (fma.rn.f32x2) on sm_100a.
// One CTA on one SM; W warps (W = 1, 4, 8: one warp, one per scheduler partition, two per partition); per warp a loop of
// CH independent dependent-chains.  Prints cycles per instruction per warp.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c)
{
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int CH, bool PACKED>
__global__ void bench(uint32_t *out, float seed, int iters)
{
    uint32_t t0, t1;
    if constexpr (PACKED) {
        f32x2 acc[CH], m[CH], x[CH];
        for (int i = 0; i < CH; ++i) {
            acc[i] = (unsigned long long)__float_as_uint(seed + i) * 0x100000001ull;
            m[i] = (unsigned long long)__float_as_uint(1.0f + 1e-7f * (i + threadIdx.x)) * 0x100000001ull;
            x[i] = (unsigned long long)__float_as_uint(1e-9f * (i + 1)) * 0x100000001ull;
        }
        __syncthreads();
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t0)::"memory");
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma2(m[i], acc[i], x[i]);
        }
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t1)::"memory");
        f32x2 s = 0;
        for (int i = 0; i < CH; ++i) s ^= acc[i];
        if (s == 0x1234) out[100] = 1;
    } else {
        float acc[CH], m[CH], x[CH];
        for (int i = 0; i < CH; ++i) {
            acc[i] = seed + i;
            m[i] = 1.0f + 1e-7f * (i + threadIdx.x);
            x[i] = 1e-9f * (i + 1);
        }
        __syncthreads();
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t0)::"memory");
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = fma1(m[i], acc[i], x[i]);
        }
        asm volatile("mov.u32 %0, %%clock;" : "=r"(t1)::"memory");
        float s = 0;
        for (int i = 0; i < CH; ++i) s += acc[i];
        if (s == 0.1234f) out[100] = 1;
    }
    if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
}

template <int CH, bool PACKED>
void run(const char *name, uint32_t *d)
{
    const int iters = 200;
    for (int warps : {1, 4, 8, 16}) {
        bench<CH, PACKED><<<1, 32 * warps>>>(d, 1.0f, iters);
        bench<CH, PACKED><<<1, 32 * warps>>>(d, 1.0f, iters);
        cudaDeviceSynchronize();
        uint32_t h[32];
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("%-6s chains=%2d warps=%2d  cycles/instr/warp %.2f   (SM-wide: %.2f cycles per warp-instruction)\n", name, CH, warps,
               mx / (iters * 8.0 * CH), mx / (iters * 8.0 * CH * warps));
    }
}

int main()
{
    uint32_t *d;
    cudaMalloc(&d, 1024);
    run<1, false>("FFMA", d);
    run<4, false>("FFMA", d);
    run<8, false>("FFMA", d);
    run<16, false>("FFMA", d);
    run<1, true>("FFMA2", d);
    run<4, true>("FFMA2", d);
    run<8, true>("FFMA2", d);
    run<16, true>("FFMA2", d);
    return 0;
}
