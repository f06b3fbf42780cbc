// l2_exchange.cu -- microbenchmark of the grid-wide all-gather through L2 used by the grid PCG kernel:
// G co-resident CTAs (one per SM), each publishes R 8-byte {value, epoch} packets, every CTA polls all N = G*R.
//   MODE 0: st.relaxed.gpu / ld.relaxed.gpu packets, every thread polls N/NT packets (combined loop)
//   MODE 1: same, but only warp 0 polls (N/32 packets per lane) and the rest of the CTA waits at __syncthreads
//   MODE 2: packets written with st.global.cg-like volatile / polled with ld.volatile
//   MODE 4: replicated packets: the producer writes one copy per consumer CTA ([consumer][N] layout), every CTA polls only
//           its private copy with all threads (no line is read by more than one CTA); MODE 5: same, warp 0 polls
//   MODE 3: counter barrier: plain stores, __threadfence, atomicAdd on one counter; pollers spin on the counter, then ld.cg the data
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void st_pkt(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_pkt(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int R, int MODE>
__global__ void __launch_bounds__(128) exch(int G, uint32_t phases, unsigned long long *ws, unsigned int *ctr, float *out, long long *cyc)
{
    extern __shared__ float part[];
    const int N = G * R, t = threadIdx.x, cta = blockIdx.x, lane = t & 31;
    float acc = (float)t;
    long long t0 = clock64();
    for (uint32_t ph = 1; ph <= phases; ++ph) {
        unsigned long long *base = ws + (size_t)(ph & 1) * N;
        const float v = acc + (float)ph;
        if (MODE == 4 || MODE == 5) {
            unsigned long long *rep = ws + (size_t)(ph & 1) * N * G;      // [consumer][N]
            const unsigned long long pk = ((unsigned long long)ph << 32) | __float_as_uint(v);
            // thread t publishes packet (t % R) of this CTA to consumers t / R, t / R + 128 / R, ...
            for (int c = t / R; c < G; c += 128 / R) st_pkt(rep + (size_t)c * N + cta * R + (t % R), pk);
            const unsigned long long *mine = rep + (size_t)cta * N;
            if (MODE == 4 || t < 32) {
                const int stride = MODE == 4 ? 128 : 32, me = MODE == 4 ? t : lane;
                unsigned long long w[8];
                bool ok;
                do {
                    ok = true;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int i = me + stride * q;
                        if (i < N) { w[q] = ld_pkt(mine + i); ok = ok && (uint32_t)(w[q] >> 32) == ph; }
                    }
                } while (!ok);
#pragma unroll
                for (int q = 0; q < 8; ++q) { const int i = me + stride * q; if (i < N) part[i] = __uint_as_float((uint32_t)w[q]); }
            }
        } else if (MODE == 3) {
            if (t < R) reinterpret_cast<volatile float *>(base)[2 * (cta * R + t)] = v;
            __syncthreads();
            if (t == 0) {
                __threadfence();
                atomicAdd(ctr, 1u);
                while (*reinterpret_cast<volatile unsigned int *>(ctr) < (unsigned)G * ph) {}
                __threadfence();
            }
            __syncthreads();
            for (int i = t; i < N; i += 128) part[i] = __ldcg(reinterpret_cast<const float *>(base) + 2 * i);
        } else {
            const unsigned long long pk = ((unsigned long long)ph << 32) | __float_as_uint(v);
            if (t < R) { if (MODE == 2) st_vol(base + cta * R + t, pk); else st_pkt(base + cta * R + t, pk); }
            if (MODE == 1) {
                if (t < 32) {
                    for (int i0 = 0; i0 < N; i0 += 32 * 8) {
                        unsigned long long w[8];
                        bool ok;
                        do {
                            ok = true;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int i = i0 + lane + 32 * q;
                                if (i < N) { w[q] = ld_pkt(base + i); ok = ok && (uint32_t)(w[q] >> 32) == ph; }
                            }
                        } while (!ok);
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const int i = i0 + lane + 32 * q; if (i < N) part[i] = __uint_as_float((uint32_t)w[q]); }
                    }
                }
            } else {
                unsigned long long w[8];
                bool ok;
                do {
                    ok = true;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int i = t + 128 * q;
                        if (i < N) { w[q] = (MODE == 2) ? ld_vol(base + i) : ld_pkt(base + i); ok = ok && (uint32_t)(w[q] >> 32) == ph; }
                    }
                } while (!ok);
#pragma unroll
                for (int q = 0; q < 8; ++q) { const int i = t + 128 * q; if (i < N) part[i] = __uint_as_float((uint32_t)w[q]); }
            }
        }
        __syncthreads();
        float s = 0.f;
        for (int i = lane; i < N; i += 32) s += part[i];
        acc = acc * 0.5f + s * 1e-9f;
        __syncthreads();
    }
    long long t1 = clock64();
    if (t == 0) cyc[cta] = t1 - t0;
    out[cta * 128 + t] = acc;
}

template <int R, int MODE>
void run(int G, const char *name)
{
    const int N = G * R;
    unsigned long long *ws; unsigned int *ctr; float *out; long long *cyc;
    cudaMalloc(&ws, sizeof(unsigned long long) * 2 * N * G); cudaMemset(ws, 0, sizeof(unsigned long long) * 2 * N * G);
    cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
    cudaMalloc(&out, sizeof(float) * G * 128); cudaMalloc(&cyc, sizeof(long long) * G);
    auto kern = exch<R, MODE>;
    const int smem = 120 * 1024;   // one CTA per SM
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const uint32_t phases = 500;
    void *args[] = {(void *)&G, (void *)&phases, (void *)&ws, (void *)&ctr, (void *)&out, (void *)&cyc};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)kern, dim3(G), dim3(128), args, smem, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    std::vector<long long> h(G);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * G, cudaMemcpyDeviceToHost);
    printf("%-14s G=%3d R=%d N=%3d : %8.1f cycles/phase\n", name, G, R, N, (double)h[0] / phases);
    cudaFree(ws); cudaFree(ctr); cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int G : {2, 8, 32, 128}) {
        run<1, 0>(G, "pkt all-poll"); run<1, 1>(G, "pkt warp0-poll"); run<1, 4>(G, "replicated"); run<1, 5>(G, "replicated w0");
    }
    run<2, 0>(128, "pkt all-poll"); run<2, 1>(128, "pkt warp0-poll"); run<2, 4>(128, "replicated"); run<2, 5>(128, "replicated w0");
    run<2, 4>(64, "replicated"); run<4, 4>(64, "replicated"); run<4, 5>(64, "replicated w0");
    return 0;
}
