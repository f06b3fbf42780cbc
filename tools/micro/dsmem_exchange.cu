// dsmem_exchange.cu -- microbenchmark of the all-to-all partial exchange inside a thread-block cluster
// (the synchronisation point of the cluster PCG kernels), no arithmetic: how many cycles does one
// "every CTA sends its R values to all C CTAs, every warp waits until it sees all N = C*R values" phase take,
// as a function of C, R and the message format?   Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//   MODE 0: one 8-byte {value, epoch} packet per value per destination (lane d of the 16-lane group -> CTA d)
//   MODE 1: two packets per 16-byte store (lane d of the warp -> CTA d)
//   MODE 2: like 0 but the own CTA's copy is a plain local st.shared
//   MODE 3: like 0, plain 4-byte values + one 8-byte {count, epoch} "flag" packet per (source warp, destination) sent
//           after the values with st.release.cluster (values: weak 4-byte stores); consumer acquires the flags
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t a, uint32_t r)
{
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
}
__device__ __forceinline__ uint32_t ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st1(uint32_t a, float v, uint32_t ep)
{
    const uint64_t p = ((uint64_t)ep << 32) | __float_as_uint(v);
    asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(p) : "memory");
}
__device__ __forceinline__ void st2(uint32_t a, float v0, float v1, uint32_t ep)
{
    const uint64_t p0 = ((uint64_t)ep << 32) | __float_as_uint(v0), p1 = ((uint64_t)ep << 32) | __float_as_uint(v1);
    asm volatile("st.relaxed.cluster.shared::cluster.v2.u64 [%0], {%1, %2};" ::"r"(a), "l"(p0), "l"(p1) : "memory");
}
__device__ __forceinline__ void st_local(uint32_t a, float v, uint32_t ep)
{
    const uint64_t p = ((uint64_t)ep << 32) | __float_as_uint(v);
    asm volatile("st.relaxed.cluster.shared::cta.u64 [%0], %1;" ::"r"(a), "l"(p) : "memory");
}
__device__ __forceinline__ uint64_t ld1(uint32_t a)
{
    uint64_t p;
    asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(p) : "r"(a) : "memory");
    return p;
}

template <int C, int R, int MODE, int WORK>
__global__ void __launch_bounds__(R * 16) exch(uint32_t phases, float *out, long long *cyc)
{
    constexpr int N = C * R, NT = R * 16;
    constexpr int PER = (N / 32 > 8) ? N / 32 : (N >= 32 ? 8 : N / 4);
    constexpr int LW = N / PER;
    __shared__ __align__(16) uint64_t pk[2 * N];
    __shared__ float stage[R];
    __shared__ __align__(8) uint64_t total[2];
    const uint32_t t = threadIdx.x, lane = t & 31, j = t & 15, k = t / 16, cr = ctarank();
    for (int i = t; i < 2 * N; i += NT) pk[i] = 0;
    if (t < 2) total[t] = 0;
    __syncthreads();
    cluster_sync();
    const uint32_t base = smem_u32(pk);
    const uint32_t b = cr * R + k;
    const uint32_t peer1 = map_to_cta(base, j < C ? j : cr) + 8 * b;
    const uint32_t peer2 = map_to_cta(base, lane < C ? lane : cr) + 8 * (b & ~1u);
    const uint32_t mine = base + 8 * (lane % LW);
    float acc = (float)t;
    long long t0 = clock64();
    for (uint32_t ph = 1; ph <= phases; ++ph) {
        const uint32_t off = (ph & 1) * N * 8;
        // stand-in for the band-row chain: WORK dependent FMAs
#pragma unroll 1
        for (int w = 0; w < WORK; ++w) acc = fmaf(acc, 1.0000001f, 1e-9f);
        const float v = acc + (float)ph;
        if (MODE == 0) {
            if (j < C) st1(peer1 + off, v, ph);
        } else if (MODE == 1) {
            const float hi = __shfl_sync(0xffffffffu, v, 16), lo = __shfl_sync(0xffffffffu, v, 0);
            if (lane < C) st2(peer2 + off, lo, hi, ph);
        } else if (MODE == 2) {
            if (j < C) {
                if (j == cr) st_local(base + 8 * b + off, v, ph);
                else st1(peer1 + off, v, ph);
            }
        }
        if (MODE == 4 || MODE == 5) {
            // aggregated: every group leader parks its value in local smem; warp 0 forwards all R values of the
            // CTA with ONE store instruction per destination (R lanes, contiguous 8R bytes)
            if (j == 0) stage[k] = v;
            if (t < 32) {
                asm volatile("bar.sync 1, %0;" ::"r"(NT) : "memory");
                if (MODE == 4) {
                    const float mv = stage[lane < R ? lane : 0];
#pragma unroll
                    for (int d = 0; d < C; ++d)
                        if (lane < R) st1(map_to_cta(base, d) + 8 * (cr * R + lane) + off, mv, ph);
                } else {
                    const float m0 = stage[lane < R / 2 ? 2 * lane : 0], m1 = stage[lane < R / 2 ? 2 * lane + 1 : 0];
#pragma unroll
                    for (int d = 0; d < C; ++d)
                        if (lane < R / 2) st2(map_to_cta(base, d) + 8 * (cr * R + 2 * lane) + off, m0, m1, ph);
                }
            } else {
                asm volatile("bar.arrive 1, %0;" ::"r"(NT) : "memory");
            }
        }
        if (MODE == 6 || MODE == 7) {
            if (j < C) st1(peer1 + off, v, ph);
        }
        float s = 0.f;
        if (MODE == 7) {
            // one poller warp per CTA gathers all N packets and publishes {sum, epoch} locally
            constexpr int PW = (N + 31) / 32;
            if (t < 32) {
                uint64_t w[PW];
                bool okw;
                do {
                    okw = true;
#pragma unroll
                    for (int m = 0; m < PW; ++m) {
                        const int idx = lane + 32 * m;
                        w[m] = idx < N ? ld1(base + off + 8 * idx) : ((uint64_t)ph << 32);
                        okw = okw && (uint32_t)(w[m] >> 32) == ph;
                    }
                } while (!okw);
                float ps = 0.f;
#pragma unroll
                for (int m = 0; m < PW; ++m) ps += __uint_as_float((uint32_t)w[m]);
#pragma unroll
                for (int sh = 16; sh >= 1; sh /= 2) ps += __shfl_xor_sync(0xffffffffu, ps, sh);
                if (lane == 0) st_local(smem_u32(&total[ph & 1]), ps, ph);
            }
            uint64_t r;
            do { r = ld1(smem_u32(&total[ph & 1])); } while ((uint32_t)(r >> 32) != ph);
            s = __uint_as_float((uint32_t)r);
        } else {
        uint64_t q[PER];
        bool ok;
        do {
            ok = true;
#pragma unroll
            for (int m = 0; m < PER; ++m) {
                q[m] = ld1(mine + off + 8 * LW * m);
                ok = ok && (uint32_t)(q[m] >> 32) == ph;
            }
            if (MODE == 6 && !ok) {
                float z = acc;
#pragma unroll
                for (int w = 0; w < 12; ++w) z = fmaf(z, 1.0000001f, 1e-9f);
                acc = z;
            }
        } while (!ok);
#pragma unroll
        for (int m = 0; m < PER; ++m) s += __uint_as_float((uint32_t)q[m]);
        }
        acc = acc * 0.5f + s * 1e-9f;
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync();
    if (t == 0) cyc[cr] = t1 - t0;
    out[blockIdx.x * NT + t] = acc;
}

template <int C, int R, int MODE, int WORK>
void run(const char *name)
{
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * C * R * 16);
    cudaMalloc(&cyc, sizeof(long long) * C);
    auto kern = exch<C, R, MODE, WORK>;
    if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(R * 16);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const uint32_t phases = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, phases, out, cyc);
        if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    }
    std::vector<long long> h(C);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * C, cudaMemcpyDeviceToHost);
    printf("%-10s C=%2d R=%2d N=%3d threads=%3d work=%3d : %7.1f cycles/phase\n", name, C, R, C * R, R * 16, WORK, (double)h[0] / phases);
    cudaFree(out); cudaFree(cyc);
}

#define ALLMODES(C, R, W) run<C, R, 0, W>("pkt8"); run<C, R, 6, W>("backoff"); run<C, R, 7, W>("1poller");

int main()
{
    ALLMODES(4, 8, 0)
    ALLMODES(8, 8, 0)
    ALLMODES(16, 8, 0)
    ALLMODES(8, 16, 0)
    ALLMODES(16, 16, 0)
    ALLMODES(4, 32, 0)
    ALLMODES(16, 32, 0)
    ALLMODES(16, 2, 0)
    return 0;
}
