// dsmem_burst.cu -- how long does a BURST of K remote shared-memory stores from one SM to one peer SM take to become visible?
// CTA 0 sends K packets (one store instruction, K lanes, 8 or 16 bytes each, contiguous or one per 128-byte line) to CTA 1, which
// polls until it has seen all of them and answers with a single packet.  Reported: round trip minus the single-packet round trip / 2.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_burst dsmem_burst.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

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
__device__ __forceinline__ void st8(uint32_t a, uint32_t v, uint32_t ep)
{
    const uint64_t p = ((uint64_t)ep << 32) | v;
    asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(p) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, uint32_t v0, uint32_t v1, uint32_t ep)
{
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v0), "r"(ep), "r"(v1), "r"(ep) : "memory");
}
__device__ __forceinline__ uint64_t ld8(uint32_t a)
{
    uint64_t p;
    asm volatile("ld.volatile.shared::cta.u64 %0, [%1];" : "=l"(p) : "r"(a) : "memory");
    return p;
}

// K lanes send; STRIDE = bytes between packets (8/16 = contiguous, 128 = one per line); SZ = 8 or 16; NINSTR = the burst is split over
// NINSTR store instructions (K / NINSTR lanes each)
template <int K, int SZ, int STRIDE, int NINSTR>
__global__ void __launch_bounds__(32) burst(uint32_t rounds, long long *cyc)
{
    __shared__ __align__(128) unsigned char buf[32 * 128 + 128];
    const uint32_t cr = ctarank(), lane = threadIdx.x;
    for (uint32_t i = lane; i < sizeof(buf) / 4; i += 32) reinterpret_cast<uint32_t *>(buf)[i] = 0;
    __syncthreads();
    cluster_sync();
    const uint32_t mine = smem_u32(buf), other = map_to_cta(mine, cr ^ 1);
    long long t0 = clock64();
    for (uint32_t r = 1; r <= rounds; ++r) {
        if (cr == 0) {
#pragma unroll
            for (int q = 0; q < NINSTR; ++q) {
                const int lo = q * (K / NINSTR), hi = lo + K / NINSTR;
                if ((int)lane >= lo && (int)lane < hi) {
                    if (SZ == 8) st8(other + lane * STRIDE, r, r); else st16(other + lane * STRIDE, r, r, r);
                }
            }
            if (lane == 0) while ((uint32_t)(ld8(mine + 32 * 128) >> 32) != r) {}
            __syncwarp();
        } else {
            bool ok;
            do {
                ok = (int)lane >= K || (uint32_t)(ld8(mine + lane * STRIDE + (SZ == 16 ? 8 : 0)) >> 32) == r;
                ok = __all_sync(0xffffffffu, ok);
            } while (!ok);
            if (lane == 0) st8(other + 32 * 128, r, r);
        }
    }
    long long t1 = clock64();
    cluster_sync();
    if (lane == 0) cyc[cr] = t1 - t0;
}

template <int K, int SZ, int STRIDE, int NINSTR>
double run()
{
    long long *cyc;
    cudaMalloc(&cyc, 16);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(32);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const uint32_t rounds = 5000;
    for (int rep = 0; rep < 2; ++rep) { cudaLaunchKernelEx(&cfg, burst<K, SZ, STRIDE, NINSTR>, rounds, cyc); cudaDeviceSynchronize(); }
    long long h[2];
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    cudaFree(cyc);
    const double rt = (double)h[0] / rounds;
    printf("K=%2d packets of %2d B, stride %3d B, %d instr: round trip %.1f cycles\n", K, SZ, STRIDE, NINSTR, rt);
    return rt;
}

int main()
{
    setvbuf(stdout, NULL, _IONBF, 0);
    run<1, 8, 8, 1>();
    run<2, 8, 8, 1>();
    run<4, 8, 8, 1>();
    run<8, 8, 8, 1>();
    run<16, 8, 8, 1>();
    run<32, 8, 8, 1>();
    run<2, 8, 128, 1>();
    run<4, 8, 128, 1>();
    run<8, 8, 128, 1>();
    run<16, 8, 128, 1>();
    run<32, 8, 128, 1>();
    run<2, 16, 16, 1>();
    run<8, 16, 16, 1>();
    run<32, 16, 16, 1>();
    run<8, 16, 128, 1>();
    run<2, 8, 8, 2>();
    run<4, 8, 8, 4>();
    run<8, 8, 128, 8>();
    return 0;
}
