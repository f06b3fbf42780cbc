// dsmem_hop.cu -- microbenchmark of ONE all-to-all hop inside a thread-block cluster as the fast PCG kernel does it:
// one warp of every CTA sends one packet to every CTA of the cluster (lane d -> CTA d), optionally boundary-row packets
// go to the two neighbours, every warp polls until it has seen all C packets, CTA barrier, next phase.  No arithmetic.
// What does the hop cost as a function of cluster size, packet size and the extra neighbour traffic?
//   DOT = 16: {a, epoch, b, epoch} 16-byte packet;  DOT = 8: one 8-byte packet {a, epoch}
//   HALO = number of 8-byte packets each CTA sends to EACH neighbour per phase (0, 28, 10 x 16-byte when HALO16)
//   PING: two CTAs bounce one 8-byte packet (latency of a single store + poll, per direction)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_hop dsmem_hop.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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
__device__ __forceinline__ void st8(uint32_t a, uint32_t v, uint32_t ep)
{
    const uint64_t p = ((uint64_t)ep << 32) | v;
    asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(p) : "memory");
}
__device__ __forceinline__ void st8w(uint32_t a, uint32_t v, uint32_t ep)
{
    const uint64_t p = ((uint64_t)ep << 32) | v;
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(p) : "memory");
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
__device__ __forceinline__ uint4 ld16(uint32_t a)
{
    uint4 q;
    asm volatile("ld.volatile.shared::cta.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(a) : "memory");
    return q;
}

// MODE bit 0: weak 8-byte stores; bit 1: only warp 0 polls, the others wait at the CTA barrier
template <int C, int DOT, int HALO, int NT, int MODE>
__global__ void __launch_bounds__(NT) hop(uint32_t phases, uint32_t *out, long long *cyc)
{
    __shared__ __align__(16) uint4 dot[2][16];
    __shared__ __align__(16) uint64_t halo[2][2][32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5, cr = ctarank();
    for (uint32_t i = t; i < 32; i += NT) reinterpret_cast<uint4 *>(dot)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = t; i < 128; i += NT) reinterpret_cast<uint64_t *>(halo)[i] = 0;
    __syncthreads();
    cluster_sync();
    const uint32_t dot_u = smem_u32(dot), halo_u = smem_u32(halo);
    const uint32_t peer = map_to_cta(dot_u, lane < C ? lane : cr) + 16 * cr;
    const bool hasl = cr > 0, hasr = cr + 1 < C;
    const uint32_t left = map_to_cta(halo_u, hasl ? cr - 1 : cr) + 8 * (32 + lane);     // my packets land in the left CTA's "from right" half
    const uint32_t right = map_to_cta(halo_u, hasr ? cr + 1 : cr) + 8 * lane;
    constexpr int PERQ = C < 8 ? C : 8, LQ = C / PERQ;
    uint32_t acc = t;
    long long t0 = clock64();
    for (uint32_t ph = 1; ph <= phases; ++ph) {
        const uint32_t par = ph & 1;
        // boundary rows: warps 1 and 2 send HALO packets to each neighbour (lanes 0 .. HALO-1)
        if (HALO > 0 && warp == ((MODE & 4) ? 0 : 1) && (int)lane < HALO) {
            if (hasl) { if (MODE & 1) st8w(left + par * 512, acc, ph); else st8(left + par * 512, acc, ph); }
        }
        if (HALO > 0 && warp == ((MODE & 4) ? 0 : 2) && (int)lane < HALO) {
            if (hasr) { if (MODE & 1) st8w(right + par * 512, acc, ph); else st8(right + par * 512, acc, ph); }
        }
        if ((MODE & 512) && warp == 0) {
            const long long w0 = clock64();
            while (clock64() - w0 < 400) {}
        }
        if ((MODE & 128) && warp == 0 && (lane == 16 || lane == 17)) {
            const bool tol = lane == 16;
            if (tol ? hasl : hasr) st16(map_to_cta(dot_u, tol ? cr - 1 : cr + 1) + 16 * (C + (tol ? 0 : 1)) + par * 256, acc, acc + 1, ph);
        }
        if ((MODE & 256) && warp == 2 && lane < 2) {        // same two packets, but from another warp / instruction
            const bool tol = lane == 0;
            if (tol ? hasl : hasr) st16(map_to_cta(dot_u, tol ? cr - 1 : cr + 1) + 16 * (C + (tol ? 0 : 1)) + par * 256, acc, acc + 1, ph);
        }
        if ((MODE & 32) && warp == 0 && lane == cr) {
            asm volatile("st.volatile.shared::cta.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dot_u + 16 * cr + par * 256), "r"(acc), "r"(ph), "r"(acc + 1), "r"(ph) : "memory");
        } else if (warp == 0 && (int)lane < C) {
            if (DOT == 16) st16(peer + par * 256, acc, acc + 1, ph);
            else if (MODE & 1) st8w(peer + par * 256, acc, ph);
            else st8(peer + par * 256, acc, ph);
        }
        uint32_t s = 0;
        if (!(MODE & 2) || warp == ((MODE & 64) ? 1 : 0)) {
            bool ok;
            uint4 q[PERQ];
            uint64_t h = 0;
            const bool pollh = HALO > 0 && !(MODE & 16) && ((warp == ((MODE & 8) ? 0 : 3) && hasl) || (warp == ((MODE & 8) ? 1 : 4) && hasr)) && (int)lane < HALO;
            do {
                ok = true;
#pragma unroll
                for (int m = 0; m < PERQ; ++m) {
                    if (DOT == 16) {
                        q[m] = ld16(dot_u + 16 * (par * 16 + (lane % LQ) * PERQ + m));
                        ok = ok && q[m].y == ph && q[m].w == ph;
                    } else {
                        const uint64_t v = ld8(dot_u + 16 * (par * 16 + (lane % LQ) * PERQ + m));
                        q[m].x = (uint32_t)v;
                        ok = ok && (uint32_t)(v >> 32) == ph;
                    }
                }
                if (MODE & (128 | 256)) {
                    if (hasr) { const uint4 e = ld16(dot_u + 16 * (par * 16 + C)); ok = ok && e.y == ph && e.w == ph; }
                    if (hasl) { const uint4 e = ld16(dot_u + 16 * (par * 16 + C + 1)); ok = ok && e.y == ph && e.w == ph; }
                }
                if (pollh) {
                    h = ld8(halo_u + 8 * (par * 64 + ((warp == 3 || (warp == 0 && (MODE & 8))) ? 0 : 32) + lane));
                    ok = ok && (uint32_t)(h >> 32) == ph;
                }
            } while (!ok);
#pragma unroll
            for (int m = 0; m < PERQ; ++m) s += q[m].x;
            s += (uint32_t)h;
        }
        acc = acc * 3 + s;
        __syncthreads();
    }
    long long t1 = clock64();
    cluster_sync();
    if (t == 0) cyc[cr] = t1 - t0;
    out[blockIdx.x * NT + t] = acc;
}

// two CTAs bounce one packet: cycles per one-way trip (store issue -> seen by the polling thread of the other CTA)
__global__ void __launch_bounds__(32) ping(uint32_t rounds, long long *cyc)
{
    __shared__ __align__(8) uint64_t slot;
    const uint32_t cr = ctarank();
    if (threadIdx.x == 0) slot = 0;
    __syncthreads();
    cluster_sync();
    const uint32_t mine = smem_u32(&slot), other = map_to_cta(mine, cr ^ 1);
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (uint32_t r = 1; r <= rounds; ++r) {
            if (cr == 0) {
                st8(other, r, r);
                while ((uint32_t)(ld8(mine) >> 32) != r) {}
            } else {
                while ((uint32_t)(ld8(mine) >> 32) != r) {}
                st8(other, r, r);
            }
        }
    }
    long long t1 = clock64();
    cluster_sync();
    if (threadIdx.x == 0) cyc[cr] = t1 - t0;
}

template <int C, int DOT, int HALO, int NT, int MODE>
void run(const char *name)
{
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * C * NT);
    cudaMalloc(&cyc, sizeof(long long) * C);
    auto kern = hop<C, DOT, HALO, NT, MODE>;
    if (C > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(NT);
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const uint32_t phases = 4000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, phases, out, cyc);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    }
    std::vector<long long> h(C);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * C, cudaMemcpyDeviceToHost);
    printf("%-22s C=%2d dot=%2dB halo=%2d/nbr threads=%3d mode=%d : %7.1f cycles/phase\n", name, C, DOT, HALO, NT, MODE, (double)h[0] / phases);
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    setvbuf(stdout, NULL, _IONBF, 0);
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = 0;
#define CASE(...) do { if (only < 0 || only == idx) { __VA_ARGS__; } ++idx; } while (0)
    if (only < 0 || only == 99) {
        long long *cyc;
        cudaMalloc(&cyc, 16);
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(32);
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) { cudaLaunchKernelEx(&cfg, ping, 20000u, cyc); cudaDeviceSynchronize(); }
        long long h[2];
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("ping-pong: %.1f cycles per one-way trip (store -> seen by the peer's polling thread)\n", (double)h[0] / 20000 / 2);
    }
    CASE(run<2, 16, 0, 160, 0>("dot only"));
    CASE(run<4, 16, 0, 160, 0>("dot only"));
    CASE(run<8, 16, 0, 160, 0>("dot only"));
    CASE(run<16, 16, 0, 160, 0>("dot only"));
    CASE(run<8, 8, 0, 160, 0>("dot 8B"));
    CASE(run<16, 8, 0, 160, 0>("dot 8B"));
    CASE(run<16, 8, 0, 160, 1>("dot 8B weak"));
    CASE(run<8, 16, 28, 160, 0>("dot + halo 28x8B"));
    CASE(run<16, 16, 28, 160, 0>("dot + halo 28x8B"));
    CASE(run<16, 16, 14, 160, 0>("dot + halo 14x8B"));
    CASE(run<16, 16, 28, 160, 1>("dot + halo weak"));
    CASE(run<16, 16, 0, 160, 2>("dot, 1 polling warp"));
    CASE(run<8, 16, 0, 160, 2>("dot, 1 polling warp"));
    CASE(run<16, 16, 28, 160, 2>("dot+halo,1 poll warp*"));
    CASE(run<16, 16, 1, 160, 0>("dot + halo 1x8B"));
    CASE(run<16, 16, 28, 160, 4>("halo sent by warp 0"));
    CASE(run<16, 16, 28, 160, 8>("halo polled by w0,w1"));
    CASE(run<16, 16, 28, 160, 16>("halo sent, not polled"));
    CASE(run<8, 16, 28, 160, 16>("halo sent, not polled"));
    CASE(run<4, 16, 0, 128, 0>("dot only 4 warps"));
    CASE(run<4, 16, 0, 160, 2>("dot only 1 poll warp"));
    CASE(run<4, 8, 0, 160, 0>("dot 8B"));
    CASE(run<16, 16, 0, 160, 2 + 32>("1 poller, self local"));
    CASE(run<16, 16, 0, 160, 2 + 64>("1 poller = other warp"));
    CASE(run<16, 16, 0, 160, 2 + 32 + 64>("other warp+self local"));
    CASE(run<8, 16, 0, 160, 2 + 32>("1 poller, self local"));
    CASE(run<8, 16, 0, 160, 2 + 64>("1 poller = other warp"));
    CASE(run<4, 16, 0, 160, 2 + 32>("1 poller, self local"));
    CASE(run<4, 16, 0, 160, 2 + 64>("1 poller = other warp"));
    CASE(run<16, 16, 28, 160, 2 + 8>("1 poller dot+halo"));
    CASE(run<16, 16, 28, 160, 2 + 8 + 64>("other-warp poller dot+halo"));
    CASE(run<8, 16, 0, 160, 2 + 128>("halo in dot instr"));
    CASE(run<8, 16, 0, 160, 2 + 256>("halo pkts other warp"));
    CASE(run<8, 16, 0, 160, 2>("no halo (ref)"));
    CASE(run<4, 16, 0, 160, 2 + 128>("halo in dot instr"));
    CASE(run<4, 16, 0, 160, 2 + 256>("halo pkts other warp"));
    CASE(run<8, 16, 0, 160, 128>("halo in dot, all poll"));
    CASE(run<8, 16, 0, 160, 256>("halo other warp, all poll"));
    CASE(run<8, 16, 0, 160, 2 + 512>("delay, no halo"));
    CASE(run<8, 16, 28, 160, 2 + 8 + 512>("delay, halo 28 early"));
    CASE(run<16, 16, 0, 160, 2 + 512>("delay, no halo"));
    CASE(run<16, 16, 28, 160, 2 + 8 + 512>("delay, halo 28 early"));
    CASE(run<16, 16, 28, 160, 2 + 8 + 512 + 1>("delay, halo early weak"));
    CASE(run<8, 16, 0, 288, 0>("dot only 9 warps"));
    CASE(run<8, 16, 28, 288, 0>("dot + halo 9 warps"));
    return 0;
}
