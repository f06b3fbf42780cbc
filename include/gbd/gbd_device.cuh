// gbd_device.cuh -- sm_100a device primitives shared by the GBD-PCG kernels.
//
//  * thread-block-cluster addressing / DSMEM stores / cluster barrier (PTX, sm_90+)
//  * mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) used to stage the band tiles
//  * IEEE fused multiply-add wrappers and the GLASS summation trees
//
// The summation trees restate the ORDER of the reference's block reductions
// (GLASS/src/L1/reduce.cuh:5-33 and :36-69) so that results are bit-identical to the reference
// kernel; they are written from that behavioural description, not from its code.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbd {

// ----------------------------------------------------------------------------- addressing
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Every poll loop is bounded: a packet that never arrives (a peer that died, a mis-sized launch) ends the kernel with a trap -- the
// launch fails and the C ABI returns GBD_PCG_ERR_CUDA -- instead of hanging the GPU.  2^26 polls is seconds; an exchange takes
// a few hundred cycles.
struct SpinGuard {
    uint32_t spins = 0;
    __device__ __forceinline__ void tick()
    {
        if (++spins > (1u << 26)) __trap();
    }
};

// ---- shared memory through explicit 32-bit addresses.  Measured (profiles/r02_timeline_fastb.log): with many live registers the
// compiler does not keep shared-memory base addresses in registers, it re-derives them -- S2UR SR_CgaCtaId for every block of
// accesses through a C++ pointer, S2R SR_SWINHI for every store into a peer's shared memory -- and these special-register reads
// cost 50-100 cycles each on the dependent chain of an iteration.  The hot loops therefore hold ONE opaque base address and go
// through these accessors; peers' shared memory is addressed through generic pointers formed once (cluster_generic).
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ uint64_t opaque(uint64_t x) { asm volatile("" : "+l"(x)); return x; }
__device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f32x2(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
// generic address of a shared::cluster address (own or a peer's shared memory), for stores that would otherwise read SR_SWINHI
__device__ __forceinline__ uint64_t cluster_generic(uint32_t cluster_addr)
{
    uint64_t g;
    asm volatile("cvta.shared::cluster.u64 %0, %1;" : "=l"(g) : "l"((uint64_t)cluster_addr));
    return g;
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_idx()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_count()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double v)
{
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// split cluster barrier: arrive has release, wait has acquire semantics at cluster scope
__device__ __forceinline__ void cluster_arrive()
{
    __syncwarp();   // .aligned needs the whole warp converged
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync()
{
    cluster_arrive();
    cluster_wait();
}

// ----------------------------------------------------------------------------- mbarrier + TMA bulk
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared; bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    SpinGuard guard;                      // try_wait suspends the thread for a hardware-defined time per attempt
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        guard.tick();
    }
}

// ----------------------------------------------------------------------------- arithmetic
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }      // never contracted
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_(double a) { return fabs(a); }

// Halving tree with odd fix-up and a serial tail of <=3, over a compile-time count held in
// registers (fully unrolled).  Order: while s>3 { odd=s&1; s=(s-odd)/2; v[i]+=v[i+s] (i<s);
// if odd v[0]+=v[2s] }  then v[0]+=v[1] (+v[2]).
template <typename T, uint32_t CNT>
__device__ __forceinline__ T glass_tree(T (&v)[CNT])
{
    uint32_t s = CNT;
#pragma unroll
    for (int lvl = 0; lvl < 32; ++lvl) {
        if (s > 3) {
            const uint32_t odd = s & 1u;
            s = (s - odd) / 2;
#pragma unroll
            for (uint32_t i = 0; i < CNT / 2; ++i)
                if (i < s) v[i] = add_rn(v[i], v[i + s]);
            if (odd) v[0] = add_rn(v[0], v[2 * s]);
        }
    }
#pragma unroll
    for (uint32_t i = 1; i < 3; ++i)
        if (i < s) v[0] = add_rn(v[0], v[i]);
    return v[0];
}

template <uint32_t V>
struct is_pow2 { static constexpr bool value = V && !(V & (V - 1)); };

// Tree over CNT per-knot partials held in shared memory; every lane of the calling warp returns
// the total.  For power-of-two CNT >= 32 the tree is a pure stride-halving one, done with
// in-register adds for strides >= 32 and shuffles below; other counts fall back to the
// per-thread unrolled tree (each thread redundantly, same order).
template <typename T, uint32_t CNT>
__device__ __forceinline__ T glass_tree_smem(const T *part)
{
    if constexpr (is_pow2<CNT>::value && CNT >= 32) {
        constexpr uint32_t PER = CNT / 32;
        const uint32_t lane = threadIdx.x & 31u;
        T v[PER];
#pragma unroll
        for (uint32_t j = 0; j < PER; ++j) v[j] = part[lane + 32 * j];
#pragma unroll
        for (uint32_t h = PER / 2; h >= 1; h /= 2) {
#pragma unroll
            for (uint32_t j = 0; j < PER / 2; ++j)
                if (j < h) v[j] = add_rn(v[j], v[j + h]);
        }
        T x = v[0];
#pragma unroll
        for (uint32_t s = 16; s >= 1; s /= 2) x = add_rn(x, __shfl_down_sync(0xffffffffu, x, s));
        return __shfl_sync(0xffffffffu, x, 0);
    } else {
        T v[CNT];
#pragma unroll
        for (uint32_t j = 0; j < CNT; ++j) v[j] = part[j];
        return glass_tree<T, CNT>(v);
    }
}

}  // namespace gbd
