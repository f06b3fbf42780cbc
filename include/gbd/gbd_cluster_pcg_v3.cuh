// gbd_cluster_pcg_v3.cuh -- cluster-resident GBD-PCG, third generation: two matrix rows per thread.
//
// fp32, even n <= 16 (IIWA: n = 14).  Same contract and floating-point operation order as the
// reference pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218) -> bit-identical results.  Relative to v2:
//
//  * a knot row lives in an 8-lane group; lane j < n/2 owns matrix rows j and j + n/2 of S and of
//    Pinv (4 x 3n values in registers for the whole solve).  The two band-row FMA chains of a thread
//    are independent, so they interleave in the FMA pipe: same 3n-deep dependent chain latency as
//    one row, half the threads, half the window loads, half the warps running the N-way tree.
//  * the GLASS tree over the n products of a knot row (GLASS/src/L1/reduce.cuh:5-33) starts by adding
//    element i + n/2 onto element i -- both live in the same thread, so the first level costs one
//    add and no shuffle; the remaining tree over n/2 values runs with width-8 shuffles.
//  * CTAs are 8 * R threads (R = N / C knot rows): 128 threads at N = 128, C = 8.  That is the block
//    size the reference launches pcg<> with (PCG_NUM_THREADS, include/common/settings.cuh:111-113), so the
//    same body also runs behind the drop-in pcg<T,n,N> template (include/gbd_dropin/pcg.cuh) inside
//    the reference's cooperative launch: the kernel carries compile-time cluster dimensions, the first
//    cluster of the grid solves the system and the other CTAs return at once.
//  * STAGE = true : tiles arrive by 1-D TMA bulk copies into shared memory and are lifted into
//    registers from there (C-ABI kernels; > 48 KB of dynamic shared memory, opted in by the library).
//    STAGE = false: each thread loads its own rows straight from global memory (drop-in kernel: the
//    reference's launch site passes pcgSharedMemSize() bytes and never opts in to > 48 KB).
//
// Synchronisation, message aggregation and the redundant halo rows are those of v2 (st.async +
// mbarrier complete_tx, one shipping warp, two exchange points per iteration).
#pragma once
#include "gbd_cluster_pcg_v2.cuh"

namespace gbd {

template <uint32_t n, uint32_t N, uint32_t C, bool STAGE>
struct ClusterPcg3 {
    static_assert(n % 2 == 0 && n <= 16 && n >= 2, "v3 needs an even block size <= 16");
    static_assert(N % C == 0 && N >= 2 && C <= 16, "unsupported shape");
    using T = float;
    static constexpr uint32_t H = n / 2;                 // active lanes per knot row; lane j owns rows j, j + H
    static constexpr uint32_t G = 8;                     // lanes per knot row
    static constexpr uint32_t R = N / C;
    static constexpr uint32_t NT = R * G;
    static_assert(NT % 32 == 0 && NT <= 1024, "knot rows per CTA must fill whole warps");
    static constexpr uint32_t W = 3 * n;
    static constexpr uint32_t TILE = 3 * n * n;
    static constexpr uint32_t XS = (n + 3) / 4 * 4;
    static constexpr uint32_t XLEN = (R + 2) * XS;
    static constexpr uint32_t VEC = 4;
    static constexpr bool VEC_PART = R % VEC == 0;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr size_t align16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;                 // 3 mbarriers: tiles, phase A, phase B
    static constexpr size_t OFF_XP = 32;
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HU = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HT = OFF_HU + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_HS = OFF_HT + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_PV = OFF_HS + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_PE = OFF_PV + align16(sizeof(T) * N);
    static constexpr size_t OFF_S = OFF_PE + align16(sizeof(T) * N);
    static constexpr size_t OFF_P = OFF_S + (STAGE ? align16(sizeof(T) * R * TILE) : 0);
    static constexpr size_t SMEM_BYTES = OFF_P + (STAGE ? align16(sizeof(T) * R * TILE) : 0);
};

// two interleaved band-row chains over one padded window: columns ascending, one FMA per column per row
template <uint32_t n, uint32_t XS>
__device__ __forceinline__ void chain2_padded(const float (&m0)[3 * n], const float (&m1)[3 * n], const float *__restrict__ xw,
                                              float &out0, float &out1)
{
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        float x[XS];
#pragma unroll
        for (uint32_t q = 0; q < XS / 4; ++q) {
            const float4 f = reinterpret_cast<const float4 *>(xw + blk * XS)[q];
            x[4 * q] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
        }
#pragma unroll
        for (uint32_t c = 0; c < n; ++c) {
            a0 = fma_rn(m0[blk * n + c], x[c], a0);
            a1 = fma_rn(m1[blk * n + c], x[c], a1);
        }
    }
    out0 = a0;
    out1 = a1;
}

// mbarrier set-up of one CTA; the caller follows it with __syncthreads() and cluster_sync()
template <uint32_t n, uint32_t N, uint32_t C, bool STAGE>
__device__ __forceinline__ void pcg_cluster_v3_init(unsigned char *smem_raw)
{
    using K = ClusterPcg3<n, N, C, STAGE>;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        fence_mbar_init();
    }
}

// Solves systems first_sys, first_sys + sys_stride, ... < a.batch with this cluster.  Must be called by
// threads 0 .. NT-1 of every CTA of the cluster (and only by them: intra-CTA barriers are named barriers
// over NT threads, so a launch may carry idle extra threads).
template <uint32_t n, uint32_t N, uint32_t C, bool STAGE>
__device__ __forceinline__ void pcg_cluster_v3_run(const PcgArgs<float> &a, unsigned char *smem_raw, uint32_t first_sys,
                                                   uint32_t sys_stride)
{
    using K = ClusterPcg3<n, N, C, STAGE>;
    using T = float;
    constexpr uint32_t R = K::R, W = K::W, TILE = K::TILE, G = K::G, H = K::H, XS = K::XS, VEC = K::VEC, NT = K::NT;
    constexpr uint32_t HCH = XS / VEC;

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    uint64_t *barT = bars, *barA = bars + 1, *barB = bars + 2;
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);
    T *hu = reinterpret_cast<T *>(smem_raw + K::OFF_HU);
    T *ht = reinterpret_cast<T *>(smem_raw + K::OFF_HT);
    T *hs = reinterpret_cast<T *>(smem_raw + K::OFF_HS);
    T *part_v = reinterpret_cast<T *>(smem_raw + K::OFF_PV);
    T *part_e = reinterpret_cast<T *>(smem_raw + K::OFF_PE);
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u;
    const bool sender = t < 32;
    const uint32_t j = t % G, k = t / G;               // lane in the knot-row group, local knot row (< R)
    const bool act = j < H;                            // lanes H..7 idle along
    const uint32_t j0 = act ? j : 0, j1 = j0 + H;      // the two matrix rows / vector elements of this thread
    const uint32_t cr = cluster_ctarank();
    const uint32_t b = cr * R + k;
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    const uint32_t left = has_left ? cr - 1 : cr, right = has_right ? cr + 1 : cr;
    const bool own_lhalo = act && k == 0, own_rhalo = act && k == R - 1;

    const uint32_t nb = (has_left ? 1u : 0u) + (has_right ? 1u : 0u);
    const uint32_t halo_bytes = nb * XS * (uint32_t)sizeof(T);
    const uint32_t full_bytes = (C - 1) * R * (uint32_t)sizeof(T) + halo_bytes;

    auto cta_sync = [&]() { named_bar_sync(2, NT); };

    auto ship = [&](T *part, uint64_t *bar, T *halo_l_dst, T *halo_r_dst, bool with_partials, uint32_t expect) {
        if (sender) {
            named_bar_sync(1, NT);
            const uint32_t bar_u = smem_u32(bar);
            if (with_partials && C > 1) {
                if constexpr (K::VEC_PART) {
                    constexpr uint32_t CH = R / VEC;
                    for (uint32_t m = lane; m < (C - 1) * CH; m += 32) {
                        const uint32_t d = m / CH, ch = m % CH, dst = d + (d >= cr ? 1u : 0u);
                        const T *src = part + cr * R + ch * VEC;
                        st_async_vec16(map_to_cta(smem_u32(src), dst), src, map_to_cta(bar_u, dst));
                    }
                } else {
                    for (uint32_t m = lane; m < (C - 1) * R; m += 32) {
                        const uint32_t d = m / R, e = m % R, dst = d + (d >= cr ? 1u : 0u);
                        const T *src = part + cr * R + e;
                        st_async(map_to_cta(smem_u32(src), dst), *src, map_to_cta(bar_u, dst));
                    }
                }
            }
            if (has_left && lane < HCH)
                st_async_vec16(map_to_cta(smem_u32(halo_l_dst + lane * VEC), left), hs + lane * VEC, map_to_cta(bar_u, left));
            if (has_right && lane >= HCH && lane < 2 * HCH)
                st_async_vec16(map_to_cta(smem_u32(halo_r_dst + (lane - HCH) * VEC), right), hs + XS + (lane - HCH) * VEC,
                               map_to_cta(bar_u, right));
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(bar, expect);
        } else {
            named_bar_arrive(1, NT);
        }
    };

    // knot-row dot partial in GLASS order: level one (i, i + n/2) in-thread, the rest with width-8 shuffles
    auto knot_dot = [&](T x0, T y0, T x1, T y1) -> T {
        const T s = add_rn(mul_rn(x0, y0), mul_rn(x1, y1));
        return glass_tree_shfl<T, H, G>(s, j);
    };

    uint32_t phT = 0, phA = 0, phB = 0;
    for (uint32_t sys = first_sys; sys < a.batch; sys += sys_stride) {
        const size_t moff = ((size_t)sys * N + (size_t)cr * R) * TILE;
        const size_t vbase = (size_t)sys * N * n;
        const T *gS = a.S + moff, *gP = a.Pinv + moff;
        const bool tma = STAGE && K::TMA_OK && a.use_tma;

        if constexpr (STAGE) {
            if (tma) {
                if (t == 0) {
                    fence_proxy_async();
                    constexpr uint32_t total = (uint32_t)(sizeof(T) * R * TILE);
                    constexpr uint32_t CHB = 16384;
                    mbar_arrive_expect_tx(barT, 2 * total);
                    for (uint32_t o = 0; o < total; o += CHB) {
                        const uint32_t len = total - o < CHB ? total - o : CHB;
                        tma_bulk_g2s(reinterpret_cast<unsigned char *>(sS) + o, reinterpret_cast<const unsigned char *>(gS) + o, len, barT);
                        tma_bulk_g2s(reinterpret_cast<unsigned char *>(sP) + o, reinterpret_cast<const unsigned char *>(gP) + o, len, barT);
                    }
                }
            } else {
                for (uint32_t i = t; i < R * TILE; i += NT) { sS[i] = gS[i]; sP[i] = gP[i]; }
            }
        }
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) {
            const uint32_t row = i / XS, e = i % XS;
            const long kb = (long)(cr * R) + (long)row - 1;
            xp[i] = (e < n && kb >= 0 && kb < (long)N) ? a.lambda[vbase + (size_t)kb * n + e] : T(0);
            if ((row == 0 && !has_left) || (row == R + 1 && !has_right)) xr[i] = T(0);
        }
        if (t < 2 * XS) hs[t] = T(0);
        T lam0 = T(0), lam1 = T(0), gam0 = T(0), gam1 = T(0);
        if (act) {
            const size_t o = vbase + (size_t)b * n;
            lam0 = a.lambda[o + j0]; lam1 = a.lambda[o + j1];
            gam0 = a.gamma[o + j0]; gam1 = a.gamma[o + j1];
        }

        // this thread's two rows of S and of Pinv live in registers for the whole solve; the tiles the
        // reference never reads (left of block row 0, right of block row N-1) are taken as zero
        T ms0[W], ms1[W], mp0[W], mp1[W];
        const bool skip_l = b == 0, skip_r = b == N - 1;
        if constexpr (STAGE) {
            if (tma) mbar_wait(barT, phT);
            phT ^= 1u;
            cta_sync();
            const T *rowS = sS + k * TILE, *rowP = sP + k * TILE;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                const bool z = !act || (skip_l && c < n) || (skip_r && c >= 2 * n);
                ms0[c] = z ? T(0) : rowS[c * n + j0];
                ms1[c] = z ? T(0) : rowS[c * n + j1];
                mp0[c] = z ? T(0) : rowP[c * n + j0];
                mp1[c] = z ? T(0) : rowP[c * n + j1];
            }
        } else {
            const T *rowS = gS + (size_t)k * TILE, *rowP = gP + (size_t)k * TILE;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                const bool z = !act || (skip_l && c < n) || (skip_r && c >= 2 * n);
                ms0[c] = z ? T(0) : __ldg(rowS + c * n + j0);
                ms1[c] = z ? T(0) : __ldg(rowS + c * n + j1);
                mp0[c] = z ? T(0) : __ldg(rowP + c * n + j0);
                mp1[c] = z ? T(0) : __ldg(rowP + c * n + j1);
            }
            cta_sync();
        }
        const T *wp = xp + k * XS, *wr = xr + k * XS;
        T *own_p = xp + (k + 1) * XS, *own_r = xr + (k + 1) * XS;

        // ---- r = gamma - S*lambda ; exchange boundary rows of r            (pcg.cuh:118-126)
        T c0, c1;
        chain2_padded<n, XS>(ms0, ms1, wp, c0, c1);
        T r0 = gam0 - c0, r1 = gam1 - c1;
        if (act) { own_r[j0] = r0; own_r[j1] = r1; }
        if (own_lhalo) { hs[j0] = r0; hs[j1] = r1; }
        if (own_rhalo) { hs[XS + j0] = r0; hs[XS + j1] = r1; }
        ship(part_e, barA, xr + (R + 1) * XS, xr, false, halo_bytes);
        mbar_wait(barA, phA);
        phA ^= 1u;
        cta_sync();
        // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                             (pcg.cuh:130-149)
        T rt0, rt1;
        chain2_padded<n, XS>(mp0, mp1, wr, rt0, rt1);
        {
            const T x = knot_dot(r0, rt0, r1, rt1);
            if (j == 0) part_e[b] = x;
        }
        if (own_lhalo) { hs[j0] = rt0; hs[j1] = rt1; }
        if (own_rhalo) { hs[XS + j0] = rt0; hs[XS + j1] = rt1; }
        ship(part_e, barB, ht + XS, ht, true, full_bytes);
        mbar_wait(barB, phB);
        phB ^= 1u;
        T eta = glass_tree_part<T, N>(part_e);
        T p0 = rt0, p1 = rt1, u0 = T(0), u1 = T(0);
        if (act) { own_p[j0] = p0; own_p[j1] = p1; }
        if (own_lhalo) { xp[j0] = has_left ? ht[j0] : T(0); xp[j1] = has_left ? ht[j1] : T(0); }
        if (own_rhalo) {
            xp[(R + 1) * XS + j0] = has_right ? ht[XS + j0] : T(0);
            xp[(R + 1) * XS + j1] = has_right ? ht[XS + j1] : T(0);
        }

        uint32_t iter = 0;
        uint8_t max_iter_exit = 1;
        for (; iter < a.max_iter; ++iter) {
            cta_sync();
            // ---- upsilon = S*p ; v = p.upsilon                             (pcg.cuh:156-167)
            chain2_padded<n, XS>(ms0, ms1, wp, u0, u1);
            {
                const T x = knot_dot(p0, u0, p1, u1);
                if (j == 0) part_v[b] = x;
            }
            if (own_lhalo) { hs[j0] = u0; hs[j1] = u1; }
            if (own_rhalo) { hs[XS + j0] = u0; hs[XS + j1] = u1; }
            ship(part_v, barA, hu + XS, hu, true, full_bytes);
            mbar_wait(barA, phA);
            phA ^= 1u;
            const T alpha = eta / glass_tree_part<T, N>(part_v);               // :169
            // ---- lambda += alpha p ; r -= alpha upsilon (own rows + halo copies)   (:172-176)
            lam0 = fma_rn(alpha, p0, lam0); lam1 = fma_rn(alpha, p1, lam1);
            r0 = fma_rn(-alpha, u0, r0); r1 = fma_rn(-alpha, u1, r1);
            if (act) { own_r[j0] = r0; own_r[j1] = r1; }
            if (own_lhalo && has_left) { xr[j0] = fma_rn(-alpha, hu[j0], xr[j0]); xr[j1] = fma_rn(-alpha, hu[j1], xr[j1]); }
            if (own_rhalo && has_right) {
                T *h = xr + (R + 1) * XS;
                h[j0] = fma_rn(-alpha, hu[XS + j0], h[j0]);
                h[j1] = fma_rn(-alpha, hu[XS + j1], h[j1]);
            }
            cta_sync();
            // ---- r~ = Pinv*r ; eta' = r.r~                                 (:180-193)
            chain2_padded<n, XS>(mp0, mp1, wr, rt0, rt1);
            {
                const T x = knot_dot(r0, rt0, r1, rt1);
                if (j == 0) part_e[b] = x;
            }
            if (own_lhalo) { hs[j0] = rt0; hs[j1] = rt1; }
            if (own_rhalo) { hs[XS + j0] = rt0; hs[XS + j1] = rt1; }
            ship(part_e, barB, ht + XS, ht, true, full_bytes);
            mbar_wait(barB, phB);
            phB ^= 1u;
            const T eta_new = glass_tree_part<T, N>(part_e);
            if (abs_(eta_new) < a.exit_tol) { ++iter; max_iter_exit = 0; break; }   // :195
            const T beta = eta_new / eta;                                       // :199-200
            eta = eta_new;
            // ---- p = r~ + beta p (own rows + halo copies)                   (:203-206)
            p0 = fma_rn(beta, p0, rt0); p1 = fma_rn(beta, p1, rt1);
            if (act) { own_p[j0] = p0; own_p[j1] = p1; }
            if (own_lhalo && has_left) { xp[j0] = fma_rn(beta, xp[j0], ht[j0]); xp[j1] = fma_rn(beta, xp[j1], ht[j1]); }
            if (own_rhalo && has_right) {
                T *h = xp + (R + 1) * XS;
                h[j0] = fma_rn(beta, h[j0], ht[XS + j0]);
                h[j1] = fma_rn(beta, h[j1], ht[XS + j1]);
            }
        }

        // ---- outputs                                                        (:212-215)
        if (act) {
            const size_t o = vbase + (size_t)b * n;
            a.lambda[o + j0] = lam0; a.lambda[o + j1] = lam1;
            if (a.r_out) { a.r_out[o + j0] = r0; a.r_out[o + j1] = r1; }
            if (a.p_out) { a.p_out[o + j0] = p0; a.p_out[o + j1] = p1; }
        }
        if (cr == 0 && t == 0) {
            store_result(a, sys, iter, max_iter_exit);
        }
        cta_sync();
        // batches: no CTA starts the next system's exchanges while a peer may still be reading this system's last packets /
        // phases (neighbour-only prologue exchanges do not order far CTAs).  Once per solve, off the iteration path; the
        // drop-in pcg<> (batch == 1, extra idle threads in the block) never gets here.
        if (a.batch > 1) cluster_sync();
    }
}

// C-ABI kernel: persistent clusters looping over a batch of systems
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
__global__ void __launch_bounds__(ClusterPcg3<n, N, C, true>::NT, MINB)
pcg_cluster_kernel_v3(const PcgArgs<float> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_v3_init<n, N, C, true>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, all mbarriers initialised, before any DSMEM traffic
    pcg_cluster_v3_run<n, N, C, true>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
