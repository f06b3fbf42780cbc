// gbd_cluster_pcg_v2.cuh -- cluster-resident GBD-PCG, second generation (n <= 32).
//
// Same contract and the same floating-point operation order as gbd_cluster_pcg.cuh (and therefore
// as the reference pcg<T,n,N>, GBD-PCG/include/pcg.cuh:54-218); what changes is how the two
// per-iteration synchronisation points are implemented:
//
//  * no barrier.cluster in the iteration loop.  Every cross-CTA datum (a knot row's dot partial for
//    all C CTAs, a boundary row of upsilon / r~ for the neighbour) is pushed with
//    `st.async ... mbarrier::complete_tx::bytes` straight into the consumer's shared memory and
//    counted on the consumer's mbarrier; the consumer arms the expected byte count once per phase
//    and waits on its own barrier.  Data and signal travel together, no release/acquire fence, no
//    L1 invalidation (barrier.cluster costs ~380 cycles + CCTL.IVALL on this part).
//  * each knot row sits in one half-warp (16 lanes, n <= 16) or one warp (n <= 32): the per-knot
//    GLASS tree over the n products is done with width-limited shuffles, no smem round trip.
//  * messages are aggregated: a CTA's R partials and its two boundary rows are first written to its
//    own shared memory, one warp (released by a named barrier the other warps only *arrive* at)
//    then ships them as 16-byte st.async vectors -- R/4 messages per peer CTA + 4 per neighbour
//    instead of R + n scalar ones.  Measured on B200: every complete_tx on one mbarrier costs
//    ~5 cycles at the receiver, so message COUNT, not bytes, set the iteration time of the first
//    scalar-message version (0.90 / 1.39 / 4.0 us per iteration at N = 32 / 128 / 512).
//  * vector windows are padded to a multiple of 4 elements per knot row: 128-bit smem loads.
//  * one __syncthreads per half iteration (own p / r rows visible before the band-row chains).
#pragma once
#include "gbd_cluster_pcg.cuh"

namespace gbd {

__device__ __forceinline__ void st_async(uint32_t addr, float v, uint32_t mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(addr), "f"(v),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async(uint32_t addr, double v, uint32_t mbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(addr), "d"(v),
                 "r"(mbar)
                 : "memory");
}

// 16-byte vector messages (4 floats / 2 doubles) read from local shared memory
__device__ __forceinline__ void st_async_vec16(uint32_t dst, const float *src, uint32_t mbar)
{
    const float4 f = *reinterpret_cast<const float4 *>(src);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_vec16(uint32_t dst, const double *src, uint32_t mbar)
{
    const double2 f = *reinterpret_cast<const double2 *>(src);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(dst),
                 "d"(f.x), "d"(f.y), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// GLASS tree over CNT values held one per lane in lanes 0..CNT-1 of a G-lane group; total in lane 0.
template <typename T, uint32_t CNT, uint32_t G>
__device__ __forceinline__ T glass_tree_shfl(T x, uint32_t lane_in_group)
{
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t s = CNT;
#pragma unroll
    for (int lvl = 0; lvl < 8; ++lvl) {
        if (s > 3) {
            const uint32_t odd = s & 1u;
            s = (s - odd) / 2;
            const T y = __shfl_down_sync(FULL, x, s, G);
            T z = T(0);
            if (odd) z = __shfl_sync(FULL, x, 2 * s, G);      // pre-level value of element 2s (odd fix-up)
            x = add_rn(x, y);
            if (odd && lane_in_group == 0) x = add_rn(x, z);
        }
    }
    const T y1 = __shfl_sync(FULL, x, 1, G), y2 = __shfl_sync(FULL, x, 2, G);
    if (lane_in_group == 0) {
        if (s >= 2) x = add_rn(x, y1);
        if (s >= 3) x = add_rn(x, y2);
    }
    return x;
}

template <uint32_t CNT>
__host__ __device__ constexpr uint32_t part_index(uint32_t b)
{
    return b;   // plain layout: a CTA's R partials are contiguous, so they travel as 16-byte vectors
}

// N-way GLASS tree over partials stored with part_index(); every lane returns the total.
template <typename T, uint32_t CNT>
__device__ __forceinline__ T glass_tree_part(const T *part)
{
    if constexpr (is_pow2<CNT>::value && CNT >= 32) {
        constexpr uint32_t PER = CNT / 32;
        const uint32_t lane = threadIdx.x & 31u;
        T v[PER];
#pragma unroll
        for (uint32_t q = 0; q < PER; ++q) v[q] = part[lane + 32 * q];
#pragma unroll
        for (uint32_t h = PER / 2; h >= 1; h /= 2) {
#pragma unroll
            for (uint32_t q = 0; q < PER / 2; ++q)
                if (q < h) v[q] = add_rn(v[q], v[q + h]);
        }
        T x = v[0];
        // XOR butterfly: lane l < s adds the same pair as the halving tree does (a + b == b + a bit for bit), and every
        // lane ends with the total -- no broadcast shuffle
#pragma unroll
        for (uint32_t s = 16; s >= 1; s /= 2) x = add_rn(x, __shfl_xor_sync(0xffffffffu, x, s));
        return x;
    } else {
        T v[CNT];
#pragma unroll
        for (uint32_t q = 0; q < CNT; ++q) v[q] = part[q];
        return glass_tree<T, CNT>(v);
    }
}

template <typename T, uint32_t n, uint32_t N, uint32_t C>
struct ClusterPcg2 {
    static_assert(N % C == 0 && N >= 2 && n <= 32 && C <= 16, "unsupported shape");
    static constexpr uint32_t G = n <= 16 ? 16 : 32;   // lanes per knot row
    static constexpr uint32_t R = N / C;
    static constexpr uint32_t W = 3 * n;
    static constexpr uint32_t TILE = 3 * n * n;
    static constexpr uint32_t NT = (R * G + 31) / 32 * 32;
    static constexpr uint32_t XS = (n + 3) / 4 * 4;     // padded row stride of the vector windows
    static constexpr uint32_t XLEN = (R + 2) * XS;
    static constexpr uint32_t VEC = 16 / sizeof(T);     // elements per 16-byte message
    static constexpr bool VEC_PART = R % VEC == 0;      // partials travel as vectors (else one scalar each)
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr size_t align16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;                // 3 mbarriers: tiles, phase A, phase B
    static constexpr size_t OFF_S = 32;
    static constexpr size_t OFF_P = OFF_S + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XP = OFF_P + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HU = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HT = OFF_HU + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_HS = OFF_HT + align16(sizeof(T) * 2 * XS);    // outgoing boundary rows (staging)
    static constexpr size_t OFF_PV = OFF_HS + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_PE = OFF_PV + align16(sizeof(T) * N);
    static constexpr size_t SMEM_BYTES = OFF_PE + align16(sizeof(T) * N);
};

// band-row chain over a padded window: columns c = 0..3n-1 ascending, one FMA each
template <typename T, uint32_t n, uint32_t XS>
__device__ __forceinline__ T chain_padded(const T (&m)[3 * n], const T *__restrict__ xw)
{
    T x[3 * XS];
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (uint32_t q = 0; q < 3 * XS / 4; ++q) {
            const float4 f = reinterpret_cast<const float4 *>(xw)[q];
            x[4 * q] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
        }
    } else {
#pragma unroll
        for (uint32_t q = 0; q < 3 * XS / 2; ++q) {
            const double2 f = reinterpret_cast<const double2 *>(xw)[q];
            x[2 * q] = f.x; x[2 * q + 1] = f.y;
        }
    }
    T acc = T(0);
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk)
#pragma unroll
        for (uint32_t c = 0; c < n; ++c) acc = fma_rn(m[blk * n + c], x[blk * XS + c], acc);
    return acc;
}

// mbarrier set-up of one CTA; the caller follows it with a CTA barrier and cluster_sync()
template <typename T, uint32_t n, uint32_t N, uint32_t C>
__device__ __forceinline__ void pcg_cluster_v2_init(unsigned char *smem_raw)
{
    using K = ClusterPcg2<T, n, N, C>;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        fence_mbar_init();
    }
}

// Solves systems first_sys, first_sys + sys_stride, ... < a.batch with this cluster.  Called by threads 0 .. NT-1 of every CTA of
// the cluster (EXACT_BLOCK: the launch carries exactly NT threads and the plain CTA barrier is used; otherwise a named barrier
// over NT threads, so the drop-in pcg<T,n,N> can run it under a larger caller-chosen block).
template <typename T, uint32_t n, uint32_t N, uint32_t C, bool PROF, bool EXACT_BLOCK>
__device__ __forceinline__ void pcg_cluster_v2_run(const PcgArgs<T> &a, unsigned char *smem_raw, uint32_t first_sys, uint32_t sys_stride)
{
    auto cta_sync = [&]() { if constexpr (EXACT_BLOCK) __syncthreads(); else named_bar_sync(2, ClusterPcg2<T, n, N, C>::NT); };
    using K = ClusterPcg2<T, n, N, C>;
    constexpr uint32_t R = K::R, W = K::W, TILE = K::TILE, G = K::G, XS = K::XS, VEC = K::VEC, NT = K::NT;
    constexpr uint32_t HCH = XS / VEC;             // 16-byte messages per boundary row

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    uint64_t *barT = bars, *barA = bars + 1, *barB = bars + 2;
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);   // p (prologue: lambda) rows, one halo row each side
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);   // r rows, one halo row each side
    T *hu = reinterpret_cast<T *>(smem_raw + K::OFF_HU);   // incoming upsilon boundary rows [from left | from right]
    T *ht = reinterpret_cast<T *>(smem_raw + K::OFF_HT);   // incoming r~ boundary rows
    T *hs = reinterpret_cast<T *>(smem_raw + K::OFF_HS);   // outgoing boundary rows [my first | my last]
    T *part_v = reinterpret_cast<T *>(smem_raw + K::OFF_PV);
    T *part_e = reinterpret_cast<T *>(smem_raw + K::OFF_PE);

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u;
    const bool sender = t < 32;                    // warp 0 ships this CTA's messages
    const uint32_t j = t % G;                      // lane inside the knot-row group
    const uint32_t g = t / G;
    const bool group_live = g < R;
    const uint32_t k = group_live ? g : R - 1;     // local knot row (clamped for padding lanes)
    const bool is_row = group_live && j < n;
    const uint32_t cr = cluster_ctarank();
    const uint32_t b = cr * R + k;
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    const uint32_t left = has_left ? cr - 1 : cr, right = has_right ? cr + 1 : cr;
    const bool first_row = group_live && g == 0 && j < XS, last_row = group_live && g == R - 1 && j < XS;
    const bool own_lhalo = group_live && g == 0 && j < n, own_rhalo = group_live && g == R - 1 && j < n;
    const uint32_t jn = j < n ? j : 0;

    const uint32_t nb = (has_left ? 1u : 0u) + (has_right ? 1u : 0u);
    const uint32_t halo_bytes = nb * XS * (uint32_t)sizeof(T);
    const uint32_t full_bytes = (C - 1) * R * (uint32_t)sizeof(T) + halo_bytes;

    // Ship phase: the sender warp waits until every warp has staged its partial / boundary rows, pushes
    // them to the peers (16-byte messages counted on the peers' mbarrier) and only then arms OUR
    // mbarrier, so that a completed phase also orders the locally staged partials for all local readers.
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

    uint32_t phT = 0, phA = 0, phB = 0;
    for (uint32_t sys = first_sys; sys < a.batch; sys += sys_stride) {
        const size_t moff = ((size_t)sys * N + (size_t)cr * R) * TILE;
        const size_t vbase = (size_t)sys * N * n;
        const T *gS = a.S + moff, *gP = a.Pinv + moff;
        const bool tma = K::TMA_OK && a.use_tma;

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
        // lambda window: own knot rows plus one each side (absent neighbours read as zero); pads zeroed
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) {
            const uint32_t row = i / XS, e = i % XS;
            const long kb = (long)(cr * R) + (long)row - 1;
            xp[i] = (e < n && kb >= 0 && kb < (long)N) ? a.lambda[vbase + (size_t)kb * n + e] : T(0);
            // halo rows of r with no neighbour stay zero; rows WITH a neighbour are written remotely
            // (possibly already, by a neighbour that is ahead) and must not be touched here
            if ((row == 0 && !has_left) || (row == R + 1 && !has_right)) xr[i] = T(0);
        }
        if (t < 2 * XS) hs[t] = T(0);
        T lam = T(0), gam = T(0);
        if (is_row) {
            lam = a.lambda[vbase + (size_t)b * n + j];
            gam = a.gamma[vbase + (size_t)b * n + j];
        }
        if (tma) mbar_wait(barT, phT);
        phT ^= 1u;
        cta_sync();
        if (cr == 0)
            for (uint32_t i = t; i < n * n; i += NT) { sS[i] = T(0); sP[i] = T(0); }
        if (cr == C - 1)
            for (uint32_t i = t; i < n * n; i += NT) { sS[(R - 1) * TILE + 2 * n * n + i] = T(0); sP[(R - 1) * TILE + 2 * n * n + i] = T(0); }
        cta_sync();

        // this thread's rows of S and Pinv live in registers for the whole solve
        T ms[W], mp[W];
        {
            const T *rowS = sS + k * TILE + jn, *rowP = sP + k * TILE + jn;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                ms[c] = is_row ? rowS[c * n] : T(0);
                mp[c] = is_row ? rowP[c * n] : T(0);
            }
        }
        const T *wp = xp + k * XS, *wr = xr + k * XS;      // 3-row windows [k-1 | k | k+1]
        T *own_p = xp + (k + 1) * XS + jn, *own_r = xr + (k + 1) * XS + jn;

        // ---- r = gamma - S*lambda ; exchange boundary rows of r            (pcg.cuh:118-126)
        T r = gam - chain_padded<T, n, XS>(ms, wp);
        if (is_row) *own_r = r;
        if (own_lhalo) hs[j] = r;
        if (own_rhalo) hs[XS + j] = r;
        ship(part_e, barA, xr + (R + 1) * XS, xr, false, halo_bytes);
        mbar_wait(barA, phA);
        phA ^= 1u;
        cta_sync();
        // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                             (pcg.cuh:130-149)
        T rt = chain_padded<T, n, XS>(mp, wr);
        {
            const T x = glass_tree_shfl<T, n, G>(mul_rn(r, rt), j);
            if (group_live && j == 0) part_e[b] = x;
        }
        if (own_lhalo) hs[j] = rt;
        if (own_rhalo) hs[XS + j] = rt;
        ship(part_e, barB, ht + XS, ht, true, full_bytes);
        mbar_wait(barB, phB);
        phB ^= 1u;
        T eta = glass_tree_part<T, N>(part_e);
        T p = rt, ups = T(0);
        if (is_row) *own_p = p;
        if (own_lhalo) xp[j] = has_left ? ht[j] : T(0);
        if (own_rhalo) xp[(R + 1) * XS + j] = has_right ? ht[XS + j] : T(0);

        uint32_t iter = 0;
        uint8_t max_iter_exit = 1;
        // timeline build (mode 14): %clock stamps of iterations 8..11, [iter][point][thread] (tools/timeline.py)
        auto stamp = [&](uint32_t pt, T dep) {
            if constexpr (PROF) {
                if (a.dbg && iter >= 8 && iter < 12) {
                    uint32_t c_;
                    asm volatile("mov.u32 %0, %%clock;" : "=r"(c_) : "f"((float)dep) : "memory");
                    a.dbg[((iter - 8) * 12 + pt) * (C * NT) + cr * NT + t] = c_;
                }
            }
        };
        for (; iter < a.max_iter; ++iter) {
            cta_sync();
            stamp(0, p);
            // ---- upsilon = S*p ; v = p.upsilon                             (pcg.cuh:156-167)
            ups = chain_padded<T, n, XS>(ms, wp);
            stamp(1, ups);
            {
                const T x = glass_tree_shfl<T, n, G>(mul_rn(p, ups), j);
                if (group_live && j == 0) part_v[b] = x;
                stamp(2, x);
            }
            if (own_lhalo) hs[j] = ups;
            if (own_rhalo) hs[XS + j] = ups;
            ship(part_v, barA, hu + XS, hu, true, full_bytes);
            stamp(3, ups);
            mbar_wait(barA, phA);
            phA ^= 1u;
            stamp(4, ups);
            const T alpha = eta / glass_tree_part<T, N>(part_v);               // :169
            stamp(5, alpha);
            // ---- lambda += alpha p ; r -= alpha upsilon (own rows + halo copies)   (:172-176)
            lam = fma_rn(alpha, p, lam);
            r = fma_rn(-alpha, ups, r);
            if (is_row) *own_r = r;
            if (own_lhalo && has_left) xr[j] = fma_rn(-alpha, hu[j], xr[j]);
            if (own_rhalo && has_right) xr[(R + 1) * XS + j] = fma_rn(-alpha, hu[XS + j], xr[(R + 1) * XS + j]);
            cta_sync();
            stamp(6, r);
            // ---- r~ = Pinv*r ; eta' = r.r~                                 (:180-193)
            rt = chain_padded<T, n, XS>(mp, wr);
            stamp(7, rt);
            {
                const T x = glass_tree_shfl<T, n, G>(mul_rn(r, rt), j);
                if (group_live && j == 0) part_e[b] = x;
            }
            if (own_lhalo) hs[j] = rt;
            if (own_rhalo) hs[XS + j] = rt;
            ship(part_e, barB, ht + XS, ht, true, full_bytes);
            stamp(8, rt);
            mbar_wait(barB, phB);
            phB ^= 1u;
            stamp(9, rt);
            const T eta_new = glass_tree_part<T, N>(part_e);
            stamp(10, eta_new);
            if (abs_(eta_new) < a.exit_tol) { ++iter; max_iter_exit = 0; break; }   // :195
            const T beta = eta_new / eta;                                       // :199-200
            eta = eta_new;
            // ---- p = r~ + beta p (own rows + halo copies)                   (:203-206)
            p = fma_rn(beta, p, rt);
            if (is_row) *own_p = p;
            if (own_lhalo && has_left) xp[j] = fma_rn(beta, xp[j], ht[j]);
            if (own_rhalo && has_right) xp[(R + 1) * XS + j] = fma_rn(beta, xp[(R + 1) * XS + j], ht[XS + j]);
        }

        // ---- outputs                                                        (:212-215)
        if (is_row) {
            const size_t o = vbase + (size_t)b * n + j;
            a.lambda[o] = lam;
            if (a.r_out) a.r_out[o] = r;
            if (a.p_out) a.p_out[o] = p;
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
    (void)first_row; (void)last_row;
}

// C-ABI kernel: persistent clusters looping over a batch of systems
template <typename T, uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
__global__ void __launch_bounds__(ClusterPcg2<T, n, N, C>::NT, MINB)
pcg_cluster_kernel_v2(const PcgArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_v2_init<T, n, N, C>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, all mbarriers initialised, before any DSMEM traffic
    pcg_cluster_v2_run<T, n, N, C, PROF, true>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
