// gbd_cluster_pcg.cuh -- cluster-resident GBD-PCG: one thread-block cluster solves one system.
//
// Replaces the reference's pcg<T,n,N> kernel (GBD-PCG/include/pcg.cuh:54-218) and its helpers
// loadbdVec/bdmv (GBD-PCG/include/utils.cuh:9-85) and glass::dot/reduce
// (GLASS/src/L1/dot.cuh:52-63, reduce.cuh:5-69).  Same inputs, same outputs, same floating-point
// operation ORDER (so results are bit-identical to the reference kernel); different machine
// mapping:
//
//   reference                                  here
//   ---------                                  ----
//   1 CTA per knot row, N CTAs, cooperative    C CTAs (one cluster, C = 1..16) per system, each
//   grid, 4 grid.sync() per iteration          owning R = N/C consecutive knot rows; 2 cluster
//                                              barriers per iteration, nothing through L2
//   n of 128 threads run the 3n-long FMA       one thread per matrix row: R*n threads per CTA all
//   chain, tiles re-read from smem each time   run chains; tiles staged ONCE by TMA bulk copy and
//                                              (REGS) held in registers for the whole solve
//   halo p/r exchanged through global memory   halo vectors are kept as redundant copies that
//   + grid barrier (2 of the 4 barriers)       every CTA updates itself (same ops => same bits);
//                                              only upsilon / r~ boundary rows travel, pushed
//                                              into the neighbour's smem (DSMEM) together with
//                                              the dot partials that the barrier is needed for
//   N-way smem tree, 7 __syncthreads levels,   partials all-gathered by DSMEM stores, tree done
//   redundantly in every CTA                   per warp with register adds + shuffles (same order)
//
// A persistent loop over `batch` systems (cluster c takes systems c, c+G, ...) makes the same
// kernel the batched many-trajectory solver.
#pragma once
#include "gbd_device.cuh"

namespace gbd {

template <typename T>
struct PcgArgs {
    const T *S;        // [batch][N][3][n][n]
    const T *Pinv;     // [batch][N][3][n][n]
    const T *gamma;    // [batch][N*n]
    T *lambda;         // [batch][N*n]  in: initial guess, out: solution
    T *r_out;          // nullable; [batch][N*n] final residual   (what the reference leaves in d_r)
    T *p_out;          // nullable; [batch][N*n] final direction  (what the reference leaves in d_p)
    uint32_t *iters;   // [batch]
    uint8_t *max_iter_exit;  // [batch]  (bool in the reference; 1 byte, 0/1)
    uint32_t batch;
    uint32_t max_iter;
    T exit_tol;
    uint32_t use_tma;  // 0: plain loads (unaligned pointers)
    uint32_t *host_result = nullptr;   // nullable: mapped pinned mirror {iters, flag} of system 0 (the linsys window reads it
                                       // after one stream wait instead of two device-to-host copies)
    uint32_t *dbg = nullptr;   // timeline build only (gbd_pcg_set_debug_buffer): per-thread %clock stamps
    uint32_t *work_counter = nullptr;  // nullable, zeroed before the launch: batched launches of the v5 kernel draw each cluster's
                                       // next system from it (first come, first served) instead of a fixed stride
};

// iteration count and exit flag of system `sys` (pcg.cuh:212-215), plus the optional host mirror
template <typename T>
__device__ __forceinline__ void store_result(const PcgArgs<T> &a, uint32_t sys, uint32_t iter, uint8_t max_iter_exit)
{
    a.iters[sys] = iter;
    a.max_iter_exit[sys] = max_iter_exit;
    if (a.host_result && sys == 0) {
        volatile uint32_t *h = a.host_result;
        h[1] = max_iter_exit;
        h[0] = iter;
    }
}

template <typename T, uint32_t n, uint32_t N, uint32_t C, bool REGS>
struct ClusterPcg {
    static_assert(N % C == 0, "cluster size must divide the number of knot points");
    static_assert(N >= 2, "reference semantics need at least two knot points");
    static constexpr uint32_t R = N / C;           // knot rows per CTA
    static constexpr uint32_t ROWS = R * n;        // matrix rows (= worker threads) per CTA
    static constexpr uint32_t W = 3 * n;           // band-row width
    static constexpr uint32_t TILE = 3 * n * n;    // elements per knot row of S or Pinv
    static constexpr uint32_t NT_ROWS = (ROWS + 31) / 32 * 32;
    static constexpr uint32_t NT_HALO = (2 * n + 31) / 32 * 32;
    static constexpr uint32_t NT = NT_ROWS > NT_HALO ? NT_ROWS : NT_HALO;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr uint32_t XLEN = (R + 2) * n;  // vector with one halo knot row each side

    // shared-memory carve-up (bytes)
    static constexpr size_t align16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_S = 16;
    static constexpr size_t OFF_P = OFF_S + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XP = OFF_P + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HU = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HT = OFF_HU + align16(sizeof(T) * 2 * n);
    static constexpr size_t OFF_PROD = OFF_HT + align16(sizeof(T) * 2 * n);
    static constexpr size_t OFF_PV = OFF_PROD + align16(sizeof(T) * ROWS);
    static constexpr size_t OFF_PE = OFF_PV + align16(sizeof(T) * N);
    static constexpr size_t SMEM_BYTES = OFF_PE + align16(sizeof(T) * N);
};

// One band-row chain: acc = sum_c m[c]*x[c], c ascending, single FMA per term (utils.cuh:46-85 order).
template <typename T, uint32_t W>
__device__ __forceinline__ T chain_regs(const T (&m)[W], const T *__restrict__ xw)
{
    T acc = T(0);
#pragma unroll
    for (uint32_t c = 0; c < W; ++c) acc = fma_rn(m[c], xw[c], acc);
    return acc;
}
template <typename T, uint32_t W, uint32_t n>
__device__ __forceinline__ T chain_smem(const T *__restrict__ mrow, const T *__restrict__ xw)
{
    T acc = T(0);
#pragma unroll 14
    for (uint32_t c = 0; c < W; ++c) acc = fma_rn(mrow[c * n], xw[c], acc);
    return acc;
}

template <typename T, uint32_t n, uint32_t N, uint32_t C, bool REGS>
__global__ void __launch_bounds__(ClusterPcg<T, n, N, C, REGS>::NT, 1)
pcg_cluster_kernel(const PcgArgs<T> a)
{
    using K = ClusterPcg<T, n, N, C, REGS>;
    constexpr uint32_t R = K::R, ROWS = K::ROWS, W = K::W, TILE = K::TILE;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);   // p (and, in the prologue, lambda) with halos
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);   // r with halos
    T *hu = reinterpret_cast<T *>(smem_raw + K::OFF_HU);   // incoming upsilon boundary rows [left | right]
    T *ht = reinterpret_cast<T *>(smem_raw + K::OFF_HT);   // incoming r~ boundary rows     [left | right]
    T *prod = reinterpret_cast<T *>(smem_raw + K::OFF_PROD);
    T *part_v = reinterpret_cast<T *>(smem_raw + K::OFF_PV);
    T *part_e = reinterpret_cast<T *>(smem_raw + K::OFF_PE);

    const uint32_t t = threadIdx.x;
    const uint32_t cr = (C > 1) ? cluster_ctarank() : 0u;
    const uint32_t cid = (C > 1) ? cluster_idx() : blockIdx.x;
    const uint32_t ncl = (C > 1) ? cluster_count() : gridDim.x;
    const bool is_row = t < ROWS;
    const uint32_t k = is_row ? t / n : 0u;        // local knot row
    const uint32_t rr = is_row ? t % n : 0u;       // row inside the knot block
    const uint32_t b = cr * R + k;                 // global knot row
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    const bool is_halo = t < 2 * n;                // threads that also maintain one halo element
    const bool halo_live = is_halo && (t < n ? has_left : has_right);
    const uint32_t halo_x = (t < n) ? t : (R + 1) * n + (t - n);   // index of that element in xp/xr
    const bool is_lead = is_row && rr == 0;

    // remote (DSMEM) addresses, fixed for the whole kernel
    const uint32_t left = has_left ? cr - 1 : cr, right = has_right ? cr + 1 : cr;
    // my first knot row feeds the LEFT neighbour's right halo; my last feeds the RIGHT neighbour's left halo
    const bool push_left = is_row && k == 0 && has_left;
    const bool push_right = is_row && k == R - 1 && has_right;
    const uint32_t rem_hu_l = map_to_cta(smem_u32(hu + n + rr), left), rem_hu_r = map_to_cta(smem_u32(hu + rr), right);
    const uint32_t rem_ht_l = map_to_cta(smem_u32(ht + n + rr), left), rem_ht_r = map_to_cta(smem_u32(ht + rr), right);
    const uint32_t rem_xr_l = map_to_cta(smem_u32(xr + (R + 1) * n + rr), left), rem_xr_r = map_to_cta(smem_u32(xr + rr), right);

    if (t == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_sync();   // every CTA of the cluster is resident before any DSMEM store

    uint32_t phase = 0;
    for (uint32_t sys = cid; sys < a.batch; sys += ncl) {
        const size_t moff = ((size_t)sys * N + (size_t)cr * R) * TILE;
        const size_t voff = (size_t)sys * N * n + (size_t)cr * ROWS;
        const T *gS = a.S + moff, *gP = a.Pinv + moff;

        // ---- stage this CTA's R band rows of S and Pinv: one TMA bulk copy stream, once per solve
        if (K::TMA_OK && a.use_tma) {
            if (t == 0) {
                fence_proxy_async();
                constexpr uint32_t total = (uint32_t)(sizeof(T) * R * TILE);
                constexpr uint32_t CH = 16384;   // bytes per bulk copy
                mbar_arrive_expect_tx(bar, 2 * total);
                for (uint32_t o = 0; o < total; o += CH) {
                    const uint32_t len = total - o < CH ? total - o : CH;
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sS) + o, reinterpret_cast<const unsigned char *>(gS) + o, len, bar);
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sP) + o, reinterpret_cast<const unsigned char *>(gP) + o, len, bar);
                }
            }
        } else {
            for (uint32_t i = t; i < R * TILE; i += K::NT) {
                sS[i] = gS[i];
                sP[i] = gP[i];
            }
        }
        // lambda window (own rows + one knot row each side; absent neighbours read as zero)
        for (uint32_t i = t; i < K::XLEN; i += K::NT) {
            const long g = (long)(cr * ROWS) + (long)i - (long)n;   // index into this system's lambda
            xp[i] = (g >= 0 && g < (long)(N * n)) ? a.lambda[(size_t)sys * N * n + g] : T(0);
        }
        if (!has_left && is_halo && t < n) xr[halo_x] = T(0);
        if (!has_right && is_halo && t >= n) xr[halo_x] = T(0);
        T lam = T(0), gam = T(0);
        if (is_row) {
            lam = a.lambda[voff + t];
            gam = a.gamma[voff + t];
        }
        if (K::TMA_OK && a.use_tma) mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();
        // the two pad tiles are never written by the producer (may hold NaN patterns): zero them
        if (cr == 0)
            for (uint32_t i = t; i < n * n; i += K::NT) { sS[i] = T(0); sP[i] = T(0); }
        if (cr == C - 1)
            for (uint32_t i = t; i < n * n; i += K::NT) { sS[(R - 1) * TILE + 2 * n * n + i] = T(0); sP[(R - 1) * TILE + 2 * n * n + i] = T(0); }
        __syncthreads();

        T ms[REGS ? W : 1], mp[REGS ? W : 1];
        const T *rowS = sS + k * TILE + rr, *rowP = sP + k * TILE + rr;
        if constexpr (REGS) {
            if (is_row) {
#pragma unroll
                for (uint32_t c = 0; c < W; ++c) { ms[c] = rowS[c * n]; mp[c] = rowP[c * n]; }
            }
        }
        auto band_S = [&](const T *x) -> T {
            if constexpr (REGS) return chain_regs<T, W>(ms, x + k * n);
            else return chain_smem<T, W, n>(rowS, x + k * n);
        };
        auto band_P = [&](const T *x) -> T {
            if constexpr (REGS) return chain_regs<T, W>(mp, x + k * n);
            else return chain_smem<T, W, n>(rowP, x + k * n);
        };
        // dot partial of this knot row (lead thread), GLASS order over n products
        auto knot_partial = [&]() -> T {
            T v[n];
#pragma unroll
            for (uint32_t i = 0; i < n; ++i) v[i] = prod[k * n + i];
            return glass_tree<T, n>(v);
        };
        auto push_partial = [&](T *part, T val) {
            const uint32_t base = smem_u32(part + b);
            if constexpr (C > 1) {
#pragma unroll
                for (uint32_t d = 0; d < C; ++d) st_cluster(map_to_cta(base, d), val);
            } else {
                part[b] = val;
            }
        };

        // ---- r = gamma - S*lambda                                     (pcg.cuh:118-126)
        T r = T(0), p = T(0), ups = T(0), rt = T(0);
        if (is_row) {
            r = gam - band_S(xp);
            xr[n + t] = r;
            if (push_left) st_cluster(rem_xr_l, r);
            if (push_right) st_cluster(rem_xr_r, r);
        }
        cluster_sync();
        // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                        (pcg.cuh:130-149)
        if (is_row) {
            rt = band_P(xr);
            prod[t] = mul_rn(r, rt);
            if (push_left) st_cluster(rem_ht_l, rt);
            if (push_right) st_cluster(rem_ht_r, rt);
        }
        __syncthreads();
        if (is_lead) push_partial(part_e, knot_partial());
        cluster_sync();
        T eta = glass_tree_smem<T, N>(part_e);
        if (is_row) { p = rt; xp[n + t] = p; }
        if (is_halo) xp[halo_x] = halo_live ? ht[t] : T(0);

        uint32_t iter = 0;
        uint8_t max_iter_exit = 1;
        for (; iter < a.max_iter; ++iter) {
            __syncthreads();
            // ---- upsilon = S*p ; v = p.upsilon                         (pcg.cuh:156-167)
            if (is_row) {
                ups = band_S(xp);
                prod[t] = mul_rn(p, ups);
                if (push_left) st_cluster(rem_hu_l, ups);
                if (push_right) st_cluster(rem_hu_r, ups);
            }
            __syncthreads();
            if (is_lead) push_partial(part_v, knot_partial());
            cluster_sync();
            const T alpha = eta / glass_tree_smem<T, N>(part_v);          // :169
            // ---- lambda += alpha p ; r -= alpha upsilon  (own rows and the halo copies)   (:172-176)
            if (is_row) {
                lam = fma_rn(alpha, p, lam);
                r = fma_rn(-alpha, ups, r);
                xr[n + t] = r;
            }
            if (halo_live) xr[halo_x] = fma_rn(-alpha, hu[t], xr[halo_x]);
            __syncthreads();
            // ---- r~ = Pinv*r ; eta' = r.r~                             (:180-193)
            if (is_row) {
                rt = band_P(xr);
                prod[t] = mul_rn(r, rt);
                if (push_left) st_cluster(rem_ht_l, rt);
                if (push_right) st_cluster(rem_ht_r, rt);
            }
            __syncthreads();
            if (is_lead) push_partial(part_e, knot_partial());
            cluster_sync();
            const T eta_new = glass_tree_smem<T, N>(part_e);
            if (abs_(eta_new) < a.exit_tol) { ++iter; max_iter_exit = 0; break; }   // :195
            const T beta = eta_new / eta;                                  // :199-200
            eta = eta_new;
            // ---- p = r~ + beta p  (own rows and the halo copies)        (:203-206)
            if (is_row) { p = fma_rn(beta, p, rt); xp[n + t] = p; }
            if (halo_live) xp[halo_x] = fma_rn(beta, xp[halo_x], ht[t]);
        }

        // ---- outputs                                                    (:212-215)
        if (is_row) {
            a.lambda[voff + t] = lam;
            if (a.r_out) a.r_out[voff + t] = r;
            if (a.p_out) a.p_out[voff + t] = p;
        }
        if (cr == 0 && t == 0) {
            store_result(a, sys, iter, max_iter_exit);
        }
        __syncthreads();   // smem of this system is dead before the next one is staged
    }
    cluster_sync();        // no CTA leaves while a neighbour could still address its smem
}

}  // namespace gbd
