// gbd_grid_pcg.cuh -- grid-resident GBD-PCG: the whole GPU (one CTA per R knot rows) solves ONE system.
//
// Used where a system does not fit one thread-block cluster (n = 64, N = 256: 25 MB of tiles) and as
// the kernel behind the drop-in pcg<T,n,N> template, which must run under the REFERENCE's launch
// geometry (cooperative grid of N CTAs, caller-chosen block size; include/pcg/sqp.cuh:230).
// Same inputs/outputs and the same floating-point operation order as the reference kernel
// (GBD-PCG/include/pcg.cuh:54-218), so results are bit-identical; what differs:
//
//  * 2 synchronisation points per iteration instead of 4 grid.sync(): the halo vectors p / r are
//    redundant copies that every CTA updates itself (same operations => same bits), so only the
//    boundary rows of upsilon / r~ travel, together with the dot partials.
//  * no barrier object and no fence: every travelling value is an 8-byte "packet" {epoch, value}
//    written with one relaxed store and polled with relaxed loads.  A packet validates itself, so
//    there is no ordering between packets to enforce, no atomics and no __threadfence; the epoch
//    advances by one per phase and never repeats within 2^32 phases.
//  * the N-way sum is done once per warp with register adds + shuffles in the reference's tree
//    order instead of 7 __syncthreads levels in shared memory; the per-knot sum with shuffles.
//  * tiles are staged once by TMA bulk copies and (n <= 32) kept in registers.
#pragma once
#include "gbd_cluster_pcg_v2.cuh"

namespace gbd {

// ---- packets: {epoch:32 | payload:32}; a double travels as two packets
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
template <typename T>
struct Pkt;
template <>
struct Pkt<float> {
    static constexpr uint32_t WORDS = 1;
    static __device__ __forceinline__ void put(unsigned long long *slot, float v, uint32_t epoch)
    {
        st_pkt(slot, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v));
    }
    static __device__ __forceinline__ float get(const unsigned long long *slot, uint32_t epoch)
    {
        unsigned long long w;
        SpinGuard guard;
        do { guard.tick(); w = ld_pkt(slot); } while ((uint32_t)(w >> 32) != epoch);
        return __uint_as_float((uint32_t)w);
    }
};
template <>
struct Pkt<double> {
    static constexpr uint32_t WORDS = 2;
    static __device__ __forceinline__ void put(unsigned long long *slot, double v, uint32_t epoch)
    {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v), e = (unsigned long long)epoch << 32;
        st_pkt(slot, e | (b & 0xffffffffull));
        st_pkt(slot + 1, e | (b >> 32));
    }
    static __device__ __forceinline__ double get(const unsigned long long *slot, uint32_t epoch)
    {
        unsigned long long lo, hi;
        SpinGuard guard;
        do { guard.tick(); lo = ld_pkt(slot); } while ((uint32_t)(lo >> 32) != epoch);
        do { guard.tick(); hi = ld_pkt(slot + 1); } while ((uint32_t)(hi >> 32) != epoch);
        return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
    }
};

template <typename T, uint32_t n, uint32_t N, uint32_t R>
struct GridPcg {
    static_assert(N % R == 0 && N / R >= 2, "need at least two CTAs");
    static constexpr uint32_t CTAS = N / R;
    static constexpr bool SMALL = n <= 32;                       // knot row inside one warp, tiles in registers
    static constexpr uint32_t G = n <= 16 ? 16 : (n <= 32 ? 32 : (n + 31) / 32 * 32);   // lanes per knot row
    static constexpr uint32_t NT_MIN = R * G;                    // threads that own matrix rows
    static constexpr uint32_t W = 3 * n, TILE = 3 * n * n;
    static constexpr uint32_t XS = (n + 3) / 4 * 4;
    static constexpr uint32_t XLEN = (R + 2) * XS;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr uint32_t PW = Pkt<T>::WORDS;
    // workspace (u64 words): per phase type {A, B}, per CONSUMER CTA a private region of N partial packets + the
    // left neighbour's last row + the right neighbour's first row (n packets each).  Producers write one copy of
    // a partial per consumer, so no line is ever polled by more than one CTA: with one shared copy all CTAS x NT
    // threads spin on the same N/16 lines and the L2 slice serialises them (measured, tools/micro/l2_exchange.cu:
    // 7200 -> 3000 cycles per all-gather at 128 CTAs, N = 256).
    static constexpr size_t REGION_WORDS = (size_t)PW * (N + 2 * n);
    static constexpr size_t PH_WORDS = REGION_WORDS * CTAS;
    static constexpr size_t WS_WORDS = 2 * PH_WORDS;
    static constexpr size_t align16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_S = 16;
    static constexpr size_t OFF_P = OFF_S + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XP = OFF_P + align16(sizeof(T) * R * TILE);
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HU = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_HT = OFF_HU + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_PROD = OFF_HT + align16(sizeof(T) * 2 * XS);
    static constexpr size_t OFF_PART = OFF_PROD + align16(sizeof(T) * R * G);
    static constexpr size_t SMEM_BYTES = OFF_PART + align16(sizeof(T) * N);
};

// chain over a padded window with the matrix row in shared memory (column stride n)
template <typename T, uint32_t n, uint32_t XS>
__device__ __forceinline__ T chain_padded_smem(const T *__restrict__ mrow, const T *__restrict__ xw)
{
    T acc = T(0);
    constexpr uint32_t V = 16 / sizeof(T);              // vector elements per 128-bit window load
    constexpr uint32_t CB = 4 * V;                       // columns per register block of the window
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        // fully unrolled: the matrix loads of later column blocks are issued while the FMA chain of the current one runs
#pragma unroll
        for (uint32_t c0 = 0; c0 + CB <= n; c0 += CB) {
            T x[CB];
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) {
                if constexpr (sizeof(T) == 4) {
                    const float4 f = *reinterpret_cast<const float4 *>(xw + blk * XS + c0 + 4 * q);
                    x[4 * q] = f.x; x[4 * q + 1] = f.y; x[4 * q + 2] = f.z; x[4 * q + 3] = f.w;
                } else {
                    const double2 f = *reinterpret_cast<const double2 *>(xw + blk * XS + c0 + 2 * q);
                    x[2 * q] = f.x; x[2 * q + 1] = f.y;
                }
            }
#pragma unroll
            for (uint32_t c = 0; c < CB; ++c) acc = fma_rn(mrow[(blk * n + c0 + c) * n], x[c], acc);
        }
#pragma unroll 4
        for (uint32_t c = n / CB * CB; c < n; ++c) acc = fma_rn(mrow[(blk * n + c) * n], xw[blk * XS + c], acc);
    }
    return acc;
}

// The body is a __device__ function so that both the C-ABI kernel and the drop-in pcg<T,n,N> template can wrap it.
//   ws          workspace of GridPcg::WS_WORDS u64 (zero-initialised once; reused across launches)
//   epoch_base  every phase of this launch uses epochs epoch_base+1, +2, ...; returns the last epoch used
template <typename T, uint32_t n, uint32_t N, uint32_t R>
__device__ __forceinline__ uint32_t pcg_grid_body(const T *__restrict__ gS_all, const T *__restrict__ gP_all,
                                                  const T *__restrict__ g_gamma, T *g_lambda, T *r_out, T *p_out,
                                                  uint32_t *d_iters, uint8_t *d_flag, uint32_t max_iter, T exit_tol,
                                                  unsigned long long *ws, uint32_t epoch_base, bool use_tma,
                                                  unsigned char *smem_raw)
{
    using K = GridPcg<T, n, N, R>;
    constexpr uint32_t W = K::W, TILE = K::TILE, G = K::G, XS = K::XS, PW = K::PW, CTAS = K::CTAS;
    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);
    T *hu = reinterpret_cast<T *>(smem_raw + K::OFF_HU);
    T *ht = reinterpret_cast<T *>(smem_raw + K::OFF_HT);
    T *prod = reinterpret_cast<T *>(smem_raw + K::OFF_PROD);
    T *part = reinterpret_cast<T *>(smem_raw + K::OFF_PART);

    const uint32_t t = threadIdx.x, bd = blockDim.x;
    const uint32_t cta = blockIdx.x;
    const uint32_t j = t % G, g = t / G;
    const bool group_live = g < R;
    const uint32_t k = group_live ? g : R - 1;
    const bool is_row = group_live && j < n;
    const uint32_t jn = j < n ? j : 0;
    const uint32_t b = cta * R + k;
    const bool has_left = cta > 0, has_right = cta + 1 < CTAS;
    const bool own_lhalo = group_live && g == 0 && j < n, own_rhalo = group_live && g == R - 1 && j < n;

    unsigned long long *wsA = ws, *wsB = ws + K::PH_WORDS;
    // region of consumer CTA c: [N partials | n from its left neighbour | n from its right neighbour]
    auto part_slot = [&](unsigned long long *base, uint32_t c, uint32_t i) { return base + K::REGION_WORDS * c + (size_t)PW * i; };
    auto row_slot = [&](unsigned long long *base, uint32_t c, uint32_t which, uint32_t e) {
        return base + K::REGION_WORDS * c + (size_t)PW * (N + (size_t)which * n + e);
    };
    uint32_t epoch = epoch_base;

    // publish this CTA's boundary rows (and optionally its partials); poll everyone's
    auto exchange = [&](unsigned long long *base, T myval, T mypartial, bool with_partials, T *halo_in) {   // mypartial by value
        ++epoch;
        {   // one copy of this knot row's partial per consumer CTA, spread over the lanes that hold it
            bool sender;
            uint32_t lanes, idx;
            if constexpr (K::SMALL) {
                mypartial = __shfl_sync(0xffffffffu, mypartial, 0, G);
                sender = group_live; lanes = G; idx = j;
            } else if constexpr (n == 64) {
                sender = group_live && j < 32; lanes = 32; idx = j;        // knot_dot leaves the total in lanes 0..31
            } else {
                sender = group_live && j == 0; lanes = 1; idx = 0;
            }
            if (with_partials && sender)
                for (uint32_t c = idx; c < CTAS; c += lanes) Pkt<T>::put(part_slot(base, c, b), mypartial, epoch);
        }
        if (own_lhalo && has_left) Pkt<T>::put(row_slot(base, cta - 1, 1, j), myval, epoch);     // I am its right neighbour
        if (own_rhalo && has_right) Pkt<T>::put(row_slot(base, cta + 1, 0, j), myval, epoch);    // I am its left neighbour
        constexpr uint32_t QMAX = 4, HMAX = 2;
        if (PW == 1 && N <= QMAX * bd && 2 * n <= HMAX * bd) {
            // one combined poll: all of this thread's packets are in flight together and re-read until every
            // one carries this phase's epoch (a sequential spin per packet costs one L2 round trip each)
            unsigned long long w[QMAX], wh[HMAX];
            bool hlive[HMAX];
#pragma unroll
            for (uint32_t q = 0; q < HMAX; ++q) {
                const uint32_t i = t + q * bd;
                hlive[q] = i < 2 * n && (i < n ? has_left : has_right);
            }
            const unsigned long long *mine = part_slot(base, cta, 0);     // [N partials | n from left | n from right]
            bool ok;
            SpinGuard guard;
            do {
                guard.tick();
                ok = true;
#pragma unroll
                for (uint32_t q = 0; q < QMAX; ++q) {
                    const uint32_t i = t + q * bd;
                    if (with_partials && i < N) {
                        w[q] = ld_pkt(mine + i);
                        ok = ok && (uint32_t)(w[q] >> 32) == epoch;
                    }
                }
#pragma unroll
                for (uint32_t q = 0; q < HMAX; ++q)
                    if (hlive[q]) {
                        wh[q] = ld_pkt(mine + N + t + q * bd);
                        ok = ok && (uint32_t)(wh[q] >> 32) == epoch;
                    }
            } while (!ok);
#pragma unroll
            for (uint32_t q = 0; q < QMAX; ++q) {
                const uint32_t i = t + q * bd;
                if (with_partials && i < N) part[i] = __uint_as_float((uint32_t)w[q]);
            }
#pragma unroll
            for (uint32_t q = 0; q < HMAX; ++q) {
                const uint32_t i = t + q * bd;
                if (hlive[q]) halo_in[(i < n ? 0 : XS) + (i < n ? i : i - n)] = __uint_as_float((uint32_t)wh[q]);
            }
        } else {
            if (with_partials)
                for (uint32_t i = t; i < N; i += bd) part[i] = Pkt<T>::get(part_slot(base, cta, i), epoch);
            // the left neighbour's LAST row and the right neighbour's FIRST row
            for (uint32_t i = t; i < 2 * n; i += bd) {
                const bool from_left = i < n;
                const uint32_t e = from_left ? i : i - n;
                if (from_left ? has_left : has_right)
                    halo_in[(from_left ? 0 : XS) + e] =
                        Pkt<T>::get(row_slot(base, cta, from_left ? 0 : 1, e), epoch);
            }
        }
        __syncthreads();
    };

    // knot-row dot partial in GLASS order; valid in lane j == 0 of each live group
    auto knot_dot = [&](T x, T y) -> T {
        const T pr = mul_rn(x, y);
        if constexpr (K::SMALL) {
            return glass_tree_shfl<T, n, G>(pr, j);
        } else if constexpr (n == 64) {
            // two warps per knot row: the first GLASS level (i, i + 32) crosses the warps through shared memory,
            // strides 16 .. 1 are XOR-butterfly shuffles inside the lower warp (a + b == b + a bit for bit)
            if (is_row && j >= 32) prod[k * G + j - 32] = pr;
            __syncthreads();
            T x = pr;
            if (j < 32) {
                x = add_rn(pr, prod[k * G + j]);
#pragma unroll
                for (uint32_t sh = 16; sh >= 1; sh /= 2) x = add_rn(x, __shfl_xor_sync(0xffffffffu, x, sh));
            }
            return x;
        } else {
            if (is_row) prod[k * G + j] = pr;
            __syncthreads();
            T out = T(0);
            if (group_live && j == 0) {
                T v[n];
#pragma unroll
                for (uint32_t i = 0; i < n; ++i) v[i] = prod[k * G + i];
                out = glass_tree<T, n>(v);
            }
            return out;
        }
    };

    // ---- stage tiles once
    const size_t moff = (size_t)cta * R * TILE;
    const T *gS = gS_all + moff, *gP = gP_all + moff;
    const bool tma = K::TMA_OK && use_tma;
    if (t == 0) {
        mbar_init(barT, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tma) {
        if (t == 0) {
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
        for (uint32_t i = t; i < R * TILE; i += bd) { sS[i] = gS[i]; sP[i] = gP[i]; }
    }
    for (uint32_t i = t; i < (R + 2) * XS; i += bd) {
        const uint32_t row = i / XS, e = i % XS;
        const long kb = (long)(cta * R) + (long)row - 1;
        xp[i] = (e < n && kb >= 0 && kb < (long)N) ? g_lambda[(size_t)kb * n + e] : T(0);
        xr[i] = T(0);
    }
    for (uint32_t i = t; i < 2 * XS; i += bd) { hu[i] = T(0); ht[i] = T(0); }
    T lam = T(0), gam = T(0);
    if (is_row) {
        lam = g_lambda[(size_t)b * n + j];
        gam = g_gamma[(size_t)b * n + j];
    }
    if (tma) mbar_wait(barT, 0);
    __syncthreads();
    if (cta == 0)
        for (uint32_t i = t; i < n * n; i += bd) { sS[i] = T(0); sP[i] = T(0); }
    if (cta == CTAS - 1)
        for (uint32_t i = t; i < n * n; i += bd) { sS[(R - 1) * TILE + 2 * n * n + i] = T(0); sP[(R - 1) * TILE + 2 * n * n + i] = T(0); }
    __syncthreads();

    T ms[K::SMALL ? W : 1], mp[K::SMALL ? W : 1];
    const T *rowS = sS + k * TILE + jn, *rowP = sP + k * TILE + jn;
    if constexpr (K::SMALL) {
#pragma unroll
        for (uint32_t c = 0; c < W; ++c) {
            ms[c] = is_row ? rowS[c * n] : T(0);
            mp[c] = is_row ? rowP[c * n] : T(0);
        }
    }
    const T *wp = xp + k * XS, *wr = xr + k * XS;
    T *own_p = xp + (k + 1) * XS + jn, *own_r = xr + (k + 1) * XS + jn;
    auto band_S = [&](const T *win) -> T {
        if constexpr (K::SMALL) return chain_padded<T, n, XS>(ms, win);
        else return is_row ? chain_padded_smem<T, n, XS>(rowS, win) : T(0);
    };
    auto band_P = [&](const T *win) -> T {
        if constexpr (K::SMALL) return chain_padded<T, n, XS>(mp, win);
        else return is_row ? chain_padded_smem<T, n, XS>(rowP, win) : T(0);
    };

    // ---- r = gamma - S*lambda ; exchange boundary rows of r                 (pcg.cuh:118-128)
    T r = gam - band_S(wp);
    if (is_row) *own_r = r;
    exchange(wsA, r, T(0), false, hu);
    if (own_lhalo && has_left) xr[j] = hu[j];
    if (own_rhalo && has_right) xr[(R + 1) * XS + j] = hu[XS + j];
    __syncthreads();
    // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                                  (pcg.cuh:130-149)
    T rt = band_P(wr);
    exchange(wsB, rt, knot_dot(r, rt), true, ht);
    T eta = glass_tree_part<T, N>(part);
    T p = rt, ups = T(0);
    if (is_row) *own_p = p;
    if (own_lhalo) xp[j] = has_left ? ht[j] : T(0);
    if (own_rhalo) xp[(R + 1) * XS + j] = has_right ? ht[XS + j] : T(0);

    uint32_t iter = 0;
    uint8_t max_iter_exit = 1;
    for (; iter < max_iter; ++iter) {
        __syncthreads();
        // ---- upsilon = S*p ; alpha = eta / (p.upsilon)                       (pcg.cuh:156-169)
        ups = band_S(wp);
        exchange(wsA, ups, knot_dot(p, ups), true, hu);
        const T alpha = eta / glass_tree_part<T, N>(part);
        // ---- lambda += alpha p ; r -= alpha upsilon (own rows + halo copies)  (:172-176)
        lam = fma_rn(alpha, p, lam);
        r = fma_rn(-alpha, ups, r);
        if (is_row) *own_r = r;
        if (own_lhalo && has_left) xr[j] = fma_rn(-alpha, hu[j], xr[j]);
        if (own_rhalo && has_right) xr[(R + 1) * XS + j] = fma_rn(-alpha, hu[XS + j], xr[(R + 1) * XS + j]);
        __syncthreads();
        // ---- r~ = Pinv*r ; eta' = r.r~                                       (:180-193)
        rt = band_P(wr);
        exchange(wsB, rt, knot_dot(r, rt), true, ht);
        const T eta_new = glass_tree_part<T, N>(part);
        if (abs_(eta_new) < exit_tol) { ++iter; max_iter_exit = 0; break; }       // :195
        const T beta = eta_new / eta;                                           // :199-200
        eta = eta_new;
        // ---- p = r~ + beta p (own rows + halo copies)                         (:203-206)
        p = fma_rn(beta, p, rt);
        if (is_row) *own_p = p;
        if (own_lhalo && has_left) xp[j] = fma_rn(beta, xp[j], ht[j]);
        if (own_rhalo && has_right) xp[(R + 1) * XS + j] = fma_rn(beta, xp[(R + 1) * XS + j], ht[XS + j]);
    }

    if (is_row) {                                                               // :212-215
        const size_t o = (size_t)b * n + j;
        g_lambda[o] = lam;
        if (r_out) r_out[o] = r;
        if (p_out) p_out[o] = p;
    }
    if (cta == 0 && t == 0) {
        d_iters[0] = iter;
        d_flag[0] = max_iter_exit;
    }
    return epoch;
}

template <typename T>
struct GridArgs {
    PcgArgs<T> a;
    unsigned long long *ws;
    uint32_t epoch_base;
};

template <typename T, uint32_t n, uint32_t N, uint32_t R>
__global__ void __launch_bounds__(GridPcg<T, n, N, R>::NT_MIN < 128 ? 128 : GridPcg<T, n, N, R>::NT_MIN, 1)
pcg_grid_kernel(const GridArgs<T> ga)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_grid_body<T, n, N, R>(ga.a.S, ga.a.Pinv, ga.a.gamma, ga.a.lambda, ga.a.r_out, ga.a.p_out, ga.a.iters,
                              ga.a.max_iter_exit, ga.a.max_iter, ga.a.exit_tol, ga.ws, ga.epoch_base, ga.a.use_tma != 0,
                              smem_raw);
}

}  // namespace gbd
