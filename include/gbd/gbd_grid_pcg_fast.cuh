// gbd_grid_pcg_fast.cuh -- grid-resident GBD-PCG, TOLERANCE-PARITY ("fast") family: the whole GPU (one CTA per R knot rows)
// solves ONE system that does not fit a thread-block cluster (BASELINE config 5: n = 64, N = 256, 25 MB of tiles).
//
// Same contract as the reference pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218) and the same recurrence, band-row arithmetic and
// parity policy as gbd_cluster_pcg_fast.cuh (Chronopoulos-Gear preconditioned CG, six half-tile chains per band row, correctly
// rounded reciprocals); checked bit for bit against oracle/pcg_fast_oracle.c (G = 1: this kernel's reduction order).  Against the
// bit-exact grid kernel (gbd_grid_pcg.cuh: two all-gathers of all N per-knot partials + boundary rows through L2 per iteration,
// two 3n-long FMA chains) one iteration is:
//
//   u = Pinv r     Pinv rows live in REGISTERS (3n = 192 floats per thread as FFMA2 pairs), r window in shared memory
//   u boundary rows to the two neighbour CTAs   -- one 8-byte {value, epoch} packet per element through L2, neighbour-only
//   w = S u        S rows in shared memory (98 KB per CTA), six independent chains per row
//   ONE all-gather: every CTA's {r.u, w.u} pair as one 16-byte packet per consumer CTA (CTAS packets instead of 2 N), with the
//                   boundary rows of w riding along to the neighbours, which keep redundant r / s on their near-halo rows
//
// (The cluster kernels compute u on the halo rows redundantly and need no u exchange; here a CTA has 2 knot rows and no room for its
// neighbours' Pinv rows, so u travels -- to the neighbours only, which costs one L2 round trip instead of an all-gather.)
// Packets are double-buffered by epoch parity per channel; every poll is bounded (SpinGuard).
#pragma once
#include "gbd_grid_pcg.cuh"
#include "gbd_cluster_pcg_fast.cuh"

namespace gbd {

// CL = CTAs per thread-block cluster (1: no clusters).  With CL > 1 the all-gather is two-level: pairs meet in the cluster leader's shared
// memory (DSMEM), the CTAS / CL leaders exchange cluster sums through L2, every leader hands the total to its peers (DSMEM) -- fewer
// participants on the L2 path (its cost follows their number: profiles/r01c_micro_l2_exchange.log), same balanced tree (CL a power of two).
template <uint32_t n, uint32_t N, uint32_t R, uint32_t CL = 1>
struct GridPcgFast {
    using T = float;
    static_assert(n % 32 == 0 && n <= 64, "a knot row is n/32 whole warps; its Pinv row (3n floats) lives in registers");
    static_assert(N % R == 0 && R >= 2, "two boundary rows per CTA");
    static constexpr uint32_t CTAS = N / R;
    static_assert(CTAS >= 2 && (CTAS & (CTAS - 1)) == 0, "the pair tree assumes a power-of-two CTA count");
    static_assert(CL >= 1 && (CL & (CL - 1)) == 0 && CTAS % CL == 0 && CL <= 8, "cluster size");
    static constexpr uint32_t NCL = CTAS / CL;           // clusters = participants of the L2 exchange
    static_assert(NCL <= 32 || CL == 1, "one cluster sum per lane of the leader's first warp");
    static constexpr uint32_t NT = R * n, NW = NT / 32, H = n / 2, XS = n, TILE = 3 * n * n;
    static_assert((NW & (NW - 1)) == 0 && NW <= 32, "warp sums are added in a balanced tree");
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    // workspace of one CONSUMER CTA (u64 words): [2 parities][CTAS] 16-byte pair packets, then [2][2 sides][n] w packets, [2][2][n] u packets
    static constexpr size_t DOT_WORDS = 2 * 2 * CTAS, HALO_WORDS = 2 * 2 * n;
    static constexpr size_t REGION_WORDS = DOT_WORDS + 2 * HALO_WORDS;
    static constexpr size_t WS_WORDS = REGION_WORDS * CTAS;
    static constexpr size_t a16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_SUM = 16;                                   // [NW] {r.u, w.u} warp sums
    static constexpr size_t OFF_INBOX = OFF_SUM + a16(8 * NW);              // [2][CL] x 16 B pairs from the cluster's CTAs (leader), then [2] x 16 B total
    static constexpr size_t OFF_PAIRS = OFF_INBOX + 16 * (2 * CL + 2);      // [CTAS] {gamma_c, delta_c}   (CL > 1: [0] = the total)
    static constexpr size_t OFF_XR = OFF_PAIRS + a16(8 * CTAS);             // r rows a-1 .. a+R, interleaved pairs
    static constexpr size_t OFF_XU = OFF_XR + a16(4 * (R + 2) * XS);        // u rows a-1 .. a+R (prologue: lambda0)
    static constexpr size_t OFF_S = OFF_XU + a16(4 * (R + 2) * XS);         // S rows a .. a+R-1
    static constexpr size_t OFF_P = OFF_S + a16(4 * R * TILE);              // Pinv rows (staging only)
    static constexpr size_t SMEM_BYTES = OFF_P + a16(4 * R * TILE);
    __host__ __device__ static constexpr uint32_t pos(uint32_t e) { return e < H ? 2 * e : 2 * (e - H) + 1; }
};

__device__ __forceinline__ void st_pkt2(unsigned long long *p, float a, float b, uint32_t epoch)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(((unsigned long long)epoch << 32) | __float_as_uint(a)),
                 "l"(((unsigned long long)epoch << 32) | __float_as_uint(b))
                 : "memory");
}
__device__ __forceinline__ void ld_pkt2(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

template <uint32_t n, uint32_t N, uint32_t R, uint32_t CL = 1>
__global__ void __launch_bounds__(GridPcgFast<n, N, R, CL>::NT, 1)
pcg_grid_kernel_fast(const GridArgs<float> ga)
{
    using K = GridPcgFast<n, N, R, CL>;
    constexpr uint32_t CTAS = K::CTAS, NT = K::NT, NW = K::NW, H = K::H, XS = K::XS, TILE = K::TILE, NCL = K::NCL;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const PcgArgs<float> &a = ga.a;
    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    float2 *sums = reinterpret_cast<float2 *>(smem_raw + K::OFF_SUM);
    float2 *pairs = reinterpret_cast<float2 *>(smem_raw + K::OFF_PAIRS);
    float *xr = reinterpret_cast<float *>(smem_raw + K::OFF_XR);
    float *xu = reinterpret_cast<float *>(smem_raw + K::OFF_XU);
    float *sS = reinterpret_cast<float *>(smem_raw + K::OFF_S);
    float *sP = reinterpret_cast<float *>(smem_raw + K::OFF_P);

    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t cta = blockIdx.x;
    const uint32_t g = t / n, j = t % n;                    // own knot row of the CTA, element
    const int row_a = (int)(cta * R), b = row_a + (int)g;
    const bool has_left = cta > 0, has_right = cta + 1 < CTAS;
    const uint32_t pj = K::pos(j);
    // halo element of this thread: the threads of row 0 keep element j of row a-1, those of row R-1 element j of row a+R
    const bool lh = g == 0 && has_left, rh = g == R - 1 && has_right, hl = lh || rh;
    const uint32_t hrow = lh ? 0u : R + 1;                  // its row in the windows (rows a-1 .. a+R)
    // channels: this CTA's region, and where its boundary rows go in the neighbours' regions
    unsigned long long *const my = ga.ws + K::REGION_WORDS * cta;
    unsigned long long *const my_wh = my + K::DOT_WORDS, *const my_uh = my_wh + K::HALO_WORDS;
    unsigned long long *const nb = ga.ws + K::REGION_WORDS * (lh ? cta - 1 : (rh ? cta + 1 : cta));
    // a row-0 thread feeds its LEFT neighbour's "from the right" side (1), a row-(R-1) thread its right neighbour's side 0
    const uint32_t nb_side = lh ? 1u : 0u, my_side = lh ? 0u : 1u;
    uint32_t epoch = ga.epoch_base;

    // ---- staging: S and Pinv rows of this CTA (pad tiles zeroed), lambda0 rows a-1 .. a+R into the u window
    {
        const float *srcS = a.S + (size_t)row_a * TILE, *srcP = a.Pinv + (size_t)row_a * TILE;
        const uint32_t bytes = R * TILE * 4u;
        if (t == 0) {
            mbar_init(barT, 1);
            fence_mbar_init();
        }
        if (t < 4 * (2 * CL + 2)) reinterpret_cast<uint32_t *>(smem_raw + K::OFF_INBOX)[t] = 0u;      // epoch 0 is never sent
        __syncthreads();
        if constexpr (CL > 1) cluster_sync();               // inboxes cleared before any peer writes into them
        const bool tma = K::TMA_OK && a.use_tma;
        if (tma) {
            if (t == 0) {
                fence_proxy_async();
                constexpr uint32_t CHB = 16384;
                mbar_arrive_expect_tx(barT, 2 * bytes);
                for (uint32_t o = 0; o < bytes; o += CHB) {
                    const uint32_t len = bytes - o < CHB ? bytes - o : CHB;
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sS) + o, reinterpret_cast<const unsigned char *>(srcS) + o, len, barT);
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sP) + o, reinterpret_cast<const unsigned char *>(srcP) + o, len, barT);
                }
            }
        } else {
            for (uint32_t i = t; i < bytes / 4; i += NT) { sS[i] = srcS[i]; sP[i] = srcP[i]; }
        }
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) {
            const int kb = row_a - 1 + (int)(i / XS);
            const uint32_t e = i % XS;
            xu[(i / XS) * XS + K::pos(e)] = (kb >= 0 && kb < (int)N) ? a.lambda[(size_t)kb * n + e] : 0.f;
            xr[i] = 0.f;
        }
        if (tma) mbar_wait(barT, 0);
        __syncthreads();
        // the two tiles the reference never reads hold anything: zero them (left of knot row 0, right of knot row N-1)
        if (cta == 0) for (uint32_t i = t; i < n * n; i += NT) sS[i] = 0.f;
        if (cta == CTAS - 1) for (uint32_t i = t; i < n * n; i += NT) sS[(size_t)(R - 1) * TILE + 2 * n * n + i] = 0.f;
        __syncthreads();
    }
    // this thread's Pinv row stays in registers as pairs {m[c], m[c + H]} per tile
    f32x2 mp[3 * H];
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        const bool z = (b == 0 && blk == 0) || (b == (int)N - 1 && blk == 2);
        const float *tile = sP + (size_t)g * TILE + blk * n * n;
#pragma unroll
        for (uint32_t c = 0; c < H; ++c) mp[blk * H + c] = z ? 0ull : pack2(tile[c * n + j], tile[(c + H) * n + j]);
    }
    const float *const srow = sS + (size_t)g * TILE + j;   // S[blk][c][j] at srow[(blk * n + c) * n]

    // band row of the register matrix (u = Pinv r): three packed chains, combined as chain_pairs does
    auto chain_regs = [&](const float *xw) -> float {
        f32x2 acc[3];
#pragma unroll
        for (uint32_t blk = 0; blk < 3; ++blk) {
            f32x2 s = 0ull;
#pragma unroll
            for (uint32_t q = 0; q < H / 2; ++q) {
                const float4 f = reinterpret_cast<const float4 *>(xw + blk * XS)[q];
                const f32x2 x0 = pack2(f.x, f.y), x1 = pack2(f.z, f.w);
                s = q == 0 ? mul2(mp[blk * H], x0) : fma2(mp[blk * H + 2 * q], x0, s);
                s = fma2(mp[blk * H + 2 * q + 1], x1, s);
            }
            acc[blk] = s;
        }
        float lo, hi;
        unpack2(add2(add2(acc[0], acc[1]), acc[2]), lo, hi);
        return __fadd_rn(lo, hi);
    };
    // one tile of a band row of the shared-memory matrix (w = S u, and the prologue S lambda0): the tile's two half chains with scalar
    // FMAs (the halves of a packed FMA are independent IEEE operations: same bits as chain_pairs); xrow = the window row it multiplies
    auto chain_tile = [&](const float *xrow, uint32_t blk, float &lo, float &hi) {
        float l = 0.f, h = 0.f;
        const float *mb = srow + (size_t)blk * n * n;
#pragma unroll 8
        for (uint32_t q = 0; q < H / 2; ++q) {
            const float4 f = reinterpret_cast<const float4 *>(xrow)[q];
            const float *m = mb + (size_t)(2 * q) * n;
            const float m0 = m[0], m1 = m[n], m2 = m[(size_t)H * n], m3 = m[(size_t)(H + 1) * n];
            l = q == 0 ? __fmul_rn(m0, f.x) : __fmaf_rn(m0, f.x, l);
            h = q == 0 ? __fmul_rn(m2, f.y) : __fmaf_rn(m2, f.y, h);
            l = __fmaf_rn(m1, f.z, l);
            h = __fmaf_rn(m3, f.w, h);
        }
        lo = l;
        hi = h;
    };
    auto combine = [&](const float (&lo)[3], const float (&hi)[3]) -> float {
        return __fadd_rn(__fadd_rn(__fadd_rn(lo[0], lo[1]), lo[2]), __fadd_rn(__fadd_rn(hi[0], hi[1]), hi[2]));
    };
    auto chain_smem = [&](const float *xw) -> float {
        float lo[3], hi[3];
#pragma unroll
        for (uint32_t blk = 0; blk < 3; ++blk) chain_tile(xw + blk * XS, blk, lo[blk], hi[blk]);
        return combine(lo, hi);
    };
    auto warp_sum = [&](float v) -> float {
#pragma unroll
        for (uint32_t sft = 1; sft < 32; sft <<= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, sft));
        return v;
    };
    // neighbour-only exchange of one boundary-row element through channel `uh` (u rows; the prologue sends r0 through it)
    auto halo_exchange = [&](unsigned long long *mine, size_t chan_off, float v, uint32_t ep) -> float {
        const uint32_t par = ep & 1u;
        if (hl) Pkt<float>::put(nb + chan_off + (size_t)(par * 2 + nb_side) * n + j, v, ep);
        float got = 0.f;
        if (hl) got = Pkt<float>::get(mine + (size_t)(par * 2 + my_side) * n + j, ep);
        return got;
    };

    // ---- r = gamma - S lambda0 on the own rows; the near-halo rows of r come from the neighbours once
    const size_t o = (size_t)b * n + j;
    float x = a.lambda[o];
    float r = __fsub_rn(a.gamma[o], chain_smem(xu + g * XS));
    __syncthreads();                                        // lambda0 window dead: xu becomes the u window
    for (uint32_t i = t; i < (R + 2) * XS; i += NT) xu[i] = 0.f;
    ++epoch;
    float r2 = halo_exchange(my_uh, K::DOT_WORDS + K::HALO_WORDS, r, epoch);
    float u = 0.f, w = 0.f, w2 = 0.f, p = 0.f, s = 0.f, s2 = 0.f;
    float alpha = 0.f, beta = 0.f, gam = 0.f, den = 0.f;
    uint32_t iter = 0;
    bool first = true, done = false;
    __syncthreads();

    auto step = [&]() {
        xr[(g + 1) * XS + pj] = r;
        if (hl) xr[hrow * XS + pj] = r2;
        __syncthreads();
        u = chain_regs(xr + g * XS);
        xu[(g + 1) * XS + pj] = u;
        ++epoch;
        const uint32_t par = epoch & 1u;
        // the boundary rows of u leave for the neighbours now; the tile that needs the neighbour's row is multiplied LAST, so that the
        // L2 round trip hides behind the two tiles that only need this CTA's own rows (the three tile chains of a row are independent)
        if (hl) Pkt<float>::put(nb + K::DOT_WORDS + K::HALO_WORDS + (size_t)(par * 2 + nb_side) * n + j, u, epoch);
        // scalars that only need the previous gamma and denominator: off the dependent chain
        float rgam = first ? 0.f : rcp_fast(gam), q = __fmul_rn(den, rgam);
        asm volatile("" : "+f"(rgam), "+f"(q));
        __syncthreads();
        const bool rot = g == 0;                                 // row 0 waits for its LEFT tile: order 1, 2, 0; the other rows 0, 1, 2
        const uint32_t t0 = rot ? 1u : 0u, t1 = rot ? 2u : 1u, t2 = rot ? 0u : 2u;
        float la, ha, lb, hb, lc, hc;
        chain_tile(xu + (g + t0) * XS, t0, la, ha);
        chain_tile(xu + (g + t1) * XS, t1, lb, hb);
        if (hl) xu[hrow * XS + pj] = Pkt<float>::get(my_uh + (size_t)(par * 2 + my_side) * n + j, epoch);
        __syncthreads();
        chain_tile(xu + (g + t2) * XS, t2, lc, hc);
        {
            const float lo[3] = {rot ? lc : la, rot ? la : lb, rot ? lb : lc}, hi[3] = {rot ? hc : ha, rot ? ha : hb, rot ? hb : hc};
            w = combine(lo, hi);
        }
        if (hl) Pkt<float>::put(nb + K::DOT_WORDS + (size_t)(par * 2 + nb_side) * n + j, w, epoch);     // w boundary rows ride along
        const float sr = warp_sum(__fmul_rn(r, u)), sw = warp_sum(__fmul_rn(w, u));
        if (lane == 0) sums[warp] = make_float2(sr, sw);
        __syncthreads();
        float gam_new, del_new;
        {   // the CTA's pair: balanced tree over its warp sums (every thread, same bits)
            float vg[NW], vd[NW];
#pragma unroll
            for (uint32_t i = 0; i < NW; ++i) { const float2 f = sums[i]; vg[i] = f.x; vd[i] = f.y; }
            const float cg = tree_sum<NW>(vg), cd = tree_sum<NW>(vd);
            if constexpr (CL == 1) {
                // one 16-byte packet to every consumer CTA; then poll this CTA's own region
                for (uint32_t c = t; c < CTAS; c += NT) st_pkt2(ga.ws + K::REGION_WORDS * c + (size_t)(par * CTAS + cta) * 2, cg, cd, epoch);
                for (uint32_t c = t; c < CTAS; c += NT) {
                    unsigned long long pa, pb;
                    SpinGuard guard;
                    do {
                        guard.tick();
                        ld_pkt2(my + (size_t)(par * CTAS + c) * 2, pa, pb);
                    } while ((uint32_t)(pa >> 32) != epoch || (uint32_t)(pb >> 32) != epoch);
                    pairs[c] = make_float2(__uint_as_float((uint32_t)pa), __uint_as_float((uint32_t)pb));
                }
            } else if (warp == 0) {
                // two-level: (1) the pair to the cluster leader's inbox (DSMEM); the leader adds its cluster's CL pairs in rank order,
                // (2) exchanges cluster sums with the other leaders through L2 and adds them (one per lane + butterfly = the balanced
                // tree), (3) hands the total to every CTA of its cluster (DSMEM); (4) every CTA polls its own total slot
                const uint32_t inbox_u = smem_u32(smem_raw + K::OFF_INBOX), crank = cluster_ctarank(), cl_id = cta / CL;
                if (lane == 0) st_pair_cluster(map_to_cta(inbox_u, 0) + 16u * (par * CL + crank), cg, cd, epoch);
                if (crank == 0) {
                    float lg[CL], ld[CL];
                    bool ok;
                    SpinGuard guard;
                    do {
                        guard.tick();
                        ok = true;
#pragma unroll
                        for (uint32_t m = 0; m < CL; ++m) {
                            const uint4 q4 = ld_pair(inbox_u + 16u * (par * CL + m));
                            ok = ok && q4.y == epoch && q4.w == epoch;
                            lg[m] = __uint_as_float(q4.x);
                            ld[m] = __uint_as_float(q4.z);
                        }
                    } while (!ok);
                    const float clg = tree_sum<CL>(lg), cld = tree_sum<CL>(ld);
                    if (lane < NCL) st_pkt2(ga.ws + K::REGION_WORDS * (lane * CL) + (size_t)(par * CTAS + cl_id) * 2, clg, cld, epoch);
                    float tg = 0.f, td = 0.f;
                    if (lane < NCL) {
                        unsigned long long pa, pb;
                        SpinGuard g2;
                        do {
                            g2.tick();
                            ld_pkt2(my + (size_t)(par * CTAS + lane) * 2, pa, pb);
                        } while ((uint32_t)(pa >> 32) != epoch || (uint32_t)(pb >> 32) != epoch);
                        tg = __uint_as_float((uint32_t)pa);
                        td = __uint_as_float((uint32_t)pb);
                    }
                    tg = warp_sum(tg);
                    td = warp_sum(td);
                    if (lane < CL) st_pair_cluster(map_to_cta(inbox_u, lane) + 16u * (2 * CL + par), tg, td, epoch);
                }
                uint4 q4;
                SpinGuard g3;
                do {
                    g3.tick();
                    q4 = ld_pair(inbox_u + 16u * (2 * CL + par));
                } while (q4.y != epoch || q4.w != epoch);
                if (lane == 0) pairs[0] = make_float2(__uint_as_float(q4.x), __uint_as_float(q4.z));
            }
        }
        if (hl) w2 = Pkt<float>::get(my_wh + (size_t)(par * 2 + my_side) * n + j, epoch);
        __syncthreads();
        if constexpr (CL == 1) {
            // every warp adds the CTAS pairs in the same balanced tree (ascending CTA order): PL consecutive pairs per lane, then a butterfly
            constexpr uint32_t PL = CTAS >= 32 ? CTAS / 32 : 1;
            float vg[PL], vd[PL];
#pragma unroll
            for (uint32_t i = 0; i < PL; ++i) {
                const uint32_t c = lane * PL + i;
                const float2 f = c < CTAS ? pairs[c] : make_float2(0.f, 0.f);
                vg[i] = f.x;
                vd[i] = f.y;
            }
            gam_new = warp_sum(tree_sum<PL>(vg));
            del_new = warp_sum(tree_sum<PL>(vd));
        } else {
            const float2 f = pairs[0];
            gam_new = f.x;
            del_new = f.y;
        }
        done = !first && fabsf(gam_new) < a.exit_tol;                                        // pcg.cuh:195
        if (first) {
            beta = 0.f;
            den = del_new;
        } else {
            beta = __fmul_rn(gam_new, rgam);
            den = __fmaf_rn(-__fmul_rn(beta, gam_new), q, del_new);
        }
        alpha = __fmul_rn(gam_new, rcp_fast(den));
        gam = gam_new;
        first = false;
    };

    step();
    uint8_t max_iter_exit = 1;
    for (; iter < a.max_iter; ++iter) {
        // ---- p = u + beta p ; s = w + beta s ; lambda += alpha p ; r -= alpha s  (own row + the redundant near-halo element)
        s = __fmaf_rn(beta, s, w);
        r = __fmaf_rn(-alpha, s, r);
        s2 = __fmaf_rn(beta, s2, w2);
        r2 = __fmaf_rn(-alpha, s2, r2);
        p = __fmaf_rn(beta, p, u);
        x = __fmaf_rn(alpha, p, x);
        step();
        if (done) { ++iter; max_iter_exit = 0; break; }
    }
    a.lambda[o] = x;
    if (a.r_out) a.r_out[o] = r;
    if (a.p_out) a.p_out[o] = p;
    if (cta == 0 && t == 0) store_result(a, 0, iter, max_iter_exit);
    if constexpr (CL > 1) cluster_sync();                   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
