// EXPERIMENT, NOT BUILT: two matrix rows per thread for the single-solve fast kernel.  Bit-exact vs oracle/pcg_fast_oracle.c (G = 16), but
// measured SLOWER than one row per thread (profiles/r02_ab_fast2_two_rows_per_thread.log: 1.18 vs 0.82 us/iteration at N = 128, 0.89 vs
// 0.69 at N = 32): with 3 warps per CTA and ~250 live registers the halved window traffic is outweighed by the lost thread-level
// parallelism.  Kept as a record; include path was include/gbd/.

// gbd_cluster_pcg_fast2.cuh -- cluster-resident GBD-PCG, tolerance-parity family, TWO matrix rows per thread.
//
// Same recurrence, same exchange and -- operation by operation -- the same floating-point results as gbd_cluster_pcg_fast.cuh
// (16 lanes per knot row): checked bit for bit against the same CPU restatement (oracle/pcg_fast_oracle.c, G = 16).  What changes
// is the thread mapping.  Measured (profiles/r02_timeline_fast.log, tools/micro/chain_bench.cu): a band-row chain is bound by
// the DELIVERY of the vector window from shared memory, not by the FMAs -- with one matrix row per thread every one of the
// 16 lanes of a knot row loads the whole 3 x 64-byte window (160 threads x 192 B = 240 clk at 128 B/clk per SM for u = Pinv r
// with 8 + 2 knot rows, which is what the chain takes).  Here a knot row is an 8-lane group; lane j < n/2 owns the matrix rows j
// and j + n/2 of Pinv and of S (168 registers for n = 14) and the window it loads feeds both, which halves the shared-memory
// traffic of a band product; in the interleaved window layout {x[c], x[c + n/2]} its two elements are one 64-bit pair, so r and u
// are written and w is sent with one store per thread.
#pragma once
#include "gbd_cluster_pcg_fast.cuh"

namespace gbd {

template <uint32_t n, uint32_t N, uint32_t C>
struct ClusterPcgFast2 {
    using T = float;
    static_assert(n >= 2 && n <= 16 && n % 2 == 0, "a knot row lives in an 8-lane group; rows are held as n/2 register pairs per tile");
    static_assert(N % C == 0 && C >= 1 && C <= 16, "unsupported cluster shape");
    static constexpr uint32_t G = 8, XS = 16;
    static constexpr uint32_t H = n / 2;                 // pairs per tile: {x[c], x[c + H]}; lane j < H owns elements j and j + H
    static constexpr uint32_t R = N / C;                 // own knot rows per CTA
    static_assert(R >= 4 && R % 4 == 0, "own rows fill whole warps; two boundary rows travel each way");
    static constexpr uint32_t NOWN = R * G;              // threads of the own-row warps
    static constexpr uint32_t NT = NOWN + 32;
    static constexpr uint32_t HW = NOWN / 32;            // the halo warp (last warp of the CTA): lanes 0-7 left, 8-15 right
    static_assert(NT <= 1024, "too many knot rows per CTA");
    static constexpr uint32_t NRED = R * 16;             // parked product pairs, index 16 g + element (the layout of the 16-lane kernel)
    static constexpr uint32_t LN = 8;                    // lanes of the halo warp that add the CTA's products
    static constexpr uint32_t PPL = NRED / (2 * LN);
    static constexpr uint32_t W = 3 * n;
    static constexpr uint32_t TILE = 3 * n * n;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr uint32_t HALO_PAR = 2 * 2 * XS;     // halo packets per parity: [side][slot][window position]
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_NEXT = 8;
    static constexpr size_t OFF_SC = 16;
    static constexpr size_t OFF_DOT = 32;                // [2][C] x 16 B   {gamma, epoch, delta, epoch} from every CTA
    static constexpr size_t OFF_RED = OFF_DOT + 2 * C * 16;          // [NRED] x 8 B  {r.u, w.u} products of the own rows
    static constexpr size_t OFF_HALO = OFF_RED + NRED * 8;           // [2][2][2][XS] x 8 B  w boundary rows from the neighbours
    static constexpr size_t OFF_XL = OFF_HALO + 2 * HALO_PAR * 8;    // lambda0 rows a-3 .. a+R+2 (prologue only)
    static constexpr size_t OFF_XR = OFF_XL + sizeof(T) * (R + 6) * XS;   // r rows a-2 .. a+R+1
    static constexpr size_t OFF_XU = OFF_XR + sizeof(T) * (R + 4) * XS;   // u rows a-1 .. a+R
    static constexpr size_t OFF_S = OFF_XU + sizeof(T) * (R + 2) * XS;    // S rows a-2 .. a+R+1 (staging)
    static constexpr size_t OFF_P = OFF_S + sizeof(T) * (R + 4) * TILE;   // Pinv rows a-1 .. a+R (staging)
    static constexpr size_t SMEM_BYTES = OFF_P + sizeof(T) * (R + 2) * TILE;
};

// two band rows (elements j and j + H of one knot row) times one window; per row exactly the operations of chain_pairs
template <uint32_t n, uint32_t XS>
__device__ __forceinline__ void chain_pairs_two(const f32x2 (&m)[2 * 3 * (n / 2)], const float *__restrict__ xw, float &out0, float &out1)
{
    constexpr uint32_t H = n / 2;
    f32x2 acc[2][3];
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        f32x2 x[(H + 1) / 2 * 2];
#pragma unroll
        for (uint32_t q = 0; q < (H + 1) / 2; ++q) {
            const float4 f = reinterpret_cast<const float4 *>(xw + blk * XS)[q];
            x[2 * q] = pack2(f.x, f.y);
            x[2 * q + 1] = pack2(f.z, f.w);
        }
#pragma unroll
        for (uint32_t k = 0; k < 2; ++k) {
            f32x2 s = mul2(m[(k * 3 + blk) * H], x[0]);
#pragma unroll
            for (uint32_t c = 1; c < H; ++c) s = fma2(m[(k * 3 + blk) * H + c], x[c], s);
            acc[k][blk] = s;
        }
    }
    float lo, hi;
    unpack2(add2(add2(acc[0][0], acc[0][1]), acc[0][2]), lo, hi);
    out0 = __fadd_rn(lo, hi);
    unpack2(add2(add2(acc[1][0], acc[1][1]), acc[1][2]), lo, hi);
    out1 = __fadd_rn(lo, hi);
}

template <uint32_t n, uint32_t N, uint32_t C>
__device__ __forceinline__ void pcg_cluster_fast2_init(unsigned char *smem_raw)
{
    using K = ClusterPcgFast2<n, N, C>;
    uint32_t *z = reinterpret_cast<uint32_t *>(smem_raw);
    for (uint32_t i = threadIdx.x; i < K::OFF_XL / 4; i += blockDim.x) z[i] = 0u;       // epoch 0 is never sent
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR), 1);
        fence_mbar_init();
    }
}

// 16-byte packet pair {v0, epoch, v1, epoch} into a peer's shared memory (st_pair_cluster of gbd_cluster_pcg_fast.cuh) and its poll
__device__ __forceinline__ bool pair_ok(const uint4 &q, uint32_t ep) { return q.y == ep && q.w == ep; }

// Solves systems first_sys, first_sys + sys_stride, ... < a.batch with this cluster (or, with a.work_counter, systems drawn
// from the counter after the first one).  Called by all NT threads of every CTA, after init + CTA barrier + cluster_sync.
template <uint32_t n, uint32_t N, uint32_t C, bool EXACT_BLOCK = true>
__device__ __forceinline__ void pcg_cluster_fast2_run(const PcgArgs<float> &a, unsigned char *smem_raw, uint32_t first_sys, uint32_t sys_stride)
{
    using K = ClusterPcgFast2<n, N, C>;
    constexpr uint32_t R = K::R, TILE = K::TILE, G = K::G, XS = K::XS, NT = K::NT, HW = K::HW, LN = K::LN, PPL = K::PPL, H = K::H;
    constexpr unsigned FULL = 0xffffffffu;
    auto cta_sync = [&]() { if constexpr (EXACT_BLOCK) __syncthreads(); else named_bar_sync(3, NT); };

    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    float *xl = reinterpret_cast<float *>(smem_raw + K::OFF_XL);
    float *xr = reinterpret_cast<float *>(smem_raw + K::OFF_XR);
    float *xu = reinterpret_cast<float *>(smem_raw + K::OFF_XU);
    float *sS = reinterpret_cast<float *>(smem_raw + K::OFF_S);
    float *sP = reinterpret_cast<float *>(smem_raw + K::OFF_P);
    float2 *red = reinterpret_cast<float2 *>(smem_raw + K::OFF_RED);
    volatile float *sc = reinterpret_cast<volatile float *>(smem_raw + K::OFF_SC);
    const uint32_t dot_u = smem_u32(smem_raw + K::OFF_DOT), halo_u = smem_u32(smem_raw + K::OFF_HALO), next_u = smem_u32(smem_raw + K::OFF_NEXT);

    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t j = t % G, g = t / G;
    const uint32_t cr = cluster_ctarank();
    const bool hw = warp == HW;                            // the halo warp: groups R (row a-1, far a-2) and R+1 (row a+R, far a+R+1)
    const bool left_grp = g == R;
    const int row_a = (int)(cr * R);                       // first own knot row
    const int b = hw ? (left_grp ? row_a - 1 : row_a + (int)R) : row_a + (int)g;      // this group's knot row
    const bool live = j < H && g < R + 2 && b >= 0 && b < (int)N;
    const bool own = !hw && j < H;
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    const bool hl = hw && live;                            // live halo thread (its neighbour exists)
    const uint32_t jn = j < H ? j : 0;                     // this thread's elements: jn and jn + H = the pair at window position 2 jn
    const int b2 = left_grp ? b - 1 : b + 1;               // far halo row kept element-wise by the halo threads
    // rows of this group in the windows: xr holds rows a-2 .. a+R+1, xu rows a-1 .. a+R, xl rows a-3 .. a+R+2
    const uint32_t row_xr = hw ? (left_grp ? 1u : R + 2) : g + 2, far_xr = left_grp ? 0u : R + 3;
    const uint32_t row_xu = hw ? (left_grp ? 0u : R + 1) : g + 1;
    // w boundary rows: own rows 0, 1 go to the left neighbour's right-side slots 0 (near), 1 (far); rows R-1, R-2 to the right
    // neighbour's left-side slots 0, 1.  Halo buffer: [parity][side][slot][XS] packets of 8 bytes, indexed by window position, so
    // a thread's two elements are one 16-byte pair of packets.
    const bool send_l = own && has_left && g < 2, send_r = own && has_right && g + 2 >= R;
    const uint32_t addr_l = map_to_cta(halo_u, send_l ? cr - 1 : cr) + 8u * ((2u + (g & 1u)) * XS + 2u * jn);
    const uint32_t addr_r = map_to_cta(halo_u, send_r ? cr + 1 : cr) + 8u * (((R - 1 - g) & 1u) * XS + 2u * jn);
    const uint32_t my_halo = halo_u + 8u * ((left_grp ? 0u : 2u) * XS + 2u * jn);       // slot 0; slot 1 is XS packets on
    const uint32_t peer_dot = map_to_cta(dot_u, lane < C ? lane : cr) + 16u * cr;
    constexpr uint32_t HALO_PAR_BYTES = 8u * K::HALO_PAR;

    // halo warp: the CTA's parked {r.u, w.u} product pairs -> their two sums, identical in every lane (order of the 16-lane kernel)
    auto cta_sum = [&](float &sum_g, float &sum_d) {
        f32x2 v[PPL];
#pragma unroll
        for (uint32_t m = 0; m < PPL; ++m) {
            const float4 f = *reinterpret_cast<const float4 *>(smem_raw + K::OFF_RED + 16u * (LN * m + (lane & (LN - 1))));
            v[m] = add2(pack2(f.x, f.y), pack2(f.z, f.w));
        }
#pragma unroll
        for (uint32_t cnt = PPL; cnt > 1; cnt = (cnt + 1) / 2) {
#pragma unroll
            for (uint32_t i = 0; i < cnt / 2; ++i) v[i] = add2(v[2 * i], v[2 * i + 1]);
            if (cnt & 1u) v[cnt / 2] = v[cnt - 1];
        }
        float cg, cd;
        unpack2(v[0], cg, cd);
#pragma unroll
        for (uint32_t sft = LN / 2; sft >= 1; sft >>= 1) {
            cg = __fadd_rn(cg, __shfl_xor_sync(FULL, cg, sft));
            cd = __fadd_rn(cd, __shfl_xor_sync(FULL, cd, sft));
        }
        sum_g = cg;
        sum_d = cd;
    };

    const bool draw = a.work_counter != nullptr;
    uint32_t phT = 0, ep = 0, seq = 0;
    for (uint32_t sys = first_sys; sys < a.batch;) {
        const size_t vbase = (size_t)sys * N * n;
        const float *gS = a.S + (size_t)sys * N * TILE, *gP = a.Pinv + (size_t)sys * N * TILE;
        const bool tma = K::TMA_OK && a.use_tma;
        // staged rows: S rows [a-2, a+R+2), Pinv rows [a-1, a+R+1), clipped to the system
        const int s_lo = row_a - 2 < 0 ? 0 : row_a - 2, s_hi = row_a + (int)R + 2 > (int)N ? (int)N : row_a + (int)R + 2;
        const int p_lo = row_a - 1 < 0 ? 0 : row_a - 1, p_hi = row_a + (int)R + 1 > (int)N ? (int)N : row_a + (int)R + 1;
        float *dS = sS + (size_t)(s_lo - (row_a - 2)) * TILE, *dP = sP + (size_t)(p_lo - (row_a - 1)) * TILE;
        const float *srcS = gS + (size_t)s_lo * TILE, *srcP = gP + (size_t)p_lo * TILE;
        const uint32_t bytesS = (uint32_t)(s_hi - s_lo) * TILE * 4u, bytesP = (uint32_t)(p_hi - p_lo) * TILE * 4u;
        if (tma) {
            if (t == 0) {
                fence_proxy_async();
                constexpr uint32_t CHB = 16384;
                mbar_arrive_expect_tx(barT, bytesS + bytesP);
                for (uint32_t o = 0; o < bytesS; o += CHB)
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(dS) + o, reinterpret_cast<const unsigned char *>(srcS) + o,
                                 bytesS - o < CHB ? bytesS - o : CHB, barT);
                for (uint32_t o = 0; o < bytesP; o += CHB)
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(dP) + o, reinterpret_cast<const unsigned char *>(srcP) + o,
                                 bytesP - o < CHB ? bytesP - o : CHB, barT);
            }
        } else {
            for (uint32_t i = t; i < bytesS / 4; i += NT) dS[i] = srcS[i];
            for (uint32_t i = t; i < bytesP / 4; i += NT) dP[i] = srcP[i];
        }
        // lambda0 window rows a-3 .. a+R+2 in the interleaved layout (rows outside the system and the pad slots read as zero);
        // r and u windows cleared; the parked products of the elements that do not exist (14, 15 of a row) stay zero for the whole solve
        for (uint32_t i = t; i < (R + 6) * XS; i += NT) {
            const int kb = row_a - 3 + (int)(i / XS);
            const uint32_t e = i % XS;
            xl[(i / XS) * XS + (e < n ? ClusterPcgFast<n, N, C>::pos(e) : e)] = (e < n && kb >= 0 && kb < (int)N) ? a.lambda[vbase + (size_t)kb * n + e] : 0.f;
        }
        for (uint32_t i = t; i < (R + 4) * XS; i += NT) xr[i] = 0.f;
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) xu[i] = 0.f;
        for (uint32_t i = t; i < K::NRED; i += NT) red[i] = make_float2(0.f, 0.f);
        // element k of this thread is jn + k H, k = 0, 1
        // second recurrence pair of a thread, one form for both roles (v1 = beta v1 + src ; v2 += sa v1):
        //   own rows:  src = u,  v1 = p,  v2 = lambda, sa = +alpha        halo rows (far row):  src = w2, v1 = s2, v2 = r2, sa = -alpha
        float v1[2] = {0.f, 0.f}, v2[2] = {0.f, 0.f};
        float gam_rhs[2] = {0.f, 0.f}, gam_rhs2[2] = {0.f, 0.f};
#pragma unroll
        for (uint32_t k = 0; k < 2; ++k) {
            if (own) v2[k] = a.lambda[vbase + (size_t)b * n + jn + k * H];
            if (live) gam_rhs[k] = a.gamma[vbase + (size_t)b * n + jn + k * H];
            if (hl) gam_rhs2[k] = a.gamma[vbase + (size_t)b2 * n + jn + k * H];
        }
        if (tma) mbar_wait(barT, phT);
        phT ^= 1u;
        cta_sync();

        // this thread's two rows of Pinv (every live group) and of S (own rows) stay in registers for the whole solve, as pairs
        f32x2 mp[2 * 3 * H], ms[2 * 3 * H];
        float r[2];
        {
            f32x2 m1[2 * 3 * H];
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k) {
                f32x2 mk[3 * H];
                lift_row_pairs<n, N>(mk, sP + (size_t)row_xu * TILE, b, jn + k * H, live);
#pragma unroll
                for (uint32_t c = 0; c < 3 * H; ++c) mp[k * 3 * H + c] = mk[c];
                lift_row_pairs<n, N>(mk, sS + (size_t)row_xr * TILE, b, jn + k * H, live);
#pragma unroll
                for (uint32_t c = 0; c < 3 * H; ++c) m1[k * 3 * H + c] = mk[c];
            }
            // ---- r = gamma - S*lambda on the own rows AND on the two halo rows each side       (pcg.cuh:118-126)
            float t0, t1;
            chain_pairs_two<n, XS>(m1, xl + row_xr * XS, t0, t1);
            r[0] = __fsub_rn(gam_rhs[0], t0);
            r[1] = __fsub_rn(gam_rhs[1], t1);
#pragma unroll
            for (uint32_t c = 0; c < 2 * 3 * H; ++c) ms[c] = own ? m1[c] : 0ull;
            if (hl) {
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k) {
                    f32x2 mk[3 * H];
                    lift_row_pairs<n, N>(mk, sS + (size_t)far_xr * TILE, b2, jn + k * H, true);
#pragma unroll
                    for (uint32_t c = 0; c < 3 * H; ++c) m1[k * 3 * H + c] = mk[c];
                }
                chain_pairs_two<n, XS>(m1, xl + far_xr * XS, t0, t1);
                v2[0] = __fsub_rn(gam_rhs2[0], t0);
                v2[1] = __fsub_rn(gam_rhs2[1], t1);
            }
        }
        float u[2] = {0.f, 0.f}, w[2] = {0.f, 0.f}, w2[2] = {0.f, 0.f}, s[2] = {0.f, 0.f};
        float alpha = 0.f, beta = 0.f;
        float gam = 0.f, den = 0.f;                                 // halo warp only: current gamma and CG denominator (1/alpha = den/gam)
        uint32_t iter = 0;
        bool first = true, done = false;
        float2 *const xr_own = reinterpret_cast<float2 *>(xr + row_xr * XS) + jn, *const xr_far_p = reinterpret_cast<float2 *>(xr + far_xr * XS) + jn,
                     *const xu_own = reinterpret_cast<float2 *>(xu + row_xu * XS) + jn;
        const float *const win_r = xr + (row_xr - 1) * XS, *const win_u = xu + (hw ? 0u : row_xu - 1) * XS;
        float2 *const red_own = red + (own ? 16u * g + jn : 0u);    // this thread's two parked product pairs: red_own[0], red_own[H]

        auto step = [&]() {
            if (live) *xr_own = make_float2(r[0], r[1]);
            if (hl) *xr_far_p = make_float2(v2[0], v2[1]);
            cta_sync();
            chain_pairs_two<n, XS>(mp, win_r, u[0], u[1]);
            if (live) *xu_own = make_float2(u[0], u[1]);
            if (own) {
                red_own[0].x = __fmul_rn(r[0], u[0]);
                red_own[H].x = __fmul_rn(r[1], u[1]);
            }
            cta_sync();
            ++ep;
            const uint32_t par = ep & 1u;
            if (!hw) {
                float wn0, wn1;
                chain_pairs_two<n, XS>(ms, win_u, wn0, wn1);
                if (own) {
                    red_own[0].y = __fmul_rn(wn0, u[0]);
                    red_own[H].y = __fmul_rn(wn1, u[1]);
                }
                named_bar_arrive(1, NT);
                if (send_l) st_pair_cluster(addr_l + par * HALO_PAR_BYTES, wn0, wn1, ep);
                if (send_r) st_pair_cluster(addr_r + par * HALO_PAR_BYTES, wn0, wn1, ep);
                w[0] = wn0;
                w[1] = wn1;
                named_bar_sync(2, NT);                               // sleep until the halo warp has published the scalars
                alpha = sc[0];
                beta = sc[1];
                done = sc[2] != 0.f;
            } else {
                // scalars that only need the previous gamma and denominator: off the dependent chain
                float rgam = first ? 0.f : rcp_fast(gam), q = __fmul_rn(den, rgam);            // q = 1 / alpha
                asm volatile("" : "+f"(rgam), "+f"(q));              // computed HERE, not sunk below the exchange to their first use
                named_bar_sync(1, NT);                               // the own-row warps have parked their products
                float cg, cd;
                cta_sum(cg, cd);
                if (lane < C) st_pair_cluster(peer_dot + 16u * (par * C), cg, cd, ep);
                // gather the C pairs: every lane reads all of them
                uint4 k0 = make_uint4(0, 0, 0, 0), k1 = make_uint4(0, 0, 0, 0);
                float gam_new, del_new;
                bool ok;
                uint32_t spins = 0;
                do {
                    ok = true;
                    // the C pairs are added as they are read, in the balanced tree of tree_sum (ascending CTA order): a stack of
                    // partial sums, one per level, instead of 2 C live registers
                    float sg[5], sd[5];
#pragma unroll
                    for (uint32_t m = 0; m < C; ++m) {
                        const uint4 qd = ld_pair(dot_u + 16u * (par * C + m));
                        ok = ok && qd.y == ep && qd.w == ep;
                        float cg = __uint_as_float(qd.x), cd = __uint_as_float(qd.z);
#pragma unroll
                        for (uint32_t lv = 0; lv < 5; ++lv) {
                            if ((m >> lv) & 1u) {                    // a left sibling of this level is waiting: combine and carry up
                                cg = __fadd_rn(sg[lv], cg);
                                cd = __fadd_rn(sd[lv], cd);
                            } else {
                                sg[lv] = cg;
                                sd[lv] = cd;
                                break;
                            }
                        }
                    }
                    static_assert((C & (C - 1)) == 0, "the streaming tree assumes a power-of-two cluster size");
                    constexpr uint32_t TOP = C == 1 ? 0 : (C == 2 ? 1 : (C == 4 ? 2 : (C == 8 ? 3 : 4)));
                    gam_new = sg[TOP];
                    del_new = sd[TOP];
                    if (hl) {                                        // touched in the same rounds, but the exit does not wait for them
                        k0 = ld_pair(my_halo + par * HALO_PAR_BYTES);
                        k1 = ld_pair(my_halo + par * HALO_PAR_BYTES + 8u * XS);
                    }
                    if (++spins > (1u << 24)) __trap();              // a lost packet is an error (launch failure), not a hang
                } while (!ok);
                done = !first && fabsf(gam_new) < a.exit_tol;                                // pcg.cuh:195
                if (first) {
                    beta = 0.f;
                    den = del_new;
                } else {
                    beta = __fmul_rn(gam_new, rgam);
                    den = __fmaf_rn(-__fmul_rn(beta, gam_new), q, del_new);
                }
                alpha = __fmul_rn(gam_new, rcp_fast(den));
                gam = gam_new;
                if (lane == 0) { sc[0] = alpha; sc[1] = beta; sc[2] = done ? 1.f : 0.f; }
                named_bar_arrive(2, NT);                             // the own-row warps go on with their updates ...
                // ... while this warp finishes the wait for ITS boundary elements of w: only its own update needs them
                if (hl) {
                    uint32_t spins2 = 0;
                    while (!(pair_ok(k0, ep) && pair_ok(k1, ep))) {
                        k0 = ld_pair(my_halo + par * HALO_PAR_BYTES);
                        k1 = ld_pair(my_halo + par * HALO_PAR_BYTES + 8u * XS);
                        if (++spins2 > (1u << 24)) __trap();
                    }
                    w[0] = __uint_as_float(k0.x);
                    w[1] = __uint_as_float(k0.z);
                    w2[0] = __uint_as_float(k1.x);
                    w2[1] = __uint_as_float(k1.z);
                }
            }
            first = false;
        };

        step();
        if (draw && cr == 0 && t == 0) {
            // every CTA has entered this solve (its partials arrived), so it has consumed the previous post
            ++seq;
            const uint32_t nx = atomicAdd(a.work_counter, 1u) + sys_stride;
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) st_packet<false>(map_to_cta(next_u, c), __uint_as_float(nx), seq);
        } else if (draw) {
            ++seq;
        }
        uint8_t max_iter_exit = 1;
        for (; iter < a.max_iter; ++iter) {
            // ---- p = u + beta p ; s = w + beta s ; lambda += alpha p ; r -= alpha s  (own rows + halo copies)
            const float sa = hw ? -alpha : alpha;
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k) {
                s[k] = __fmaf_rn(beta, s[k], w[k]);
                r[k] = __fmaf_rn(-alpha, s[k], r[k]);
                v1[k] = __fmaf_rn(beta, v1[k], hw ? w2[k] : u[k]);
                v2[k] = __fmaf_rn(sa, v1[k], v2[k]);
            }
            step();
            if (done) { ++iter; max_iter_exit = 0; break; }
        }

        // ---- outputs                                                        (pcg.cuh:212-215)
        if (own) {
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k) {
                const size_t o = vbase + (size_t)b * n + jn + k * H;
                a.lambda[o] = v2[k];
                if (a.r_out) a.r_out[o] = r[k];
                if (a.p_out) a.p_out[o] = v1[k];
            }
        }
        if (cr == 0 && t == 0) store_result(a, sys, iter, max_iter_exit);
        cta_sync();
        if (draw) {
            uint64_t qn;
            uint32_t spins = 0;
            do {
                qn = ld_packet_local(next_u);
                if (++spins > (1u << 26)) __trap();
            } while (!packet_ok(qn, seq));
            sys = __float_as_uint(packet_val(qn));
        } else {
            sys += sys_stride;
        }
    }
}

// C-ABI kernel: persistent clusters looping over a batch of systems
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
__global__ void __launch_bounds__(ClusterPcgFast2<n, N, C>::NT, MINB)
pcg_cluster_kernel_fast2(const PcgArgs<float> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_fast2_init<n, N, C>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, packet buffers cleared, before any DSMEM traffic
    pcg_cluster_fast2_run<n, N, C>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
