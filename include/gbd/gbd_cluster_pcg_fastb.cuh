// gbd_cluster_pcg_fastb.cuh -- cluster-resident GBD-PCG, tolerance-parity family, THROUGHPUT (batch) kernel.
//
// Same contract, same recurrence and -- row by row, sum by sum -- the same floating-point operations as the packed-row
// kernels of gbd_cluster_pcg_fast.cuh (ClusterPcgFast<n, N, C, n>), so it is checked bit for bit against the same CPU
// restatement (oracle/pcg_fast_oracle.c, lanes = n) and held to the same tolerance policy against the reference
// pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218).  What changes is who computes what.
//
// Measured on the single-solve kernels (profiles/r02c_ab.log): a band-row chain is bound by operand DELIVERY, not by FMA
// issue.  With one matrix row per thread every thread loads its whole 3 x 16-float vector window (192 B) from shared memory
// for 42 FMAs; at 128 B/clk per SM, 32 knot rows x 14 threads x 192 B is 672 clk per band product.  Here a thread owns FOUR
// rows of ONE matrix (rows q, q+4, q+8, q+12 of a knot block, as 84 register pairs) and the window it loads feeds all four:
//
//   P-threads (warps 0 .. R/8-1):  4 lanes per knot row hold the Pinv rows; they run  u = Pinv r,  keep p and lambda
//   S-threads (warps R/8 .. R/4-1): 4 lanes per knot row hold the S rows;   they run  w = S u,     keep s and r
//
// so a band product costs 4 R x 192 B = 192 clk of shared-memory bandwidth at R = 32 and 84 packed FMAs per thread, and a
// CTA of 32 knot rows is 8 warps x <= 255 registers.  A system of N = 128 knots occupies 4 SMs: 37 systems in flight.
// Lane q of a knot row owns the elements q, q + n/2, q + 4, q + 4 + n/2: in the interleaved window layout {x[c], x[c + n/2]}
// these are two adjacent pairs, so a thread reads and writes its own elements of r and u with two 64-bit accesses.
// With every register holding a matrix element, anything else a thread does is serialised on shared-memory latency
// (measured, profiles/r02_timeline_fastb.log), so the dot products never pass through memory: a thread adds its four
// products, a warp (eight knot rows) adds its 32 lane sums in a shuffle butterfly, and one float per warp is parked for the
// exchange warp.  While the S-threads run their product the P-threads are idle, so P-warp 0 is the exchange warp (what the halo
// warp is in gbd_cluster_pcg_fast.cuh) and r.u is reduced then, off the dependent chain; the exchange warp sends the CTA's
// {gamma, delta} pair to every CTA, polls for the C pairs, forms alpha and beta and publishes them.  The two halo rows each
// side (redundant r, s, w; u on the near one) belong to lanes 0 .. 2n-1 of the first S-warp, which are idle during the P
// product: they compute the two near-halo rows of u from the Pinv tiles left in the staging buffer (same chain order).
// The CPU restatement of this operation order is oracle/pcg_fast_oracle.c with G = 0.
#pragma once
#include <type_traits>
#include "gbd_cluster_pcg_fast.cuh"

namespace gbd {

// UX = true: the near-halo rows of u are not recomputed by this CTA (three lanes per element from staged Pinv tiles) but TRAVEL from the
// neighbours that own them -- 16-byte {value, epoch, value, epoch} pair packets straight into the consumer's shared memory, read by
// the four S-threads of the boundary row directly as their window pairs, in the tile they multiply LAST, so the hop hides behind the
// two tiles that only need this CTA's rows.  Then only ONE redundant row of r / s each side is kept and one boundary row of w travels.
template <uint32_t n, uint32_t N, uint32_t C, bool UX = false>
struct ClusterPcgFastB {
    using T = float;
    static_assert(n >= 2 && n <= 16 && n % 2 == 0, "rows are held as n/2 register pairs per tile");
    static_assert(N % C == 0 && C >= 1 && C <= 16, "unsupported cluster shape");
    static constexpr uint32_t XS = 16, H = n / 2, Q = 4;
    static constexpr uint32_t RPT = H > 4 ? 4 : 2;       // matrix rows per thread: q, q + H | q + 4, q + 4 + H
    static constexpr uint32_t R = N / C;                 // own knot rows per CTA
    static_assert(R >= 8 && R % 8 == 0, "a warp covers eight knot rows");
    static constexpr uint32_t NP = Q * R;                // P-threads; as many S-threads
    static constexpr uint32_t NT = 2 * NP;
    static_assert(NT <= 1024 && 2 * n <= 32, "too many knot rows per CTA");
    static constexpr uint32_t SW0 = NP / 32;             // first S-warp: its lanes 0 .. 2n-1 also own the halo rows; also warps per role
    static_assert(SW0 <= 4, "the warp sums of a role are read with one 128-bit load");
    static constexpr bool SPLIT3 = !UX && C > 1 && SW0 * 10 >= 2 * n;   // near-halo u by three lanes per element (needs 3 x 2n lanes in the S-warps)
    static_assert(!UX || R >= 16, "u travels: the first and the last S-warp each hold ONE boundary row");
    static constexpr uint32_t TILE = 3 * n * n;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr uint32_t HALO_PAR = 2 * 2 * XS;     // halo packets per parity: [side][slot][XS]
    static constexpr size_t a16(size_t x) { return (x + 15) / 16 * 16; }
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_NEXT = 8;
    static constexpr size_t OFF_SC = 16;
    static constexpr size_t OFF_DOT = 32;                                   // [2][C] x 16 B {gamma, epoch, delta, epoch}
    static constexpr size_t OFF_RU = OFF_DOT + 2 * C * 16;                  // [4] r.u sums of the P-warps
    static constexpr size_t OFF_WU = OFF_RU + 16;                           // [4] w.u sums of the S-warps
    static constexpr size_t OFF_HALO = OFF_WU + 16;                         // [2][2][2][XS] x 8 B  w boundary rows from the neighbours
    static constexpr size_t OFF_UH = OFF_HALO + 2 * HALO_PAR * 8;           // UX: [2][2 sides][8] x 16 B  pair packets of the neighbours' boundary rows of u
    static constexpr size_t OFF_XL = OFF_UH + 2 * 2 * 8 * 16;               // lambda0 rows a-3 .. a+R+2 (prologue only)
    static constexpr size_t OFF_XR = OFF_XL + sizeof(T) * (R + 6) * XS;     // r rows a-2 .. a+R+1
    static constexpr size_t OFF_XU = OFF_XR + sizeof(T) * (R + 4) * XS;     // u rows a-1 .. a+R
    static constexpr size_t OFF_S = OFF_XU + sizeof(T) * (R + 2) * XS;      // S rows a-2 .. a+R+1 (staging)
    static constexpr size_t OFF_P = OFF_S + sizeof(T) * (R + 4) * TILE;     // Pinv rows a-1 .. a+R (staging; rows 0 and R+1 stay in use)
    static constexpr size_t SMEM_BYTES = OFF_P + sizeof(T) * (R + 2) * TILE;
    __host__ __device__ static constexpr uint32_t pos(uint32_t e) { return e < H ? 2 * e : 2 * (e - H) + 1; }
};

// RPT band rows times one window: row k is 3 x H register pairs {m[c], m[c + H]} at m[k * 3H ..], the window 3 knot rows of XS
// floats stored as pairs {x[c], x[c + H]}.  Per row exactly the operations of chain_pairs (gbd_cluster_pcg_fast.cuh).
template <uint32_t n, uint32_t XS, uint32_t RPT>
__device__ __forceinline__ void chain_pairs_multi(const f32x2 (&m)[RPT * 3 * (n / 2)], uint32_t xw, float (&out)[RPT])
{
    constexpr uint32_t H = n / 2;
    f32x2 acc[RPT][3];
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        f32x2 x[(H + 1) / 2 * 2];
#pragma unroll
        for (uint32_t q = 0; q < (H + 1) / 2; ++q) {
            const float4 f = lds_f32x4(xw + 4u * (blk * XS + 4u * q));
            x[2 * q] = pack2(f.x, f.y);
            x[2 * q + 1] = pack2(f.z, f.w);
        }
#pragma unroll
        for (uint32_t k = 0; k < RPT; ++k) {
            f32x2 s = mul2(m[(k * 3 + blk) * H], x[0]);
#pragma unroll
            for (uint32_t c = 1; c < H; ++c) s = fma2(m[(k * 3 + blk) * H + c], x[c], s);
            acc[k][blk] = s;
        }
    }
#pragma unroll
    for (uint32_t k = 0; k < RPT; ++k) {
        float lo, hi;
        unpack2(add2(add2(acc[k][0], acc[k][1]), acc[k][2]), lo, hi);
        out[k] = __fadd_rn(lo, hi);
    }
}

// one band row (knot row b, element j) taken from a staged tile row in shared memory, times a window: the operations of
// chain_pairs with scalar FMAs (the halves of a packed FMA are independent IEEE operations, so the bits are the same)
template <uint32_t n, uint32_t N, uint32_t XS>
__device__ __forceinline__ float chain_smem_row(uint32_t tile_row, uint32_t xw, int b, uint32_t j)
{
    constexpr uint32_t H = n / 2;
    float lo[3], hi[3];
#pragma unroll
    for (uint32_t blk = 0; blk < 3; ++blk) {
        const bool z = (b == 0 && blk == 0) || (b == (int)N - 1 && blk == 2);
        float sl = 0.f, sh = 0.f;
#pragma unroll
        for (uint32_t c = 0; c < H; ++c) {
            const float ml = z ? 0.f : lds_f32(tile_row + 4u * ((blk * n + c) * n + j)), mh = z ? 0.f : lds_f32(tile_row + 4u * ((blk * n + c + H) * n + j));
            const float2 v = lds_f32x2(xw + 4u * (blk * XS + 2 * c));
            const float vl = v.x, vh = v.y;
            sl = c == 0 ? __fmul_rn(ml, vl) : __fmaf_rn(ml, vl, sl);
            sh = c == 0 ? __fmul_rn(mh, vh) : __fmaf_rn(mh, vh, sh);
        }
        lo[blk] = sl;
        hi[blk] = sh;
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(lo[0], lo[1]), lo[2]), __fadd_rn(__fadd_rn(hi[0], hi[1]), hi[2]));
}

template <uint32_t n, uint32_t N, uint32_t C, bool UX = false>
__device__ __forceinline__ void pcg_cluster_fastb_init(unsigned char *smem_raw)
{
    using K = ClusterPcgFastB<n, N, C, UX>;
    uint32_t *z = reinterpret_cast<uint32_t *>(smem_raw);
    for (uint32_t i = threadIdx.x; i < K::OFF_XL / 4; i += blockDim.x) z[i] = 0u;       // epoch 0 is never sent
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR), 1);
        fence_mbar_init();
    }
}

template <uint32_t n, uint32_t N, uint32_t C, bool PROF = false, bool UX = false>
__device__ __forceinline__ void pcg_cluster_fastb_run(const PcgArgs<float> &a, unsigned char *smem_raw, uint32_t first_sys, uint32_t sys_stride)
{
    using K = ClusterPcgFastB<n, N, C, UX>;
    constexpr uint32_t R = K::R, TILE = K::TILE, XS = K::XS, NT = K::NT, NP = K::NP, H = K::H, RPT = K::RPT, Q = K::Q, NW = K::SW0;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t ROWB = 4u * XS;                         // bytes of one window row

    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    float *xl = reinterpret_cast<float *>(smem_raw + K::OFF_XL);          // C++ pointers: staging and prologue only
    float *xr = reinterpret_cast<float *>(smem_raw + K::OFF_XR);
    float *xu = reinterpret_cast<float *>(smem_raw + K::OFF_XU);
    float *sS = reinterpret_cast<float *>(smem_raw + K::OFF_S);
    float *sP = reinterpret_cast<float *>(smem_raw + K::OFF_P);
    // the one base address the iteration loop works from (see gbd_device.cuh: explicit shared-memory addresses)
    const uint32_t sb = opaque(smem_u32(smem_raw));

    const uint32_t t = opaque((uint32_t)threadIdx.x), lane = t & 31u, warp = t >> 5;      // (opaque: never re-read from SR_TID)
    const bool isP = t < NP;                                   // warp-uniform role
    const uint32_t tt = isP ? t : t - NP;
    const uint32_t g = tt / Q, q = tt % Q;                     // own knot row of the CTA; lane q owns the elements q, q + H, q + 4, q + 4 + H
    const uint32_t cr = opaque(cluster_ctarank());
    const int row_a = (int)(cr * R);
    const int b = row_a + (int)g;
    const bool xw = warp == 0;                                 // the exchange warp
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    // halo lanes: side 0 = left (near row a-1, far a-2), side 1 = right (near a+R, far a+R+1)
    const bool hlane = warp == K::SW0 && lane < 2 * n;
    const uint32_t hside = lane >= n ? 1u : 0u, hj = hlane ? lane - hside * n : 0u;
    const bool hl = hlane && (hside ? has_right : has_left);   // live halo lane (its neighbour exists)
    const int hb = hside ? row_a + (int)R : row_a - 1, hb2 = hside ? hb + 1 : hb - 1;
    const uint32_t hrow_xr = hside ? R + 2 : 1u, hfar_xr = hside ? R + 3 : 0u, hrow_xu = hside ? R + 1 : 0u;
    const uint32_t hpj = K::pos(hj);
    constexpr uint32_t HALO_PAR_BYTES = 8u * K::HALO_PAR;
    const uint32_t my_halo = opaque(sb + (uint32_t)K::OFF_HALO + 8u * ((hside ? 2u : 0u) * XS + hj));   // slot 0 (near); slot 1 (far) is XS packets on
    // halo lanes: their two r elements, their u element, their Pinv tile row and its r window
    const uint32_t h_xr = opaque(sb + (uint32_t)K::OFF_XR + hrow_xr * ROWB + 4u * hpj), h_xr2 = opaque(sb + (uint32_t)K::OFF_XR + hfar_xr * ROWB + 4u * hpj);
    const uint32_t h_xu = opaque(sb + (uint32_t)K::OFF_XU + hrow_xu * ROWB + 4u * hpj);
    const uint32_t h_tile = opaque(sb + (uint32_t)K::OFF_P + hrow_xu * TILE * 4u), h_win = opaque(sb + (uint32_t)K::OFF_XR + (hside ? R + 1 : 0u) * ROWB);

    // near-halo u: element ue of the 2 n halo elements is computed by three adjacent lanes of the S-warps (ten elements per warp)
    const uint32_t ue = (warp - NW) * 10u + lane / 3u, ublk = lane % 3u;
    const uint32_t uside = ue >= n ? 1u : 0u, uj = ue - uside * n;
    const bool ulive = K::SPLIT3 && !isP && lane < 30 && ue < 2 * n && (uside ? has_right : has_left);
    const uint32_t u_tile = opaque(sb + (uint32_t)K::OFF_P + (uside ? R + 1 : 0u) * TILE * 4u + 4u * (ublk * n * n + (ulive ? uj : 0u)));
    const uint32_t u_win = opaque(sb + (uint32_t)K::OFF_XR + ((uside ? R + 1 : 0u) + ublk) * ROWB);
    const uint32_t u_out = sb + (uint32_t)K::OFF_XU + (uside ? R + 1 : 0u) * ROWB + 4u * K::pos(ulive ? uj : 0u);
    // rows of this thread: elements q, q + H (one pair of the interleaved window row), q + 4, q + 4 + H (the pair 32 bytes on); rows
    // beyond n do not exist (zero matrix rows, results dropped)
    uint32_t ek[RPT];
    bool rowv[RPT];
#pragma unroll
    for (uint32_t k = 0; k < RPT; ++k) {
        ek[k] = q + (k & 1u) * H + (k >> 1) * 4u;
        rowv[k] = q + (k >> 1) * 4u < H;
    }
    const bool pair0 = q < H, pair1 = RPT > 2 && q + 4 < H;
    const uint32_t a_own_r = opaque(sb + (uint32_t)K::OFF_XR + (g + 2) * ROWB + 8u * q);           // its first pair of r; the second is 32 bytes on
    constexpr uint32_t U_MINUS_R = (uint32_t)(K::OFF_XU - K::OFF_XR) - ROWB;                           // same elements of the u window
    const uint32_t win_r = opaque(sb + (uint32_t)K::OFF_XR + (g + 1) * ROWB), win_u = win_r + U_MINUS_R;
    // S-threads of the boundary rows send w: own rows 0, 1 to the left neighbour's right-side slots 0 (near), 1 (far); rows R-1, R-2
    // to the right neighbour's left-side slots 0, 1.  Generic pointers into the peers' shared memory, formed once.
    constexpr uint32_t HROWS = UX ? 1u : 2u;                   // boundary rows of w that travel each way
    const bool send_l = !isP && has_left && g < HROWS, send_r = !isP && has_right && g + HROWS >= R;
    // UX: the P-threads of the first / last own row send u; the S-threads of those rows read the neighbour's u from packets
    const bool usend_l = UX && isP && has_left && g == 0, usend_r = UX && isP && has_right && g + 1 == R;
    const bool upk = UX && !isP && ((g == 0 && has_left) || (g + 1 == R && has_right));
    const uint32_t ublk_pk = g == 0 ? 0u : 2u;                 // the tile whose window comes from packets
    const uint32_t uh_u = sb + (uint32_t)K::OFF_UH;
    uint64_t gaddr_u = 0;
    uint32_t my_uh = 0;
    if constexpr (UX) {
        gaddr_u = opaque(cluster_generic(map_to_cta(uh_u, usend_l ? cr - 1 : (usend_r ? cr + 1 : cr)) + 16u * ((usend_l ? 8u : 0u) + q)));
        my_uh = opaque(uh_u + 16u * (g == 0 ? 0u : 8u));         // side 0 = from the left neighbour, side 1 = from the right
    }
    const uint32_t halo_u = sb + (uint32_t)K::OFF_HALO, dot_u = opaque(sb + (uint32_t)K::OFF_DOT), next_u = sb + (uint32_t)K::OFF_NEXT;
    const uint64_t gaddr_l = opaque(cluster_generic(map_to_cta(halo_u, send_l ? cr - 1 : cr) + 8u * ((2u + (g & 1u)) * XS + q)));
    const uint64_t gaddr_r = opaque(cluster_generic(map_to_cta(halo_u, send_r ? cr + 1 : cr) + 8u * (((R - 1 - g) & 1u) * XS + q)));
    constexpr uint32_t EOFF[4] = {0u, 8u * H, 32u, 32u + 8u * H};                                      // packet slot of element k, relative to q's
    const uint64_t gpeer_dot = opaque(cluster_generic(map_to_cta(dot_u, lane < C ? lane : cr) + 16u * cr));
    const uint32_t a_sc = sb + (uint32_t)K::OFF_SC;

    // timeline build: %clock stamps of iteration PROF_ITER, held in registers and written after the solve
    constexpr uint32_t PROF_ITER = 9, NSTAMP = 16;
    uint32_t tk[NSTAMP];
    if constexpr (PROF) {
#pragma unroll
        for (uint32_t i = 0; i < NSTAMP; ++i) tk[i] = 0;
    }
    bool prof_now = false;
    auto stamp = [&](uint32_t pt, float dep) {
        if constexpr (PROF) {
            uint32_t c_;
            asm volatile("mov.u32 %0, %%clock;" : "=r"(c_) : "f"(dep) : "memory");
            if (prof_now) tk[pt] = c_;
        }
    };

    // a warp's 32 lane sums -> their total in every lane (XOR butterfly: a balanced tree in lane order)
    auto warp_sum = [&](float v) -> float {
#pragma unroll
        for (uint32_t sft = 1; sft < 32; sft <<= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, sft));
        return v;
    };
    // this thread's four products of its own elements of x (r or u, read back from the window) and its last band product
    auto own_products = [&](uint32_t a_own, const float (&o)[RPT]) -> float {
        float2 x01 = make_float2(0.f, 0.f), x23 = make_float2(0.f, 0.f);
        if (pair0) x01 = lds_f32x2(a_own);
        if (pair1) x23 = lds_f32x2(a_own + 32u);
        const float p0 = pair0 ? __fmul_rn(x01.x, o[0]) : 0.f, p1 = pair0 ? __fmul_rn(x01.y, o[1]) : 0.f;
        float p2 = 0.f, p3 = 0.f;
        if constexpr (RPT > 2) {
            p2 = pair1 ? __fmul_rn(x23.x, o[2]) : 0.f;
            p3 = pair1 ? __fmul_rn(x23.y, o[3]) : 0.f;
        }
        return __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
    };
    // the exchange warp: the NW warp sums of a role -> the CTA's sum (balanced tree, ascending)
    auto role_sum = [&](uint32_t addr) -> float {
        const float4 f = lds_f32x4(addr);
        float v[4] = {f.x, f.y, f.z, f.w};
        float w[NW];
#pragma unroll
        for (uint32_t i = 0; i < NW; ++i) w[i] = v[i];
        return tree_sum<NW>(w);
    };
    const uint32_t a_ru = sb + (uint32_t)K::OFF_RU, a_wu = sb + (uint32_t)K::OFF_WU;

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
        // r and u windows cleared
        for (uint32_t i = t; i < (R + 6) * XS; i += NT) {
            const int kb = row_a - 3 + (int)(i / XS);
            const uint32_t e = i % XS;
            xl[(i / XS) * XS + (e < n ? K::pos(e) : e)] = (e < n && kb >= 0 && kb < (int)N) ? a.lambda[vbase + (size_t)kb * n + e] : 0.f;
        }
        for (uint32_t i = t; i < (R + 4) * XS; i += NT) xr[i] = 0.f;
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) xu[i] = 0.f;
        // v1 / v2: the two recurrences of this thread's rows (P-threads: p and lambda; S-threads: s and r); out: its last band product
        float v1[RPT], v2[RPT], out[RPT];
        float rhs[RPT];
#pragma unroll
        for (uint32_t k = 0; k < RPT; ++k) {
            v1[k] = 0.f;
            out[k] = 0.f;
            const size_t o = vbase + (size_t)b * n + ek[k];
            v2[k] = (isP && rowv[k]) ? a.lambda[o] : 0.f;
            rhs[k] = (!isP && rowv[k]) ? a.gamma[o] : 0.f;
        }
        float hr = 0.f, hr2 = 0.f, hs = 0.f, hs2 = 0.f, hw = 0.f, hw2 = 0.f;     // halo lanes: r, s, w of the near / far halo row
        float hrhs = 0.f, hrhs2 = 0.f;
        if (hl) {
            hrhs = a.gamma[vbase + (size_t)hb * n + hj];
            hrhs2 = a.gamma[vbase + (size_t)hb2 * n + hj];
        }
        if (tma) mbar_wait(barT, phT);
        phT ^= 1u;
        __syncthreads();

        // this thread's four rows of its matrix stay in registers for the whole solve, as pairs
        f32x2 mm[RPT * 3 * H];
#pragma unroll
        for (uint32_t k = 0; k < RPT; ++k) {
            f32x2 mk[3 * H];
            lift_row_pairs<n, N>(mk, (isP ? sP + (size_t)(g + 1) * TILE : sS + (size_t)(g + 2) * TILE), b, rowv[k] ? ek[k] : 0u, rowv[k]);
#pragma unroll
            for (uint32_t c = 0; c < 3 * H; ++c) mm[k * 3 * H + c] = mk[c];
        }
        // ---- r = gamma - S*lambda on the own rows (S-threads) and on the two halo rows each side   (pcg.cuh:118-126)
        if (!isP) {
            float sl[RPT];
            chain_pairs_multi<n, XS, RPT>(mm, sb + (uint32_t)K::OFF_XL + (g + 2) * ROWB, sl);
#pragma unroll
            for (uint32_t k = 0; k < RPT; ++k) v2[k] = __fsub_rn(rhs[k], sl[k]);
        }
        if (hl) {
            hr = __fsub_rn(hrhs, chain_smem_row<n, N, XS>(sb + (uint32_t)K::OFF_S + hrow_xr * TILE * 4u, sb + (uint32_t)K::OFF_XL + hrow_xr * ROWB, hb, hj));
            if constexpr (!UX)
                hr2 = __fsub_rn(hrhs2, chain_smem_row<n, N, XS>(sb + (uint32_t)K::OFF_S + hfar_xr * TILE * 4u, sb + (uint32_t)K::OFF_XL + hfar_xr * ROWB, hb2, hj));
        }
        float alpha = 0.f, beta = 0.f;
        float gam = 0.f, den = 0.f;                              // exchange warp: current gamma and CG denominator
        uint32_t iter = 0;
        bool first = true, done = false;

        auto step = [&]() {
            if (!isP) {
                if (pair0) sts_f32x2(a_own_r, v2[0], v2[1]);
                if constexpr (RPT > 2) {
                    if (pair1) sts_f32x2(a_own_r + 32u, v2[2], v2[3]);
                }
                if (hl) {
                    sts_f32(h_xr, hr);
                    if constexpr (!UX) sts_f32(h_xr2, hr2);
                }
            }
            __syncthreads();
            stamp(1, v2[0]);
            ++ep;
            const uint32_t par = ep & 1u;
            if (isP) {
                // ---- u = Pinv r on the own rows; the r.u products parked
                stamp(12, v2[0]);
                chain_pairs_multi<n, XS, RPT>(mm, win_r, out);
                stamp(13, out[0] + out[1]);
                if (pair0) sts_f32x2(a_own_r + U_MINUS_R, out[0], out[1]);
                if constexpr (RPT > 2) {
                    if (pair1) sts_f32x2(a_own_r + U_MINUS_R + 32u, out[2], out[3]);
                }
                if constexpr (UX) {
                    if (usend_l || usend_r) {                        // this row of u is the neighbour's near-halo row: pair q, and pair q + 4
                        const uint64_t ga = gaddr_u + 256u * par;
                        if (pair0)
                            asm volatile("st.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ga), "r"(__float_as_uint(out[0])), "r"(ep), "r"(__float_as_uint(out[1])), "r"(ep) : "memory");
                        if constexpr (RPT > 2) {
                            if (pair1)
                                asm volatile("st.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ga + 64u), "r"(__float_as_uint(out[2])), "r"(ep), "r"(__float_as_uint(out[3])), "r"(ep) : "memory");
                        }
                    }
                }
            } else if constexpr (K::SPLIT3) {
                // ---- u on the two near halo rows (their Pinv tiles are still in the staging buffer): three lanes per element, one
                // per tile, each running the tile's two half chains; then combined in the order of chain_pairs
                float sl = 0.f, sh = 0.f;
                if (ulive) {
#pragma unroll
                    for (uint32_t c = 0; c < H; ++c) {
                        const float ml = lds_f32(u_tile + 4u * (c * n)), mh = lds_f32(u_tile + 4u * ((c + H) * n));
                        const float2 v = lds_f32x2(u_win + 8u * c);
                        sl = c == 0 ? __fmul_rn(ml, v.x) : __fmaf_rn(ml, v.x, sl);
                        sh = c == 0 ? __fmul_rn(mh, v.y) : __fmaf_rn(mh, v.y, sh);
                    }
                }
                const uint32_t l0 = lane - ublk;
                const float a0 = __shfl_sync(FULL, sl, l0), a1 = __shfl_sync(FULL, sl, l0 + 1), a2 = __shfl_sync(FULL, sl, l0 + 2);
                const float b0 = __shfl_sync(FULL, sh, l0), b1 = __shfl_sync(FULL, sh, l0 + 1), b2 = __shfl_sync(FULL, sh, l0 + 2);
                if (ulive && ublk == 0) sts_f32(u_out, __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), __fadd_rn(__fadd_rn(b0, b1), b2)));
            } else if (hl) {
                // ---- u on the near halo row, one lane per element
                sts_f32(h_xu, chain_smem_row<n, N, XS>(h_tile, h_win, hb, hj));
            }
            stamp(2, out[0]);
            __syncthreads();
            stamp(3, out[0]);
            uint64_t k0 = 0, k1 = 0;                                 // halo lanes: their two w packets
            if (!isP) {
                // ---- w = S u on the own rows; the w.u products parked; boundary rows sent to the neighbours
                stamp(14, v2[0]);
                if constexpr (UX) {
                    // tile by tile: the first S-warp (it holds own row 0) multiplies its LEFT tile last, the others their right tile, so
                    // that the neighbour's row of u -- read from its packets by the four threads of the boundary row -- has time to arrive
                    f32x2 acc[RPT][3];
                    auto do_tile = [&](auto BLK) {
                        constexpr uint32_t blk = decltype(BLK)::value;
                        f32x2 x[(H + 1) / 2 * 2];
                        if (upk && ublk_pk == blk) {
                            bool ok;
                            SpinGuard guard;
                            do {
                                guard.tick();
                                ok = true;
#pragma unroll
                                for (uint32_t c = 0; c < H; ++c) {
                                    const uint4 q4 = ld_pair(my_uh + 256u * par + 16u * c);
                                    ok = ok && q4.y == ep && q4.w == ep;
                                    x[c] = pack2(__uint_as_float(q4.x), __uint_as_float(q4.z));
                                }
                            } while (!ok);
                            if constexpr (H % 2 == 1) x[H] = 0ull;
                        } else {
#pragma unroll
                            for (uint32_t qq = 0; qq < (H + 1) / 2; ++qq) {
                                const float4 f = lds_f32x4(win_u + 4u * (blk * XS + 4u * qq));
                                x[2 * qq] = pack2(f.x, f.y);
                                x[2 * qq + 1] = pack2(f.z, f.w);
                            }
                        }
#pragma unroll
                        for (uint32_t k = 0; k < RPT; ++k) {
                            f32x2 sacc = mul2(mm[(k * 3 + blk) * H], x[0]);
#pragma unroll
                            for (uint32_t c = 1; c < H; ++c) sacc = fma2(mm[(k * 3 + blk) * H + c], x[c], sacc);
                            acc[k][blk] = sacc;
                        }
                    };
                    if (warp == NW) {
                        do_tile(std::integral_constant<uint32_t, 1>{});
                        do_tile(std::integral_constant<uint32_t, 2>{});
                        do_tile(std::integral_constant<uint32_t, 0>{});
                    } else {
                        do_tile(std::integral_constant<uint32_t, 0>{});
                        do_tile(std::integral_constant<uint32_t, 1>{});
                        do_tile(std::integral_constant<uint32_t, 2>{});
                    }
#pragma unroll
                    for (uint32_t k = 0; k < RPT; ++k) {
                        float lo, hi;
                        unpack2(add2(add2(acc[k][0], acc[k][1]), acc[k][2]), lo, hi);
                        out[k] = __fadd_rn(lo, hi);
                    }
                } else {
                    chain_pairs_multi<n, XS, RPT>(mm, win_u, out);
                }
                stamp(15, out[0] + out[1]);
                const float wsum = warp_sum(own_products(a_own_r + U_MINUS_R, out));
                if (lane == 0) sts_f32(a_wu + 4u * (warp - NW), wsum);
                stamp(4, wsum);
                __syncwarp();                                        // named barriers are warp-aligned: reconverge after lane-dependent code
                named_bar_arrive(1, NP + 32);
                if (send_l || send_r) {
                    const uint64_t ga = (send_l ? gaddr_l : gaddr_r) + par * HALO_PAR_BYTES;
#pragma unroll
                    for (uint32_t k = 0; k < RPT; ++k)
                        if (rowv[k]) {
                            const uint64_t pk = ((uint64_t)ep << 32) | (uint64_t)__float_as_uint(out[k]);
                            asm volatile("st.relaxed.cluster.u64 [%0], %1;" ::"l"(ga + EOFF[k]), "l"(pk) : "memory");
                        }
                }
                if (hl) {                                            // first touch of the slots now: a slot is seen sooner once polled
                    k0 = ld_packet_local(my_halo + par * HALO_PAR_BYTES);
                    if constexpr (!UX) k1 = ld_packet_local(my_halo + par * HALO_PAR_BYTES + 8u * XS);
                }
                stamp(5, out[0]);
            } else if (!xw) {
                // ---- P-warps: r.u of their rows while the S product runs, then asleep until the scalars are published
                const float rsum = warp_sum(own_products(a_own_r, out));
                if (lane == 0) sts_f32(a_ru + 4u * warp, rsum);
                __syncwarp();
                named_bar_arrive(3, NP);
            } else {
                // ---- exchange warp: scalars that only need the previous gamma and denominator, then r.u (both while the S product runs)
                float rgam = first ? 0.f : rcp_fast(gam), qq = __fmul_rn(den, rgam);            // qq = 1 / alpha
                asm volatile("" : "+f"(rgam), "+f"(qq));
                const float rsum = warp_sum(own_products(a_own_r, out));
                if (lane == 0) sts_f32(a_ru, rsum);
                __syncwarp();
                named_bar_sync(3, NP);                               // the other P-warps have parked their r.u sums
                const float cg = role_sum(a_ru);
                stamp(4, cg);
                __syncwarp();
                named_bar_sync(1, NP + 32);                          // the S-warps have parked their w.u sums
                stamp(5, cg);
                const float cd = role_sum(a_wu);
                if (lane < C)
                    asm volatile("st.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gpeer_dot + 16u * (par * C)), "r"(__float_as_uint(cg)), "r"(ep),
                                 "r"(__float_as_uint(cd)), "r"(ep)
                                 : "memory");
                stamp(6, cd);
                uint4 qd[C];
                bool ok;
                uint32_t spins = 0;
                do {
                    ok = true;
#pragma unroll
                    for (uint32_t m = 0; m < C; ++m) {
                        qd[m] = ld_pair(dot_u + 16u * (par * C + m));
                        ok = ok && qd[m].y == ep && qd[m].w == ep;
                    }
                    if (++spins > (1u << 24)) __trap();
                } while (!ok);
                stamp(7, __uint_as_float(qd[0].x));
                float vg[C], vd[C];
#pragma unroll
                for (uint32_t m = 0; m < C; ++m) { vg[m] = __uint_as_float(qd[m].x); vd[m] = __uint_as_float(qd[m].z); }
                const float gam_new = tree_sum<C>(vg), del_new = tree_sum<C>(vd);
                done = !first && fabsf(gam_new) < a.exit_tol;                                // pcg.cuh:195
                if (first) {
                    beta = 0.f;
                    den = del_new;
                } else {
                    beta = __fmul_rn(gam_new, rgam);
                    den = __fmaf_rn(-__fmul_rn(beta, gam_new), qq, del_new);
                }
                alpha = __fmul_rn(gam_new, rcp_fast(den));
                gam = gam_new;
                if (lane == 0) {
                    sts_f32(a_sc, alpha);
                    sts_f32(a_sc + 4u, beta);
                    sts_f32(a_sc + 8u, done ? 1.f : 0.f);
                }
                stamp(8, alpha);
                __syncwarp();
                named_bar_arrive(2, NT);
                stamp(9, alpha);
            }
            if (!xw) {
                // every warp but the exchange warp sleeps here until the scalars are published (ONE barrier site for both roles:
                // compute-sanitizer's synccheck treats syncs on one barrier from different instructions as divergence)
                __syncwarp();
                named_bar_sync(2, NT);
                stamp(9, out[0]);
                alpha = lds_f32(a_sc);
                beta = lds_f32(a_sc + 4u);
                done = lds_f32(a_sc + 8u) != 0.f;
                if (hl) {
                    uint32_t spins2 = 0;
                    while (!(packet_ok(k0, ep) && (UX || packet_ok(k1, ep)))) {
                        k0 = ld_packet_local(my_halo + par * HALO_PAR_BYTES);
                        if constexpr (!UX) k1 = ld_packet_local(my_halo + par * HALO_PAR_BYTES + 8u * XS);
                        if (++spins2 > (1u << 24)) __trap();          // a lost packet is an error (launch failure), not a hang
                    }
                    hw = packet_val(k0);
                    if constexpr (!UX) hw2 = packet_val(k1);
                }
                stamp(10, hw);
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
            if constexpr (PROF) prof_now = a.dbg != nullptr && iter == PROF_ITER;
            stamp(0, alpha);
            // ---- p = u + beta p ; lambda += alpha p   (P-threads)      s = w + beta s ; r -= alpha s   (S-threads, halo lanes)
            const float sa = isP ? alpha : -alpha;
#pragma unroll
            for (uint32_t k = 0; k < RPT; ++k) {
                v1[k] = __fmaf_rn(beta, v1[k], out[k]);
                v2[k] = __fmaf_rn(sa, v1[k], v2[k]);
            }
            hs = __fmaf_rn(beta, hs, hw);
            hr = __fmaf_rn(-alpha, hs, hr);
            if constexpr (!UX) {
                hs2 = __fmaf_rn(beta, hs2, hw2);
                hr2 = __fmaf_rn(-alpha, hs2, hr2);
            }
            step();
            stamp(11, alpha);
            if (done) { ++iter; max_iter_exit = 0; break; }
        }
        if constexpr (PROF) {
            if (a.dbg) {
#pragma unroll
                for (uint32_t i = 0; i < NSTAMP; ++i) a.dbg[i * (C * NT) + cr * NT + t] = tk[i];
            }
        }

        // ---- outputs                                                        (pcg.cuh:212-215)
#pragma unroll
        for (uint32_t k = 0; k < RPT; ++k)
            if (rowv[k]) {
                const size_t o = vbase + (size_t)b * n + ek[k];
                if (isP) {
                    a.lambda[o] = v2[k];
                    if (a.p_out) a.p_out[o] = v1[k];
                } else if (a.r_out) {
                    a.r_out[o] = v2[k];
                }
            }
        if (cr == 0 && t == 0) store_result(a, sys, iter, max_iter_exit);
        __syncthreads();
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
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false, bool UX = false>
__global__ void __launch_bounds__(ClusterPcgFastB<n, N, C, UX>::NT, MINB)
pcg_cluster_kernel_fastb(const PcgArgs<float> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_fastb_init<n, N, C, UX>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, packet buffers cleared, before any DSMEM traffic
    pcg_cluster_fastb_run<n, N, C, PROF, UX>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
