// gbd_bcr.cuh -- direct solve of the block-tridiagonal Schur system S lambda = gamma by block cyclic reduction inside
// one thread-block cluster (the "GPU direct alternative" to the reference's CPU QDLDL path,
// include/qdldl/sqp.cuh:22-49, for systems on which PCG runs into its iteration cap).
//
// NOT the reference's algorithm: the reference solves this system either iteratively (pcg<>) or with QDLDL's sequential
// sparse LDL^T on the CPU.  Results therefore agree with those to a tolerance, not bit for bit (tests state it); what is
// kept is the interface -- the same band layout S = [N][left | diag | right][n][n] column-major, gamma, lambda.
//
// Algorithm (fp32, no pivoting; S is symmetric definite).  Level l = 0 .. log2(N)-1 with stride s = 2^l:
//   phase 1  every row j = s (mod 2s) is eliminated: one warp inverts D_j (Gauss-Jordan with the rows in registers,
//            gbd_schur.cuh) and publishes W_j = D_j^-1 [L_j | U_j | b_j]  (n x (2n+1))
//   phase 2  every row i = 0 (mod 2s) absorbs its two eliminated neighbours j = i -+ s:
//              D_i -= L_i W_U(i-s) + U_i W_L(i+s);  L_i <- -L_i W_L(i-s);  U_i <- -U_i W_U(i+s);  b_i -= L_i w_b(i-s) + U_i w_b(i+s)
// then x_0 = D_0^-1 b_0 and back substitution x_j = w_b(j) - W_L(j) x_{j-s} - W_U(j) x_{j+s}, levels descending.
// Mapping: C = N/R CTAs of 4 warps, R consecutive rows per CTA, all row data in shared memory for the whole solve; one
// warp per active row, lane r owns matrix row r; neighbours in other CTAs are read through DSMEM (W_j is first copied
// into a per-warp buffer with 128-bit loads); one cluster barrier between phase 1 and phase 2 and per back-substitution level.
// Dependent chain: log2(N) x (14-pivot inversion + two n x n x (2n+1) products) -- about 20 us at n = 14, N = 128.
#pragma once
#include <type_traits>
#include "gbd_device.cuh"
#include "gbd_schur.cuh"

namespace gbd {

template <uint32_t n, uint32_t N, uint32_t C>
struct BcrShape {
    static_assert((N & (N - 1)) == 0 && N >= 2, "cyclic reduction is written for a power-of-two number of block rows");
    static_assert(N % C == 0 && C >= 1 && C <= 16, "cluster shape");
    static_assert(n + 1 <= 16, "one warp per row: lane r owns matrix row r");
    static constexpr uint32_t R = N / C;                 // rows per CTA
    static constexpr uint32_t NT = 128, WARPS = 4;
    static constexpr uint32_t nn = n * n;
    static constexpr uint32_t WC = 2 * n + 1;            // columns of W = [W_L | W_U | w_b]
    static constexpr uint32_t pad4(uint32_t x) { return (x + 3) / 4 * 4; }
    // per-row record (floats): D, L, U (n x n column-major), b, W (n x WC column-major), x
    static constexpr uint32_t OFF_D = 0, OFF_L = nn, OFF_U = 2 * nn, OFF_B = 3 * nn, OFF_W = pad4(3 * nn + n);
    static constexpr uint32_t WF = pad4(n * WC);
    static constexpr uint32_t OFF_X = OFF_W + WF;
    static constexpr uint32_t ROWF = OFF_X + pad4(n);
    static constexpr uint32_t SCRATCH = 2 * WF + 32;     // per warp: two neighbour W copies + Gauss-Jordan snapshot
    static constexpr size_t SMEM_BYTES = sizeof(float) * ((size_t)R * ROWF + WARPS * SCRATCH);
};

__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cluster_f1(uint32_t cluster_addr)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}

struct BcrArgs {
    const float *S;       // [batch][N][3][n][n]
    const float *gamma;   // [batch][N*n]
    float *lambda;        // [batch][N*n]  out
    uint32_t batch;
    const uint8_t *only_if;   // nullable [batch]: solve system i only where only_if[i] != 0 (PCG's max_iter_exit flags)
    uint32_t *dbg;        // timeline build only: %clock stamps [stamp][CTA][warp] (gbd_pcg_set_debug_buffer, tools/timeline_bcr.py)
};

// The body is a __device__ function so that both the C-ABI kernel and the drop-in pcg<T,n,N> template (-DGBD_DROPIN_DIRECT=1) can wrap
// it.  Called by exactly NT = 128 threads of every CTA of a C-CTA cluster; solves systems first_sys, first_sys + sys_stride, ...
template <uint32_t n, uint32_t N, uint32_t C, bool PROF = false>
__device__ __forceinline__ void bcr_cluster_body(const BcrArgs &a, float *bsm, uint32_t first_sys, uint32_t sys_stride)
{
    using K = BcrShape<n, N, C>;
    constexpr uint32_t R = K::R, nn = K::nn, WC = K::WC, ROWF = K::ROWF, WF = K::WF, NT = K::NT;
    float *rows = bsm;                                   // R row records
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    float *scratch = bsm + (size_t)R * ROWF + warp * K::SCRATCH;
    float *wm = scratch, *wp = scratch + WF, *snap = scratch + 2 * WF;      // W of row i-s, W of row i+s, GJ snapshot
    const uint32_t cr = cluster_ctarank();
    const uint32_t r = lane < n ? lane : 0;              // lanes n.. shadow row 0 and never store
    const bool act = lane < n;
    const uint32_t rows_u = smem_u32(rows);

    uint32_t nstamp = 0;
    auto stamp = [&]() {                                 // timeline build: one %clock per warp per call site, in program order
        if constexpr (PROF) {
            if (a.dbg && lane == 0) {
                uint32_t c_;
                asm volatile("mov.u32 %0, %%clock;" : "=r"(c_)::"memory");
                a.dbg[(nstamp * C + cr) * K::WARPS + warp] = c_;
            }
            ++nstamp;
        }
    };
    auto rec = [&](uint32_t grow) -> float * { return rows + (size_t)(grow - cr * R) * ROWF; };          // local rows only
    auto rec_cluster = [&](uint32_t grow, uint32_t off) -> uint32_t {                                     // any row of the system
        return map_to_cta(rows_u + 4u * ((grow % R) * ROWF + off), grow / R);
    };
    // copy W of row j (any CTA) into a per-warp buffer: 128-bit DSMEM loads, all lanes
    auto fetch_w = [&](uint32_t j, float *dst) {
        constexpr uint32_t Q = (WF / 4 + 31) / 32;          // all of a lane's loads are issued before the first store
        float4 v[Q];
        if (j / R == cr) {                                   // neighbour in this CTA: plain shared-memory loads
            const float4 *src = reinterpret_cast<const float4 *>(rec(j) + K::OFF_W);
#pragma unroll
            for (uint32_t q = 0; q < Q; ++q)
                if (lane + 32 * q < WF / 4) v[q] = src[lane + 32 * q];
        } else {
            const uint32_t src = rec_cluster(j, K::OFF_W);
#pragma unroll
            for (uint32_t q = 0; q < Q; ++q)
                if (lane + 32 * q < WF / 4) v[q] = ld_cluster_f4(src + 16u * (lane + 32 * q));
        }
#pragma unroll
        for (uint32_t q = 0; q < Q; ++q)
            if (lane + 32 * q < WF / 4) reinterpret_cast<float4 *>(dst)[lane + 32 * q] = v[q];
    };
    // tasks of this warp at stride s: rows first, first + 2s, ... inside this CTA (or the single row cr*R when 2s > R)
    auto for_rows = [&](uint32_t s, uint32_t residue, auto &&body) {
        if (2 * s <= R) {
            for (uint32_t tk = warp; tk < R / (2 * s); tk += K::WARPS) body(cr * R + residue + 2 * s * tk);
        } else if (warp == 0 && (cr * R) % (2 * s) == residue) {
            body(cr * R);
        }
    };

    for (uint32_t sys = first_sys; sys < a.batch; sys += sys_stride) {
        if (a.only_if && a.only_if[sys] == 0) continue;   // cluster-uniform: every CTA reads the same byte
        const float *gS = a.S + ((size_t)sys * N + (size_t)cr * R) * 3 * nn;
        const float *gb = a.gamma + (size_t)sys * N * n + (size_t)cr * R * n;
        // ---- load: tiles [left | diag | right] -> L, D, U ; the two tiles the format never defines are zero
        for (uint32_t i = t; i < R * 3 * nn; i += NT) {
            const uint32_t lr = i / (3 * nn), e = i % (3 * nn), tile = e / nn, q = e % nn;
            const uint32_t grow = cr * R + lr;
            float v = gS[i];
            if ((grow == 0 && tile == 0) || (grow == N - 1 && tile == 2)) v = 0.0f;
            rows[(size_t)lr * ROWF + (tile == 0 ? K::OFF_L : (tile == 1 ? K::OFF_D : K::OFF_U)) + q] = v;
        }
        for (uint32_t i = t; i < R * n; i += NT) rows[(size_t)(i / n) * ROWF + K::OFF_B + i % n] = gb[i];
        __syncthreads();
        stamp();

        // ---- forward reduction
        for (uint32_t s = 1; s < N; s <<= 1) {
            // From the level on where a CTA holds at most one active row (2s >= R) its four warps share that row's work:
            // warp 0 inverts, then each warp forms a quarter of the columns of W; in phase 2 each warp takes one side
            // (i-s / i+s) and one half of the columns.  Below that level there is one warp per row (path further down).
            const bool coop = 2 * s >= R;
            if (coop) {
                const bool has1 = (2 * s == R) || ((cr * R) % (2 * s) == s);
                const uint32_t j = (2 * s == R) ? cr * R + s : cr * R;
                float *rj = rec(has1 ? j : cr * R);
                float *invbuf = bsm + (size_t)R * ROWF + WF;          // warp 0's second scratch buffer, read by all four warps
                if (has1 && warp == 0) {
                    float m[2 * n];
#pragma unroll
                    for (uint32_t c = 0; c < n; ++c) {
                        m[c] = act ? rj[K::OFF_D + r + c * n] : 0.0f;
                        m[n + c] = (c == r) ? 1.0f : 0.0f;
                    }
                    schur_detail::gj_regs<n, false, true>(m, snap, lane);
                    if (act) {
#pragma unroll
                        for (uint32_t c = 0; c < n; ++c) invbuf[r + c * n] = m[n + c];
                    }
                }
                __syncthreads();
                if (has1) {
                    float iv[n];
#pragma unroll
                    for (uint32_t k = 0; k < n; ++k) iv[k] = act ? invbuf[r + k * n] : 0.0f;
                    const float *X = rj + K::OFF_L;
                    constexpr uint32_t CW = (WC + 3) / 4;
                    auto wq = [&](auto c0_tag) {
                        constexpr uint32_t c0 = decltype(c0_tag)::value;
                        constexpr uint32_t CNT = c0 >= WC ? 0 : (c0 + CW <= WC ? CW : WC - c0);
                        if constexpr (CNT > 0) {
                            float acc[CNT];
#pragma unroll
                            for (uint32_t c = 0; c < CNT; ++c) acc[c] = 0.0f;
#pragma unroll
                            for (uint32_t k = 0; k < n; k += 2) {
#pragma unroll
                                for (uint32_t c = 0; c < CNT; ++c) {
                                    const float2 x2 = *reinterpret_cast<const float2 *>(X + k + (c0 + c) * n);
                                    acc[c] = fma_rn(iv[k], x2.x, acc[c]);
                                    acc[c] = fma_rn(iv[k + 1], x2.y, acc[c]);
                                }
                            }
                            if (act) {
#pragma unroll
                                for (uint32_t c = 0; c < CNT; ++c) rj[K::OFF_W + r + (c0 + c) * n] = acc[c];
                            }
                        }
                    };
                    if (warp == 0) wq(std::integral_constant<uint32_t, 0>{});
                    else if (warp == 1) wq(std::integral_constant<uint32_t, CW>{});
                    else if (warp == 2) wq(std::integral_constant<uint32_t, 2 * CW>{});
                    else wq(std::integral_constant<uint32_t, 3 * CW>{});
                }
            } else
            // phase 1: eliminate rows j = s (mod 2s)
            for_rows(s, s % (2 * s), [&](uint32_t j) {
                float *rj = rec(j);
                float m[2 * n];
#pragma unroll
                for (uint32_t c = 0; c < n; ++c) {
                    m[c] = act ? rj[K::OFF_D + r + c * n] : 0.0f;
                    m[n + c] = (c == r) ? 1.0f : 0.0f;
                }
                schur_detail::gj_regs<n, false, true>(m, snap, lane);
                // W = D^-1 [L | U | b]: lane r forms row r; [L | U | b] is contiguous in the record.  k outermost: the WC
                // accumulators of a lane are independent chains (ILP), operands come in as 64-bit broadcast loads
                const float *X = rj + K::OFF_L;
                static_assert(n % 2 == 0, "64-bit operand loads");
                constexpr uint32_t CH = (WC + 1) / 2;              // two passes over the columns keep the accumulators in registers
                auto w_pass = [&](auto c0_tag) {                   // compile-time column base: no index arithmetic in the loop
                    constexpr uint32_t c0 = decltype(c0_tag)::value;
                    constexpr uint32_t CNT = c0 + CH <= WC ? CH : WC - c0;
                    float acc[CNT];
#pragma unroll
                    for (uint32_t c = 0; c < CNT; ++c) acc[c] = 0.0f;
#pragma unroll
                    for (uint32_t k = 0; k < n; k += 2) {
#pragma unroll
                        for (uint32_t c = 0; c < CNT; ++c) {
                            const float2 x2 = *reinterpret_cast<const float2 *>(X + k + (c0 + c) * n);
                            acc[c] = fma_rn(m[n + k], x2.x, acc[c]);
                            acc[c] = fma_rn(m[n + k + 1], x2.y, acc[c]);
                        }
                    }
                    if (act) {
#pragma unroll
                        for (uint32_t c = 0; c < CNT; ++c) rj[K::OFF_W + r + (c0 + c) * n] = acc[c];
                    }
                };
                w_pass(std::integral_constant<uint32_t, 0>{});
                w_pass(std::integral_constant<uint32_t, CH>{});
            });
            stamp();
            cluster_sync();                                // W of every eliminated row visible cluster-wide
            stamp();
            if (coop) {
                const bool has2 = (2 * s == R) || ((cr * R) % (2 * s) == 0);
                const uint32_t i = cr * R;
                float *ri = rec(i);
                const bool plus = warp >= 2;
                const uint32_t half = warp & 1u;
                const bool has_n = has2 && (plus ? i + s < N : i >= s);
                constexpr uint32_t CH = n / 2;
                float dacc[CH], nacc[CH], bacc = 0.0f;
#pragma unroll
                for (uint32_t c = 0; c < CH; ++c) { dacc[c] = 0.0f; nacc[c] = 0.0f; }
                const uint32_t off_c = plus ? K::OFF_U : K::OFF_L;
                if (has_n) {
                    fetch_w(plus ? i + s : i - s, wm);
                    __syncwarp();
                    float crow[n];
#pragma unroll
                    for (uint32_t k = 0; k < n; ++k) crow[k] = act ? ri[off_c + r + k * n] : 0.0f;      // shadow lanes read nothing another lane writes
                    auto qpass = [&](auto c0_tag, auto side_tag) {
                        constexpr uint32_t c0 = decltype(c0_tag)::value;
                        constexpr bool minus_side = decltype(side_tag)::value;
#pragma unroll
                        for (uint32_t k = 0; k < n; k += 2) {
#pragma unroll
                            for (uint32_t c = 0; c < CH; ++c) {
                                const float2 wl = *reinterpret_cast<const float2 *>(wm + k + (c0 + c) * n);
                                const float2 wu = *reinterpret_cast<const float2 *>(wm + k + (n + c0 + c) * n);
                                const float2 wd = minus_side ? wu : wl, wn = minus_side ? wl : wu;
                                dacc[c] = fma_rn(crow[k], wd.x, dacc[c]); dacc[c] = fma_rn(crow[k + 1], wd.y, dacc[c]);
                                nacc[c] = fma_rn(crow[k], wn.x, nacc[c]); nacc[c] = fma_rn(crow[k + 1], wn.y, nacc[c]);
                            }
                        }
                    };
                    if (!plus) { if (half == 0) qpass(std::integral_constant<uint32_t, 0>{}, std::true_type{}); else qpass(std::integral_constant<uint32_t, CH>{}, std::true_type{}); }
                    else       { if (half == 0) qpass(std::integral_constant<uint32_t, 0>{}, std::false_type{}); else qpass(std::integral_constant<uint32_t, CH>{}, std::false_type{}); }
                    if (half == 0) {
#pragma unroll
                        for (uint32_t k = 0; k < n; k += 2) {
                            const float2 wb = *reinterpret_cast<const float2 *>(wm + k + 2 * n * n);
                            bacc = fma_rn(crow[k], wb.x, bacc); bacc = fma_rn(crow[k + 1], wb.y, bacc);
                        }
                    }
                }
                __syncthreads();                           // every warp has read its coupling row before anyone rewrites L / U
                const uint32_t c0r = half * CH;
                if (has2 && act) {
#pragma unroll
                    for (uint32_t c = 0; c < CH; ++c) ri[off_c + r + (c0r + c) * n] = has_n ? -nacc[c] : 0.0f;
                    if (!plus && has_n) {
#pragma unroll
                        for (uint32_t c = 0; c < CH; ++c) ri[K::OFF_D + r + (c0r + c) * n] -= dacc[c];
                        if (half == 0) ri[K::OFF_B + r] -= bacc;
                    }
                }
                __syncthreads();                           // minus side done with D and b; now the plus side
                if (has2 && act && plus && has_n) {
#pragma unroll
                    for (uint32_t c = 0; c < CH; ++c) ri[K::OFF_D + r + (c0r + c) * n] -= dacc[c];
                    if (half == 0) ri[K::OFF_B + r] -= bacc;
                }
            } else
            // phase 2: rows i = 0 (mod 2s) absorb their eliminated neighbours
            for_rows(s, 0, [&](uint32_t i) {
                float *ri = rec(i);
                const bool has_m = i >= s, has_p = i + s < N;
                if (has_m) fetch_w(i - s, wm);
                if (has_p) fetch_w(i + s, wp);
                __syncwarp();
                // one neighbour at a time (register budget): lane r holds row r of the coupling block, 2 n + 1 accumulators in
                // flight, k outermost, 64-bit broadcast loads of the neighbour's W copy
                float bl = 0.0f, bu = 0.0f;
                auto absorb = [&](const float *w, uint32_t off_c, auto side_tag, float &bacc) {
                    constexpr bool minus_side = decltype(side_tag)::value;
                    float crow[n];
#pragma unroll
                    for (uint32_t k = 0; k < n; ++k) crow[k] = act ? ri[off_c + r + k * n] : 0.0f;      // shadow lanes read nothing another lane writes
                    constexpr uint32_t CH = n / 2;                 // two passes over the columns keep the accumulators in registers
                    auto a_pass = [&](auto c0_tag) {
                        constexpr uint32_t c0 = decltype(c0_tag)::value;
                        float dacc[CH], nacc[CH];
#pragma unroll
                        for (uint32_t c = 0; c < CH; ++c) { dacc[c] = 0.0f; nacc[c] = 0.0f; }
#pragma unroll
                        for (uint32_t k = 0; k < n; k += 2) {
#pragma unroll
                            for (uint32_t c = 0; c < CH; ++c) {
                                const float2 wl = *reinterpret_cast<const float2 *>(w + k + (c0 + c) * n);           // W_L of the neighbour
                                const float2 wu = *reinterpret_cast<const float2 *>(w + k + (n + c0 + c) * n);       // W_U of the neighbour
                                // minus side (j = i-s): D -= L_i W_U, L <- -L_i W_L ; plus side (j = i+s): D -= U_i W_L, U <- -U_i W_U
                                const float2 wd = minus_side ? wu : wl, wn = minus_side ? wl : wu;
                                dacc[c] = fma_rn(crow[k], wd.x, dacc[c]); dacc[c] = fma_rn(crow[k + 1], wd.y, dacc[c]);
                                nacc[c] = fma_rn(crow[k], wn.x, nacc[c]); nacc[c] = fma_rn(crow[k + 1], wn.y, nacc[c]);
                            }
                        }
                        if (act) {
#pragma unroll
                            for (uint32_t c = 0; c < CH; ++c) {
                                ri[K::OFF_D + r + (c0 + c) * n] -= dacc[c];
                                ri[off_c + r + (c0 + c) * n] = -nacc[c];
                            }
                        }
                    };
                    a_pass(std::integral_constant<uint32_t, 0>{});
                    a_pass(std::integral_constant<uint32_t, CH>{});
#pragma unroll
                    for (uint32_t k = 0; k < n; k += 2) {
                        const float2 wb = *reinterpret_cast<const float2 *>(w + k + 2 * n * n);
                        bacc = fma_rn(crow[k], wb.x, bacc); bacc = fma_rn(crow[k + 1], wb.y, bacc);
                    }
                };
                if (has_m) absorb(wm, K::OFF_L, std::true_type{}, bl);
                else if (act) {
#pragma unroll
                    for (uint32_t c = 0; c < n; ++c) ri[K::OFF_L + r + c * n] = 0.0f;
                }
                if (has_p) absorb(wp, K::OFF_U, std::false_type{}, bu);
                else if (act) {
#pragma unroll
                    for (uint32_t c = 0; c < n; ++c) ri[K::OFF_U + r + c * n] = 0.0f;
                }
                if (act) ri[K::OFF_B + r] -= bl + bu;
                __syncwarp();
            });
            stamp();
            __syncthreads();                               // the updated rows are read next by warps of this CTA only
            stamp();
        }
        // ---- root: x_0 = D_0^-1 b_0 (row 0 lives in CTA 0)
        if (cr == 0 && warp == 0) {
            float *r0 = rec(0);
            float m[2 * n];
#pragma unroll
            for (uint32_t c = 0; c < n; ++c) {
                m[c] = act ? r0[K::OFF_D + r + c * n] : 0.0f;
                m[n + c] = (c == r) ? 1.0f : 0.0f;
            }
            schur_detail::gj_regs<n, false, true>(m, snap, lane);
            float acc = 0.0f;
#pragma unroll
            for (uint32_t k = 0; k < n; ++k) acc = fma_rn(m[n + k], r0[K::OFF_B + k], acc);
            if (act) r0[K::OFF_X + r] = acc;
        }
        stamp();
        cluster_sync();
        stamp();
        // ---- back substitution, strides descending
        for (uint32_t s = N / 2; s >= 1; s >>= 1) {
            for_rows(s, s % (2 * s), [&](uint32_t j) {
                float *rj = rec(j);
                const bool has_p = j + s < N;                  // j - s >= 0 always
                const uint32_t xm = rec_cluster(j - s, K::OFF_X), xp = rec_cluster(has_p ? j + s : j, K::OFF_X);
                float acc = rj[K::OFF_W + r + 2 * n * n];       // w_b
#pragma unroll
                for (uint32_t k = 0; k < n; ++k) {
                    acc = fma_rn(-rj[K::OFF_W + r + k * n], ld_cluster_f1(xm + 4u * k), acc);
                    if (has_p) acc = fma_rn(-rj[K::OFF_W + r + (n + k) * n], ld_cluster_f1(xp + 4u * k), acc);
                }
                if (act) rj[K::OFF_X + r] = acc;
            });
            // rows at distance < R from their neighbours only read x of this CTA or of rows solved before the last cluster barrier
            if (s >= R) cluster_sync(); else __syncthreads();
        }
        // ---- output
        float *gl = a.lambda + (size_t)sys * N * n + (size_t)cr * R * n;
        stamp();
        for (uint32_t i = t; i < R * n; i += NT) gl[i] = rows[(size_t)(i / n) * ROWF + K::OFF_X + i % n];
        cluster_sync();                                    // nobody reloads rows while a peer may still read x / W
    }
}

template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
__global__ void __launch_bounds__(BcrShape<n, N, C>::NT, MINB)
bcr_cluster_kernel(const BcrArgs a)
{
    extern __shared__ __align__(16) float bsm[];
    bcr_cluster_body<n, N, C, PROF>(a, bsm, cluster_idx(), cluster_count());
}

}  // namespace gbd
