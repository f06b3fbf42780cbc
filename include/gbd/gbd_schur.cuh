// gbd_schur.cuh -- the two steps either side of the PCG solve inside one SQP iteration:
//
//   form_schur   replaces form_schur_system<T>  (include/pcg/linsys_setup.cuh:621-657: cooperative kernel
//                form_S_gamma_Pinv_kernel = form_S_gamma_and_jacobi_Pinv_blockrow :139-562, grid.sync,
//                complete_SS_Pinv_blockrow :9-137): KKT blocks (G, C, g, c) -> S, Pinv, gamma in the pcg<> layout,
//                G overwritten with the block inverses that compute_dz needs
//   compute_dz   replaces compute_dz<T>          (include/common/dz.cuh:3-136)
//
// Same inputs, outputs, layouts and the same floating-point operation ORDER as the reference (every dot product is
// one FMA per term in ascending index order, the Gauss-Jordan updates are x/piv and fma(-(c/piv), r, x) resp.
// x*pvInv and fma(-(c*pvInv), r, x), IEEE division), so the results are bit-identical to the reference kernels
// (asserted against oracle/_ref/libref_schur.so and the C oracle).  What differs is the machine mapping:
//
//   reference                                       here
//   ---------                                       ----
//   one cooperative launch, grid.sync between the   two ordinary launches (stream order is the dependency); no
//   two phases (all N CTAs must be co-resident)     co-residency requirement, so any N and batches work
//   ~60 __syncthreads per block row: every step of  5 CTA barriers per block row: the three Gauss-Jordan inversions
//   every 14-pivot inversion is a CTA barrier pair  run concurrently in three warps on __syncwarp; independent
//                                                   products (A Q^-1, B R^-1, Q^-1 q ...) share a stage
//   per-element division in the pivot update        one division per row and pivot (same operands -> same bits)
//   one mapping                                     two mappings of the same statements: a CTA per block row (one
//                                                   trajectory: lowest latency) and a WARP per block row with both Q
//                                                   blocks inverted in one pass (batches: no CTA barrier, every resident
//                                                   warp always has work)
//   matrices updated in shared memory               Gauss-Jordan on a sliding register window (n+1 registers per row,
//                                                   the pivot loop a real loop), products register-blocked 2 x 2
//   CTA k overwrites G slot k-1 with the inverse    inverses are parked in Pinv tiles that phase 2 overwrites
//   while CTA k-1 may still read it (a race that    anyway (left tile of row k, right tile of row k-1, the pad
//   only co-residency hides)                        tile of row 0) and moved into G by the phase-2 CTA that owns
//                                                   the tile: in place like the reference, race-free
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbd {

namespace schur_detail {

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// a / b, IEEE round-to-nearest, bit for bit __fdiv_rn(a, b).  __fdiv_rn leaves its inline sequence (MUFU.RCP + 5 FMAs) for a
// ~35-instruction subroutine whenever FCHK flags an operand, and an exactly zero numerator is such an operand.  In the
// Gauss-Jordan steps zero numerators are the rule (the identity half of [V | I], the off-diagonal zeros of a diagonal Q), and one
// flagged lane sends the whole warp through the subroutine: measured 19 calls per block row, 12 % of the batched kernel's
// instructions, on the dependent pivot chain.  0 / b for finite non-zero b is the zero whose sign is sign(a) xor sign(b): formed
// directly, and the division itself runs on a numerator that does not raise the flag.
__device__ __forceinline__ float div_rn(float a, float b)
{
    const bool z = a == 0.0f && b != 0.0f && fabsf(b) < __int_as_float(0x7f800000);
    const float q = __fdiv_rn(z ? 1.0f : a, b);
    return z ? __int_as_float((__float_as_int(a) ^ __float_as_int(b)) & (int)0x80000000) : q;
}

// (gj_regs: the form the direct solver gbd_bcr.cuh calls with its rows already in registers; the assembly kernels below use the
// sliding-window form gj_div_window / gj_rcp_warp, which performs the same updates.)
// Gauss-Jordan on [V | I] (DIM x 2 DIM) by ONE warp with the matrix rows in REGISTERS: lane r < DIM holds row r of the
// augmented matrix (2 DIM values, the identity half generated in place), the pivot loop is fully unrolled so every
// register index is static.  Per pivot p the reference updates the DIM+1 columns p .. p+DIM from a snapshot of the
// pivot column `col` and pivot row `rowv`:
//   several-matrices form (matrix.cuh:149-238, DIV = true):  row == p : x / piv      else : fma(-(col[row] / piv), rowv[c], x)
//   single-matrix form    (matrix.cuh:120-146, DIV = false): pvInv = 1 / piv;
//                                                            row == p : x * pvInv    else : fma(-(col[row] * pvInv), rowv[c], x)
// Here: lane p publishes its window row through shared memory; ONE division sequence serves both kinds of quotient
// (lanes 0..DIM-1 form their own row factor col[row]/piv from a register, lanes 16..16+DIM the new pivot-row element
// rowv[c]/piv); then one FMA per element.  Same operands and operations per element as the reference -> same bits,
// with ~2x fewer instructions and half the dependent latency of the shared-memory-resident version.
// A: in shared memory, V (DIM x DIM, column-major) followed by DIM x DIM floats that receive V^-1.  snap: 32 floats, 16-B aligned.
// core: rows in registers in, rows of [.. | V^-1] in registers out (a[DIM .. 2 DIM) of lane r = row r of the inverse)
// FAST_RCP (direct solver only, never on a bit-exact path): pvInv from the hardware reciprocal (1 ulp) instead of an IEEE
// division -- the reciprocal sits on the 14-step dependent pivot chain.
template <uint32_t DIM, bool DIV, bool FAST_RCP = false>
__device__ __forceinline__ void gj_regs(float (&a)[2 * DIM], float *snap, uint32_t lane)
{
    static_assert(DIM + 1 <= 16, "one warp: row factors in lanes 0..15, new pivot row in lanes 16..31");
    static_assert(!FAST_RCP || !DIV, "the fast reciprocal only exists for the single-matrix form");
    if constexpr (!DIV) {
        // single-matrix form: every quotient is a product with pvInv = 1 / piv, so no lane needs another lane's division:
        // the pivot row travels by shuffles and every lane forms the new pivot row itself -- no shared-memory round trip
        (void)snap;
#pragma unroll
        for (uint32_t p = 0; p < DIM; ++p) {
            const float piv = __shfl_sync(0xffffffffu, a[p], p);
            float pv_inv;
            if constexpr (FAST_RCP) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pv_inv) : "f"(piv));
            else pv_inv = __fdiv_rn(1.0f, piv);
            const float f = __fmul_rn(a[p], pv_inv);
#pragma unroll
            for (uint32_t c = 0; c <= DIM; ++c) {
                const float rv = __shfl_sync(0xffffffffu, a[p + c], p);
                a[p + c] = (lane == p) ? __fmul_rn(rv, pv_inv) : fma_(-f, rv, a[p + c]);
            }
        }
        return;
    }
    float *row = snap, *nrow = snap + 16;
    const uint32_t cI = lane >= 16 ? (lane - 16 <= DIM ? lane - 16 : 0) : 0;
#pragma unroll
    for (uint32_t p = 0; p < DIM; ++p) {
        if (lane == p) {
#pragma unroll
            for (uint32_t c = 0; c <= DIM; ++c) row[c] = a[p + c];
        }
        __syncwarp();
        const float piv = row[0];
        const float num = lane < 16 ? a[p] : row[cI];
        const float q = __fdiv_rn(num, piv);
        if (lane >= 16 && lane - 16 <= DIM) nrow[cI] = q;
        __syncwarp();
        // lane p takes the new pivot row, every other lane the old one: one address select, then 128-bit loads (the
        // snapshots are 16 floats each; entries past DIM are never used)
        constexpr uint32_t NV4 = (DIM + 4) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(lane == p ? nrow : row);
#pragma unroll
        for (uint32_t i = 0; i < NV4; ++i) {
            const float4 f = src[i];
            const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                const uint32_t c = 4 * i + e;
                if (c <= DIM) a[p + c] = (lane == p) ? v[e] : fma_(-q, v[e], a[p + c]);
            }
        }
        __syncwarp();
    }
}
// ---- the inversions of the assembly kernels: the same updates as gj_regs, held as a SLIDING WINDOW.  At pivot p the reference
// touches columns p .. p+DIM of [V | I]; column p is a unit vector afterwards and is never read again, and column p+DIM enters
// the range as the untouched identity column e_p.  So a lane keeps w[c] = column p + c of its row, c = 0 .. DIM: the update of
// column p + c lands in w[c - 1], the entering identity element in w[DIM], and every register index is the same at every pivot --
// the pivot loop is a real loop (one copy of ~70 instructions instead of 14: the unrolled form streamed ~70 KB of straight-line
// code through the instruction cache once per block row, 22 % of the batched kernel's stall cycles) on DIM + 1 registers
// instead of 2 DIM.  Operands and operations per element are unchanged -> same bits.
// TEAM lanes form a team around one matrix: 32 (lanes 0..15 rows and row factors, lanes 16..31 the quotients of the new pivot
// row) or 16 (two matrices per warp, a half-warp each; every lane forms both kinds of quotient).
// A: V (DIM x DIM, column-major) followed by DIM x DIM floats that receive V^-1;  snap: 32 floats per team, 16-byte aligned.
// Gsrc != nullptr: V is read straight from global memory with rho added on the diagonal (the reference's Q + rho I), so the
// block never passes through shared memory on its way into the registers; only V^-1 is written to A + DIM * DIM.
template <uint32_t DIM, uint32_t TEAM>
__device__ __forceinline__ void gj_div_window(float *A, float *snap, uint32_t lane, const float *Gsrc = nullptr, float rho = 0.0f)
{
    static_assert(DIM + 1 <= 16 && (TEAM == 16 || TEAM == 32), "DIM rows and DIM + 1 pivot-row quotients per 16 lanes");
    const uint32_t hl = lane & 15u, upper = (lane >> 4) & 1u;
    const uint32_t r = hl < DIM ? hl : 0;                        // lanes DIM..15 shadow row 0 (never stored)
    const uint32_t cI = hl <= DIM ? hl : 0;
    const bool rows = TEAM == 16 || !upper;                      // this lane holds a matrix row ...
    const bool quot = TEAM == 16 || upper;                       // ... and/or forms pivot-row quotients
    float *row = snap + (TEAM == 16 ? 32 * upper : 0u), *nrow = row + 16;
    const uint32_t me = rows ? hl : 32u;                         // my row index as a pivot (never matches on a quotient-only lane)
    float w[DIM + 1];
    if (Gsrc) {
#pragma unroll
        for (uint32_t c = 0; c < DIM; ++c) w[c] = Gsrc[c * DIM + r];
#pragma unroll
        for (uint32_t c = 0; c < DIM; ++c) w[c] = c == r ? __fadd_rn(w[c], rho) : w[c];
    } else {
#pragma unroll
        for (uint32_t c = 0; c < DIM; ++c) w[c] = A[c * DIM + r];
    }
    w[DIM] = r == 0 ? 1.0f : 0.0f;
#pragma unroll 1
    for (uint32_t p = 0; p < DIM; ++p) {
        if (me == p) {
#pragma unroll
            for (uint32_t c = 0; c <= DIM; ++c) row[c] = w[c];
        }
        __syncwarp();
        const float piv = row[0];
        float q, nq;
        if constexpr (TEAM == 16) {
            q = div_rn(w[0], piv);
            nq = div_rn(row[cI], piv);
        } else {
            q = nq = div_rn(upper ? row[cI] : w[0], piv);     // one division sequence serves both kinds of quotient
        }
        if (quot && hl <= DIM) nrow[cI] = nq;
        __syncwarp();
        // lane p takes the new pivot row, every other lane the old one: one address select, then 128-bit loads
        constexpr uint32_t NV4 = (DIM + 4) / 4;
        const float4 *src = reinterpret_cast<const float4 *>(me == p ? nrow : row);
        float v[4 * NV4];
#pragma unroll
        for (uint32_t i = 0; i < NV4; ++i) {
            const float4 f = src[i];
            v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
        }
#pragma unroll
        for (uint32_t c = 1; c <= DIM; ++c) w[c - 1] = (me == p) ? v[c] : fma_(-q, v[c], w[c]);
        w[DIM] = r == p + 1 ? 1.0f : 0.0f;
        __syncwarp();
    }
    if (rows && hl < DIM) {
#pragma unroll
        for (uint32_t c = 0; c < DIM; ++c) A[(DIM + c) * DIM + hl] = w[c];
    }
    __syncwarp();
}
template <uint32_t DIM>
__device__ __forceinline__ void gj_div_warp(float *A, float *snap, uint32_t lane, const float *Gsrc = nullptr, float rho = 0.0f)
{
    gj_div_window<DIM, 32>(A, snap, lane, Gsrc, rho);
}
// the several-matrices form on TWO matrices at once by one warp: lanes 0..15 on AX, lanes 16..31 on AY (snap: 64 floats)
template <uint32_t DIM>
__device__ __forceinline__ void gj_div_pair_warp(float *AX, float *AY, float *snap, uint32_t lane, const float *GX, const float *GY, float rho)
{
    gj_div_window<DIM, 16>(lane & 16u ? AY : AX, snap, lane, lane & 16u ? GY : GX, rho);
}
// single-matrix form (matrix.cuh:120-146): pvInv = 1 / piv; row == p : x * pvInv, else fma(-(col[row] * pvInv), rowv[c], x).
// No lane needs another lane's quotient, so the pivot row travels by shuffles.
template <uint32_t DIM>
__device__ __forceinline__ void gj_rcp_warp(float *A, float *snap, uint32_t lane)
{
    (void)snap;
    const uint32_t r = lane < DIM ? lane : 0;
    float w[DIM + 1];
#pragma unroll
    for (uint32_t c = 0; c < DIM; ++c) w[c] = A[c * DIM + r];
    w[DIM] = r == 0 ? 1.0f : 0.0f;
#pragma unroll 1
    for (uint32_t p = 0; p < DIM; ++p) {
        const float piv = __shfl_sync(0xffffffffu, w[0], p);
        const float pv_inv = __fdiv_rn(1.0f, piv);
        const float f = __fmul_rn(w[0], pv_inv);
#pragma unroll
        for (uint32_t c = 1; c <= DIM; ++c) {
            const float rv = __shfl_sync(0xffffffffu, w[c], p);
            w[c - 1] = (lane == p) ? __fmul_rn(rv, pv_inv) : fma_(-f, rv, w[c]);
        }
        w[DIM] = r == p + 1 ? 1.0f : 0.0f;
    }
    if (lane < DIM) {
#pragma unroll
        for (uint32_t c = 0; c < DIM; ++c) A[(DIM + c) * DIM + lane] = w[c];
    }
    __syncwarp();
}

// one element of C (M x NC) = A (M x K) * B (K x NC)  [TB: A * B^T with B stored NC x K], column-major,
// one FMA per term in ascending k (GLASS/src/L3/gemm.cuh:47-96)
template <bool TB>
__device__ __forceinline__ float gemm_elem(const float *A, const float *B, uint32_t M, uint32_t K, uint32_t NC, uint32_t row,
                                           uint32_t col)
{
    float res = 0.0f;
    for (uint32_t k = 0; k < K; ++k) res = fma_(A[k * M + row], TB ? B[k * NC + col] : B[col * K + k], res);
    return res;
}
// one element of mat (ROWS x COLS, column-major) * vec  (matrix.cuh:42-54)
__device__ __forceinline__ float matvec_elem(const float *mat, const float *vec, uint32_t ROWS, uint32_t COLS, uint32_t row)
{
    float res = 0.0f;
    for (uint32_t c = 0; c < COLS; ++c) res = fma_(mat[row + c * ROWS], vec[c], res);
    return res;
}
// Register-blocked forms of the two helpers above for even M: the rows (r, r+1), r even, of a column-major operand are one
// 64-bit shared-memory load.  Every output element is still one FMA per term in ascending k from 0.0f -> the same bits as
// gemm_elem / matvec_elem; what changes is the instruction count (the batched assembly is issue-bound): 2 x 2 outputs cost
// 2 loads + 4 FMAs per k instead of 8 loads + 4 FMAs.
// o = {(r, c), (r+1, c), (r, c+1), (r+1, c+1)} of A (M x K) * B (K x NC)  [TB: A * B^T, B stored NC x K]; r, c even
template <uint32_t M, uint32_t K, uint32_t NC, bool TB>
__device__ __forceinline__ void gemm_2x2(const float *A, const float *B, uint32_t r, uint32_t c, float (&o)[4])
{
    static_assert(M % 2 == 0 && NC % 2 == 0, "row pairs and column pairs");
    o[0] = o[1] = o[2] = o[3] = 0.0f;
    if constexpr (!TB && K % 2 == 0) {
#pragma unroll
        for (uint32_t k = 0; k < K; k += 2) {
            const float2 a0 = *reinterpret_cast<const float2 *>(A + k * M + r), a1 = *reinterpret_cast<const float2 *>(A + (k + 1) * M + r);
            const float2 b0 = *reinterpret_cast<const float2 *>(B + c * K + k), b1 = *reinterpret_cast<const float2 *>(B + (c + 1) * K + k);
            o[0] = fma_(a0.x, b0.x, o[0]); o[0] = fma_(a1.x, b0.y, o[0]);
            o[1] = fma_(a0.y, b0.x, o[1]); o[1] = fma_(a1.y, b0.y, o[1]);
            o[2] = fma_(a0.x, b1.x, o[2]); o[2] = fma_(a1.x, b1.y, o[2]);
            o[3] = fma_(a0.y, b1.x, o[3]); o[3] = fma_(a1.y, b1.y, o[3]);
        }
    } else {
#pragma unroll
        for (uint32_t k = 0; k < K; ++k) {
            const float2 a = *reinterpret_cast<const float2 *>(A + k * M + r);
            float b0, b1;
            if constexpr (TB) {
                const float2 bb = *reinterpret_cast<const float2 *>(B + k * NC + c);
                b0 = bb.x; b1 = bb.y;
            } else {
                b0 = B[c * K + k]; b1 = B[(c + 1) * K + k];
            }
            o[0] = fma_(a.x, b0, o[0]); o[1] = fma_(a.y, b0, o[1]);
            o[2] = fma_(a.x, b1, o[2]); o[3] = fma_(a.y, b1, o[3]);
        }
    }
}
// o = {(r, c), (r+1, c)} of A (M x K) * B (K x NC), r even
template <uint32_t M, uint32_t K>
__device__ __forceinline__ void gemm_2x1(const float *A, const float *B, uint32_t r, uint32_t c, float (&o)[2])
{
    static_assert(M % 2 == 0, "row pairs");
    o[0] = o[1] = 0.0f;
#pragma unroll
    for (uint32_t k = 0; k < K; ++k) {
        const float2 a = *reinterpret_cast<const float2 *>(A + k * M + r);
        const float b = B[c * K + k];
        o[0] = fma_(a.x, b, o[0]); o[1] = fma_(a.y, b, o[1]);
    }
}
// o = rows (r, r+1) of mat (ROWS x COLS, column-major) * vec, r even
template <uint32_t ROWS, uint32_t COLS>
__device__ __forceinline__ void matvec_2(const float *mat, const float *vec, uint32_t r, float (&o)[2])
{
    static_assert(ROWS % 2 == 0, "row pairs");
    o[0] = o[1] = 0.0f;
#pragma unroll
    for (uint32_t c = 0; c < COLS; ++c) {
        const float2 a = *reinterpret_cast<const float2 *>(mat + c * ROWS + r);
        const float x = vec[c];
        o[0] = fma_(a.x, x, o[0]); o[1] = fma_(a.y, x, o[1]);
    }
}
__device__ __forceinline__ void identity(float *A, uint32_t dim, uint32_t t, uint32_t nt)
{
    for (uint32_t i = t; i < dim * dim; i += nt) A[i] = (i % dim == i / dim) ? 1.0f : 0.0f;
}

}  // namespace schur_detail

template <uint32_t n, uint32_t m>
struct SchurShape {
    static constexpr uint32_t NT = 128;
    static constexpr uint32_t nn = n * n, mm = m * m, nm = n * m;
    static constexpr uint32_t GSET = nn + mm, CSET = nn + nm;
    // phase-1 shared memory (floats)
    // phi, BR and theta|theta^-1 live in the buffers of Q_k, Q_kp1 and Q_k^-1|Q_kp1, which are dead by the time they are written
    // (6.2 KB per block row instead of 9 KB at n = 14: eight rows' worth more resident per SM in the warp-per-row launch)
    static constexpr uint32_t P1_FLOATS = nn /*A*/ + nm /*B*/ + 2 * nn /*Qk|I*/ + 2 * nn /*Qkp1|I*/ + 2 * mm /*R|I*/ + nn /*BRBt*/ +
                                          6 * n + m + 3 + 3 * 32;                                      // + alignment slack + 3 pivot-row snapshots
    static constexpr uint32_t P1_STRIDE = (P1_FLOATS + 3) / 4 * 4;     // per warp in the warp-per-row kernel (NT / 32 rows per CTA)
    static constexpr uint32_t P2_FLOATS = 7 * nn;
    static constexpr uint32_t P2W_TILE = 16 * 20;                      // warp-per-row phase 2: tiles staged as 16 columns with ld 20
    static constexpr uint32_t P2W_FLOATS = (NT / 32) * 7 * P2W_TILE;
};

// ---- phase 1: one CTA per block row (linsys_setup.cuh:139-562); WR (batches): one WARP per block row, four rows per CTA.
// With one CTA per row the three inversions run side by side in three warps and the fourth idles, then one warp inverts theta
// while three idle: right for the latency of one trajectory, but in a batch 42 % of the resident warp-time sat at CTA barriers
// (ncu, 1024 trajectories).  WR runs the same statements with a warp as the whole team (the inversions one after the other, a
// __syncwarp where the CTA version has a barrier), so every resident warp always has work.  Same operations per element in both.
template <uint32_t n, uint32_t m, bool WR = false>
__global__ void __launch_bounds__(SchurShape<n, m>::NT, 8)
schur_phase1_kernel(uint32_t N, const float *__restrict__ G, const float *__restrict__ C, const float *__restrict__ g,
                    const float *__restrict__ c, float *__restrict__ S, float *__restrict__ Pinv, float *__restrict__ gamma, float rho)
{
    using namespace schur_detail;
    using K = SchurShape<n, m>;
    constexpr uint32_t nn = K::nn, mm = K::mm, nm = K::nm, NT = WR ? 32u : K::NT;
    auto team_sync = [] { if constexpr (WR) __syncwarp(); else __syncthreads(); };
    static_assert(n % 2 == 0, "the blocked products pair the rows of column-major n x n operands");
    extern __shared__ __align__(16) float sm[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // phase 2 may be scheduled; it waits for our completion
    const uint32_t lane = threadIdx.x & 31u, warp = WR ? 0u : threadIdx.x >> 5, t = WR ? lane : threadIdx.x;
    const uint32_t b = WR ? blockIdx.x * (K::NT / 32) + (threadIdx.x >> 5) : blockIdx.x;
    if (WR && b >= N) return;                                    // (a whole warp; WR has no CTA barriers)
    float *const sm0 = sm + (WR ? (threadIdx.x >> 5) * K::P1_STRIDE : 0u);
    float *sA = sm0, *sB = sA + nn, *sQk = sB + nm, *sQk_i = sQk + nn, *sQp = sQk_i + nn, *sQp_i = sQp + nn;
    float *sR = sQp_i + nn, *sR_i = sR + mm, *sBRBt = sR_i + mm;
    // aliases (each is written only after the last read of its host, with a team_sync in between): phi over Q_k (read by its
    // inversion only), BR over Q_kp1 (same), theta over Q_k^-1 (last read in stage A, parked for dz before that), and theta^-1
    // behind it over BR (last read in stage B)
    float *sPhi = sQk, *sBR = sQp, *sTh = sQk_i, *sTh_i = sTh + nn;
    static_assert(nm <= nn, "BR fits the Q_kp1 buffer");
    float *sqk = sBRBt + nn, *sqp = sqk + n, *srk = sqp + n, *sgam = srk + m, *sx0 = sgam + n, *sx1 = sx0 + n;
    float *sc = sx1 + n;
    float *snap = sm0 + ((sc + n - sm0) + 3) / 4 * 4;            // 3 snapshots of 32 floats, 16-byte aligned
    {   // blockIdx.y = system of a batch: every array carries a leading [batch] dimension
        const size_t sys = blockIdx.y;
        G += sys * ((size_t)K::GSET * (N - 1) + nn); C += sys * (size_t)K::CSET * (N - 1); g += sys * ((size_t)(n + m) * (N - 1) + n);
        c += sys * (size_t)n * N; S += sys * 3 * (size_t)nn * N; Pinv += sys * 3 * (size_t)nn * N; gamma += sys * (size_t)n * N;
    }
    float *Srow = S + (size_t)b * 3 * nn, *Prow = Pinv + (size_t)b * 3 * nn;

    if (b == 0) {
        // ---- leading block (:151-278): Pinv_00 = -(Q_0 + rho I), S_00 = -Q_0^-1, gamma_0 = -Q_0^-1 q_0
        for (uint32_t i = t; i < nn; i += NT) sQk[i] = (i % n == i / n) ? __fadd_rn(G[i], rho) : G[i];
        for (uint32_t i = t; i < n; i += NT) sqk[i] = g[i];
        team_sync();
        for (uint32_t i = t; i < nn; i += NT) Prow[nn + i] = sQk[i] * -1.0f;
        team_sync();
        if (warp == 0) gj_div_warp<n>(sQk, snap, lane);
        team_sync();
        for (uint32_t i = t; i < nn; i += NT) Srow[nn + i] = sQk_i[i] * -1.0f;
        for (uint32_t i = t; i < n; i += NT) gamma[i] = -matvec_elem(sQk_i, sqk, n, n, i);
        return;
    }
    // ---- block rows 1 .. N-1 (:280-560); the reference's "k" blocks are knot b-1, its "kp1" blocks knot b
    const float *Gk = G + (size_t)(b - 1) * K::GSET, *Gp = G + (size_t)b * K::GSET, *Ck = C + (size_t)(b - 1) * K::CSET;
    // (Q_k, Q_kp1 and R_k go from global memory straight into the registers of their inversions)
    // A, B and the vectors are first read in stage A: they travel global -> shared memory as 4-byte cp.async copies issued here
    // and awaited after the inversions, so their DRAM latency hides behind the pivot loops (the inversions read no staged data)
    auto cp4 = [](float *dst, const float *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    for (uint32_t i = t; i < nn; i += NT) cp4(sA + i, Ck + i);
    for (uint32_t i = t; i < nm; i += NT) cp4(sB + i, Ck + nn + i);
    for (uint32_t i = t; i < n; i += NT) {
        cp4(sqk + i, g + (size_t)(b - 1) * (n + m) + i);
        cp4(sqp + i, g + (size_t)b * (n + m) + i);
        cp4(sc + i, c + (size_t)b * n + i);
    }
    for (uint32_t i = t; i < m; i += NT) cp4(srk + i, g + (size_t)(b - 1) * (n + m) + n + i);
    // ---- the three inversions side by side, one warp each (:351-363)
    if constexpr (WR) {
        gj_div_pair_warp<n>(sQk, sQp, snap, lane, Gk, Gp, rho);       // both state-cost blocks in one pass, a half-warp each
        gj_div_warp<m>(sR, snap + 64, lane, Gk + nn, rho);
    } else {
        if (warp == 0) gj_div_warp<n>(sQk, snap, lane, Gk, rho);
        else if (warp == 1) gj_div_warp<n>(sQp, snap + 32, lane, Gp, rho);
        else if (warp == 2) gj_div_warp<m>(sR, snap + 64, lane, Gk + nn, rho);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    team_sync();
    // park the inverses for compute_dz in tiles that phase 2 overwrites (moved into G there): Q_{b-1}^-1 in the left
    // tile of row b, R_{b-1}^-1 in the right tile of row b-1, Q_{N-1}^-1 in the pad tile (left of row 0)
    for (uint32_t i = t; i < nn; i += NT) Prow[i] = sQk_i[i];
    for (uint32_t i = t; i < mm; i += NT) (Prow - 3 * nn)[2 * nn + i] = sR_i[i];
    if (b == N - 1)
        for (uint32_t i = t; i < nn; i += NT) Pinv[i] = sQp_i[i];
    // ---- stage A: phi = A Q_k^-1, BR = B R_k^-1, gam = Q_kp1^-1 q_kp1  (:385-413).  Register-blocked tasks (2 x 2, 2 x 1, two
    // rows); the task kinds start on warp boundaries (phi | BR, gam), so a stage is ONE pass over the CTA with two warps per kind
    constexpr uint32_t HN = n / 2, T_SQ = HN * HN, S_SQ = (T_SQ + 31) / 32 * 32;
    for (uint32_t slot = t; slot < S_SQ + HN * m + HN; slot += NT) {
        if (slot < S_SQ) {
            if (slot < T_SQ) {
                const uint32_t r = 2 * (slot % HN), cc = 2 * (slot / HN);
                float o[4];
                gemm_2x2<n, n, n, false>(sA, sQk_i, r, cc, o);
                *reinterpret_cast<float2 *>(sPhi + cc * n + r) = make_float2(o[0], o[1]);
                *reinterpret_cast<float2 *>(sPhi + (cc + 1) * n + r) = make_float2(o[2], o[3]);
            }
        } else if (slot < S_SQ + HN * m) {
            const uint32_t e = slot - S_SQ, r = 2 * (e % HN), cc = e / HN;
            float o[2];
            gemm_2x1<n, m>(sB, sR_i, r, cc, o);
            *reinterpret_cast<float2 *>(sBR + cc * n + r) = make_float2(o[0], o[1]);
        } else {
            const uint32_t r = 2 * (slot - S_SQ - HN * m);
            float o[2];
            matvec_2<n, n>(sQp_i, sqp, r, o);
            sgam[r] = o[0];
            sgam[r + 1] = o[1];
        }
    }
    team_sync();
    // ---- stage B: phi q_k, BR r_k, phi A^T, BR B^T  (:421-481)
    for (uint32_t slot = t; slot < S_SQ + T_SQ + 2 * HN; slot += NT) {
        if (slot < S_SQ) {
            if (slot < T_SQ) {
                const uint32_t r = 2 * (slot % HN), cc = 2 * (slot / HN);
                float o[4];
                gemm_2x2<n, n, n, true>(sPhi, sA, r, cc, o);
                *reinterpret_cast<float2 *>(sTh + cc * n + r) = make_float2(o[0], o[1]);
                *reinterpret_cast<float2 *>(sTh + (cc + 1) * n + r) = make_float2(o[2], o[3]);
            }
        } else if (slot < S_SQ + T_SQ) {
            const uint32_t e = slot - S_SQ, r = 2 * (e % HN), cc = 2 * (e / HN);
            float o[4];
            gemm_2x2<n, m, n, true>(sBR, sB, r, cc, o);
            *reinterpret_cast<float2 *>(sBRBt + cc * n + r) = make_float2(o[0], o[1]);
            *reinterpret_cast<float2 *>(sBRBt + (cc + 1) * n + r) = make_float2(o[2], o[3]);
        } else if (slot < S_SQ + T_SQ + HN) {
            const uint32_t r = 2 * (slot - S_SQ - T_SQ);
            float o[2];
            matvec_2<n, n>(sPhi, sqk, r, o);
            sx0[r] = o[0];
            sx0[r + 1] = o[1];
        } else {
            const uint32_t r = 2 * (slot - S_SQ - T_SQ - HN);
            float o[2];
            matvec_2<n, m>(sBR, srk, r, o);
            sx1[r] = o[0];
            sx1[r + 1] = o[1];
        }
    }
    team_sync();
    // ---- stage C: theta = (phi A^T + Q_kp1^-1) + BR B^T ; gamma ; S tiles  (:417, :441-443, :466-500, :527-560)
    for (uint32_t i = t; i < nn; i += NT) {
        const float th = __fadd_rn(__fadd_rn(sTh[i], sQp_i[i]), sBRBt[i]);
        sTh[i] = th;
        Srow[nn + i] = th * -1.0f;
        Srow[i] = sPhi[i] * -1.0f;
        (Srow - 3 * nn)[2 * nn + (i % n) * n + i / n] = sPhi[i] * -1.0f;      // phi^T: right tile of row b-1
    }
    for (uint32_t i = t; i < n; i += NT) {
        const float gm = __fadd_rn(__fadd_rn(sgam[i], -sc[i]), __fadd_rn(sx1[i], sx0[i]));
        gamma[(size_t)b * n + i] = gm * -1.0f;
    }
    team_sync();
    // ---- theta^-1 (:503-518)
    if (warp == 0) gj_rcp_warp<n>(sTh, snap, lane);
    team_sync();
    for (uint32_t i = t; i < nn; i += NT) Prow[nn + i] = sTh_i[i] * -1.0f;
}


// ---- phase 2: off-diagonal tiles of Pinv (linsys_setup.cuh:9-137), plus moving the parked inverses into G
template <uint32_t n, uint32_t m>
__global__ void __launch_bounds__(SchurShape<n, m>::NT)
schur_phase2_kernel(uint32_t N, float *__restrict__ G, const float *__restrict__ S, float *__restrict__ Pinv)
{
    using namespace schur_detail;
    using K = SchurShape<n, m>;
    constexpr uint32_t nn = K::nn, mm = K::mm, NT = K::NT;
    extern __shared__ __align__(16) float sm[];
    float *sTk = sm, *sTm = sTk + nn, *sTp = sTm + nn, *sPhik = sTp + nn, *sPhiT = sPhik + nn, *sL = sPhiT + nn, *sRr = sL + nn;
    const uint32_t t = threadIdx.x, b = blockIdx.x;
    {
        const size_t sys = blockIdx.y;
        G += sys * ((size_t)K::GSET * (N - 1) + nn); S += sys * 3 * (size_t)nn * N; Pinv += sys * 3 * (size_t)nn * N;
    }
    float *Prow = Pinv + (size_t)b * 3 * nn;
    const bool has_l = b != 0, has_r = b != N - 1;
    // launched with programmatic stream serialization: this grid may be scheduled while phase 1 drains; everything
    // phase 1 wrote is visible after this wait
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // parked inverses -> G (this CTA owns the tiles they are parked in)
    if (has_l)
        for (uint32_t i = t; i < nn; i += NT) G[(size_t)(b - 1) * K::GSET + i] = Prow[i];
    if (has_r)
        for (uint32_t i = t; i < mm; i += NT) G[(size_t)b * K::GSET + nn + i] = Prow[2 * nn + i];
    if (b == 0)
        for (uint32_t i = t; i < nn; i += NT) G[(size_t)(N - 1) * K::GSET + i] = Prow[i];
    for (uint32_t i = t; i < nn; i += NT) {
        sTk[i] = Prow[nn + i];
        if (has_l) { sTm[i] = (Prow - 3 * nn)[nn + i]; sPhik[i] = S[(size_t)b * 3 * nn + i]; }
        if (has_r) { sTp[i] = (Prow + 3 * nn)[nn + i]; sPhiT[(i % n) * n + i / n] = S[(size_t)(b + 1) * 3 * nn + i]; }
    }
    __syncthreads();
    // register-blocked 2 x 2 tasks, the left and the right product on warp boundaries (see gemm_2x2)
    constexpr uint32_t HN = n / 2, T_SQ = HN * HN, S_SQ = (T_SQ + 31) / 32 * 32;
    for (uint32_t slot = t; slot < S_SQ + T_SQ; slot += NT) {
        const bool lft = slot < S_SQ;
        const uint32_t e = lft ? slot : slot - S_SQ, r = 2 * (e % HN), cc = 2 * (e / HN);
        if (e >= T_SQ || !(lft ? has_l : has_r)) continue;
        float o[4];
        gemm_2x2<n, n, n, false>(sTk, lft ? sPhik : sPhiT, r, cc, o);
        float *dst = lft ? sL : sRr;
        *reinterpret_cast<float2 *>(dst + cc * n + r) = make_float2(o[0], o[1]);
        *reinterpret_cast<float2 *>(dst + (cc + 1) * n + r) = make_float2(o[2], o[3]);
    }
    __syncthreads();
    for (uint32_t slot = t; slot < S_SQ + T_SQ; slot += NT) {
        const bool lft = slot < S_SQ;
        const uint32_t e = lft ? slot : slot - S_SQ, r = 2 * (e % HN), cc = 2 * (e / HN);
        if (e >= T_SQ || !(lft ? has_l : has_r)) continue;
        float o[4];
        gemm_2x2<n, n, n, false>(lft ? sL : sRr, lft ? sTm : sTp, r, cc, o);
        float *dst = Prow + (lft ? 0u : 2 * nn);
        dst[cc * n + r] = o[0] * -1.0f;
        dst[cc * n + r + 1] = o[1] * -1.0f;
        dst[(cc + 1) * n + r] = o[2] * -1.0f;
        dst[(cc + 1) * n + r + 1] = o[3] * -1.0f;
    }
}

// ---- phase 2 for batches: one WARP per block row, four rows per CTA, no CTA barrier.  ncu on the CTA mapping at 1024
// trajectories: the shared-memory data pipe 92 % busy -- the kernel is bound by the operand loads of its four 14 x 14 x 14
// products (2 x 2 register blocks: 8 64-bit loads per 16 FMAs).  Here the tiles are staged with a leading dimension of 20 (16 rows
// + 4 pad: a multiple of 4 that keeps the 128-bit accesses of a quarter-warp on distinct banks), so
// four rows of a column are one aligned 128-bit load, and a lane owns a 4 x 4 block of outputs: 2 128-bit loads per 16 FMAs.
// Lanes 0..15 do the left product, lanes 16..31 the right one (16 blocks cover a 16 x 16 tile; the pad rows / columns are
// computed on whatever the pads hold and never stored).  phi_{b+1}^T is not transposed while staging: the right product reads
// phi_{b+1} as the transposed operand.  Every output is still one FMA per term in ascending k from 0.0f -> same bits.
namespace schur_detail {
// o[i][j] = sum_k A(r0 + i, k) * Bop(k, c0 + j);  A column-major with ld LD;  TB = false: B column-major (K x NC, ld LD),
// TB = true: Bop = B^T with B column-major (NC x K, ld LD)
constexpr uint32_t LD4 = 20;
template <uint32_t K, bool TB>
__device__ __forceinline__ void gemm_4x4_ld(const float *A, const float *B, uint32_t r0, uint32_t c0, float (&o)[4][4])
{
#pragma unroll
    for (uint32_t i = 0; i < 4; ++i)
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) o[i][j] = 0.0f;
#pragma unroll
    for (uint32_t kg = 0; kg < (K + 3) / 4; ++kg) {
        float bq[4][4];                                  // [j][kk] = Bop(4 kg + kk, c0 + j)
        if constexpr (!TB) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                const float4 f = *reinterpret_cast<const float4 *>(B + (c0 + j) * LD4 + 4 * kg);
                bq[j][0] = f.x; bq[j][1] = f.y; bq[j][2] = f.z; bq[j][3] = f.w;
            }
        }
#pragma unroll
        for (uint32_t kk = 0; kk < 4; ++kk) {
            const uint32_t k = 4 * kg + kk;
            if (k < K) {
                const float4 a4 = *reinterpret_cast<const float4 *>(A + k * LD4 + r0);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                float bb[4];
                if constexpr (TB) {
                    const float4 f = *reinterpret_cast<const float4 *>(B + k * LD4 + c0);
                    bb[0] = f.x; bb[1] = f.y; bb[2] = f.z; bb[3] = f.w;
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) bb[j] = bq[j][kk];
                }
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i)
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) o[i][j] = fma_(a[i], bb[j], o[i][j]);
            }
        }
    }
}
}  // namespace schur_detail

template <uint32_t n, uint32_t m>
__global__ void __launch_bounds__(SchurShape<n, m>::NT)
schur_phase2_warp_kernel(uint32_t N, float *__restrict__ G, const float *__restrict__ S, float *__restrict__ Pinv)
{
    using namespace schur_detail;
    using K = SchurShape<n, m>;
    static_assert(n <= 16, "a tile is staged as 16 x 16");
    constexpr uint32_t nn = K::nn, mm = K::mm, TS = K::P2W_TILE;
    extern __shared__ __align__(16) float sm[];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, b = blockIdx.x * (K::NT / 32) + wib;
    {
        const size_t sys = blockIdx.y;
        G += sys * ((size_t)K::GSET * (N - 1) + nn); S += sys * 3 * (size_t)nn * N; Pinv += sys * 3 * (size_t)nn * N;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");           // everything phase 1 wrote is visible after this wait
    if (b >= N) return;
    float *sTk = sm + wib * (7 * TS), *sTm = sTk + TS, *sTp = sTm + TS, *sPhik = sTp + TS, *sPhiN = sPhik + TS, *sL = sPhiN + TS, *sRr = sL + TS;
    float *Prow = Pinv + (size_t)b * 3 * nn;
    const bool has_l = b != 0, has_r = b != N - 1;
    // With one warp per row there are few warps per SM to hide DRAM latency behind, so every global read of the row is in flight at
    // once: the five operand tiles go global -> shared memory as 4-byte cp.async copies (no registers, no round trip per loop
    // iteration), the parked inverses as fully unrolled loads ahead of their stores.
    auto cp4 = [](float *dst, const float *src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    constexpr uint32_t RN = (nn + 31) / 32, RM = (mm + 31) / 32;
#pragma unroll
    for (uint32_t it = 0; it < RN; ++it) {
        const uint32_t i = lane + 32 * it;
        if (i < nn) {
            const uint32_t d = (i / n) * LD4 + i % n;             // column-major, leading dimension LD4
            cp4(sTk + d, Prow + nn + i);
            if (has_l) { cp4(sTm + d, Prow - 3 * nn + nn + i); cp4(sPhik + d, S + (size_t)b * 3 * nn + i); }
            if (has_r) { cp4(sTp + d, Prow + 3 * nn + nn + i); cp4(sPhiN + d, S + (size_t)(b + 1) * 3 * nn + i); }
        }
    }
    // parked inverses -> G (this warp owns the tiles they are parked in; the same lane reads an element here and overwrites it below)
    {
        float ql[RN], qr[RM];
#pragma unroll
        for (uint32_t it = 0; it < RN; ++it) { const uint32_t i = lane + 32 * it; ql[it] = ((has_l || b == 0) && i < nn) ? Prow[i] : 0.0f; }
#pragma unroll
        for (uint32_t it = 0; it < RM; ++it) { const uint32_t i = lane + 32 * it; qr[it] = (has_r && i < mm) ? Prow[2 * nn + i] : 0.0f; }
        float *Gl = G + (size_t)(has_l ? b - 1 : N - 1) * K::GSET;           // row 0 holds Q_{N-1}^-1 in its pad tile
#pragma unroll
        for (uint32_t it = 0; it < RN; ++it) { const uint32_t i = lane + 32 * it; if (i < nn) Gl[i] = ql[it]; }
#pragma unroll
        for (uint32_t it = 0; it < RM; ++it) { const uint32_t i = lane + 32 * it; if (has_r && i < mm) G[(size_t)b * K::GSET + nn + i] = qr[it]; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const bool rgt = lane >= 16, mine = rgt ? has_r : has_l;
    const uint32_t r0 = 4 * (lane & 3u), c0 = 4 * ((lane >> 2) & 3u);
    float o[4][4];
    if (mine) {
        // left: theta_k^-1 phi_k ; right: theta_k^-1 phi_{k+1}^T
        if (rgt) gemm_4x4_ld<n, true>(sTk, sPhiN, r0, c0, o);
        else gemm_4x4_ld<n, false>(sTk, sPhik, r0, c0, o);
        float *dst = rgt ? sRr : sL;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) *reinterpret_cast<float4 *>(dst + (c0 + j) * LD4 + r0) = make_float4(o[0][j], o[1][j], o[2][j], o[3][j]);
    }
    __syncwarp();
    if (mine) {
        // ... times the neighbour's theta^-1, negated; parked in the (dead) phi buffers for a coalesced write
        gemm_4x4_ld<n, false>(rgt ? sRr : sL, rgt ? sTp : sTm, r0, c0, o);
        float *dst = rgt ? sPhiN : sPhik;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            *reinterpret_cast<float4 *>(dst + (c0 + j) * LD4 + r0) = make_float4(o[0][j] * -1.0f, o[1][j] * -1.0f, o[2][j] * -1.0f, o[3][j] * -1.0f);
    }
    __syncwarp();
    for (uint32_t i = lane; i < nn; i += 32) {
        const uint32_t d = (i / n) * LD4 + i % n;
        if (has_l) Prow[i] = sPhik[d];
        if (has_r) Prow[2 * nn + i] = sPhiN[d];
    }
}

// ---- dz (dz.cuh:3-136): one CTA per knot does the state row and (k < N-1) the control row
template <uint32_t n, uint32_t m>
__global__ void __launch_bounds__(64)
compute_dz_kernel(uint32_t N, const float *__restrict__ Ginv, const float *__restrict__ C, const float *__restrict__ g,
                  const float *__restrict__ lambda, float *__restrict__ dz)
{
    using namespace schur_detail;
    using K = SchurShape<n, m>;
    constexpr uint32_t nn = K::nn, NT = 64;
    __shared__ float sx[n], su[m > 0 ? m : 1], sl[n];
    const uint32_t t = threadIdx.x, k = blockIdx.x;
    {
        const size_t sys = blockIdx.y;
        Ginv += sys * ((size_t)K::GSET * (N - 1) + nn); C += sys * (size_t)K::CSET * (N - 1); g += sys * ((size_t)(n + m) * (N - 1) + n);
        lambda += sys * (size_t)n * N; dz += sys * ((size_t)(n + m) * (N - 1) + n);
    }
    const bool last = k == N - 1;
    for (uint32_t i = t; i < n; i += NT) sl[i] = last ? 0.0f : lambda[(size_t)(k + 1) * n + i];
    __syncthreads();
    const float *A = C + (size_t)k * K::CSET, *B = A + nn;
    for (uint32_t task = t; task < n + m; task += NT) {
        if (task < n) {
            float res = 0.0f;
            if (!last)
                for (uint32_t e = 0; e < n; ++e) res = fma_(A[task * n + e], sl[e], res);                  // (A^T lambda_{k+1})
            const float s = __fadd_rn(lambda[(size_t)k * n + task], res);
            sx[task] = __fadd_rn(g[(size_t)k * (n + m) + task], -s);
        } else if (!last) {
            const uint32_t i = task - n;
            float res = 0.0f;
            for (uint32_t e = 0; e < n; ++e) res = fma_(B[i * n + e], sl[e], res);                          // (B^T lambda_{k+1})
            su[i] = __fadd_rn(g[(size_t)k * (n + m) + n + i], -res);
        }
    }
    __syncthreads();
    const float *Qi = Ginv + (size_t)k * K::GSET, *Ri = Qi + nn;
    for (uint32_t task = t; task < n + m; task += NT) {
        if (task < n) dz[(size_t)k * (n + m) + task] = matvec_elem(Qi, sx, n, n, task);
        else if (!last) dz[(size_t)k * (n + m) + n + (task - n)] = matvec_elem(Ri, su, m, m, task - n);
    }
}

// ---- row f4 (wire format): the band S -> upper-triangular CSC of the symmetric block-tridiagonal matrix, the format the
// reference hands to QDLDL (include/utils/csr.cuh:10-74: prep_csr builds col_ptr / row_ind, store_block_csr_lowertri the
// values; include/qdldl/sqp.cuh:148,164).  Column j = (block row b, row r) holds rows (b-1)n .. (b-1)n+n-1 (row r of the left
// tile) then rows bn .. bn+r (row r of the diagonal tile up to the diagonal).  Pure index work: bit-exact by construction.
__host__ __device__ inline uint32_t csr_nnz(uint32_t n, uint32_t N) { return (N - 1) * n * n + N * ((n + 1) * n / 2); }
__host__ __device__ inline uint32_t csr_col_offset(uint32_t n, uint32_t b, uint32_t r)
{
    const uint32_t tri = (n + 1) * n / 2, brow = n * n + tri;
    return (b > 0 ? tri + (b - 1) * brow + r * n : 0u) + (r + 1) * r / 2;
}
__global__ void __launch_bounds__(64)
csr_pattern_kernel(uint32_t n, uint32_t N, int32_t *__restrict__ col_ptr, int32_t *__restrict__ row_ind)
{
    for (uint32_t b = blockIdx.x; b < N; b += gridDim.x)
        for (uint32_t r = threadIdx.x; r < n; r += blockDim.x) {
            if (b == 0 && r == 0) col_ptr[0] = 0;
            const uint32_t off = csr_col_offset(n, b, r), len = (b > 0 ? n : 0u) + r + 1;
            col_ptr[b * n + r + 1] = (int32_t)(off + len);
            for (uint32_t c = 0; c < len; ++c) row_ind[off + c] = (int32_t)((b > 0 ? (b - 1) * n : 0u) + c);
        }
}
__global__ void __launch_bounds__(128)
csr_values_kernel(uint32_t n, uint32_t N, const float *__restrict__ S, float *__restrict__ val)
{
    const uint32_t b = blockIdx.x;
    const float *L = S + (size_t)b * 3 * n * n, *D = L + n * n;
    // one (row, col) element per thread iteration; consecutive threads walk a column of the tile (coalesced reads)
    for (uint32_t e = threadIdx.x; e < 2 * n * n; e += blockDim.x) {
        const bool diag = e >= n * n;
        const uint32_t q = diag ? e - n * n : e, r = q % n, c = q / n;
        const uint32_t off = csr_col_offset(n, b, r);
        if (!diag) { if (b > 0) val[off + c] = L[r + c * n]; }
        else if (c <= r) val[off + (b > 0 ? n : 0u) + c] = D[r + c * n];
    }
}

}  // namespace gbd
