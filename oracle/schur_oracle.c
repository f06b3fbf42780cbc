/* schur_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in the reference's floating-point operation order, of the two steps that sit either side of the
 * PCG solve inside one SQP iteration:
 *
 *   schur_oracle_form_f32  =  form_schur_system<float>            include/pcg/linsys_setup.cuh:621-657
 *        phase 1  form_S_gamma_and_jacobi_Pinv_blockrow            include/pcg/linsys_setup.cuh:139-562
 *        phase 2  complete_SS_Pinv_blockrow                        include/pcg/linsys_setup.cuh:9-137
 *        with     invertMatrix (1 / 2 / 3 matrices at once)        include/utils/matrix.cuh:120-238
 *                 glass::gemm (plain and TRANSPOSE_B)              GLASS/src/L3/gemm.cuh:47-96
 *                 mat_vec_prod, add_identity, loadIdentity         include/utils/matrix.cuh:42-118
 *                 store_block_bd / load_block_bd                   GBD-PCG/include/utils.cuh:96-160
 *   schur_oracle_dz_f32    =  compute_dz<float>                    include/common/dz.cuh:3-136
 *        with     gato_ATx, gato_vec_sum, gato_vec_dif             include/utils/matrix.cuh:9-40
 *
 * Where nvcc contracts a multiply-add of the reference into one FFMA (checked in the sm_100a SASS of the reference
 * kernels with line info: every `res += a*b` dot product, and `x -= (a/b)*c` / `x -= (a*pvInv)*c` in the Gauss-Jordan
 * updates) this file calls fmaf(); everything else is a separately rounded operation (build with -ffp-contract=off).
 * Divisions are IEEE (the reference is compiled without -use_fast_math).
 *
 * Layouts (all column-major inside a block, as the reference):
 *   G  : per knot k < N-1: [Q_k (n*n) | R_k (m*m)], then Q_{N-1} (n*n)          -> overwritten with the inverses
 *   C  : per knot k < N-1: [A_k (n*n) | B_k (n*m)]
 *   g  : per knot k < N-1: [q_k (n) | r_k (m)], then q_{N-1} (n)
 *   c  : per knot n values
 *   S, Pinv : [N][3][n][n] (left | diag | right), gamma : [N*n]   (the pcg<> input layout)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

/* C (m x k) = A (m x n) * B (n x k), or A * B^T with B stored k x n; column-major; one FMA per term, ind ascending */
static void gemm(uint32_t m, uint32_t n, uint32_t k, const float *A, const float *B, float *C, int transpose_b)
{
    for (uint32_t col = 0; col < k; ++col)
        for (uint32_t row = 0; row < m; ++row) {
            float res = 0.0f;
            for (uint32_t ind = 0; ind < n; ++ind)
                res = fmaf(A[ind * m + row], transpose_b ? B[ind * k + col] : B[col * n + ind], res);
            C[col * m + row] = res;                 /* alpha = 1 */
        }
}

static void mat_vec_prod(uint32_t rows, uint32_t cols, const float *mat, const float *vec, float *out)
{
    for (uint32_t row = 0; row < rows; ++row) {
        float res = 0.0f;
        for (uint32_t col = 0; col < cols; ++col) res = fmaf(mat[row + col * rows], vec[col], res);
        out[row] = res;
    }
}

static void add_identity(float *A, uint32_t dim, float factor)
{
    for (uint32_t i = 0; i < dim; ++i) A[i * dim + i] = A[i * dim + i] + factor;
}

static void load_identity(uint32_t dim, float *A)
{
    for (uint32_t i = 0; i < dim * dim; ++i) A[i] = (i % dim == i / dim) ? 1.0f : 0.0f;
}

/* Gauss-Jordan on [V | I] (dim x 2dim, column-major), the several-matrices-at-once form (matrix.cuh:149-238):
 * per pivot p: snapshot pivot column and the dim+1 row entries from column p on, then for the dim+1 columns p..p+dim
 *   row == p : x /= piv          else : x = fma(-(col_snapshot[row] / piv), row_snapshot[col], x)                  */
static void invert_div(uint32_t dim, float *A)
{
    float *colv = (float *)malloc(sizeof(float) * (2 * dim + 1)), *rowv = colv + dim;
    for (uint32_t p = 0; p < dim; ++p) {
        const uint32_t off = p * dim;
        for (uint32_t i = 0; i < dim; ++i) colv[i] = A[i + off];
        for (uint32_t i = 0; i < dim + 1; ++i) rowv[i] = A[i * dim + p + off];
        for (uint32_t ind = 0; ind < dim * (dim + 1); ++ind) {
            const uint32_t row = ind % dim, col = ind / dim;
            if (row == p) A[off + ind] = A[off + ind] / colv[p];
            else A[off + ind] = fmaf(-(colv[row] / colv[p]), rowv[col], A[off + ind]);
        }
    }
    free(colv);
}

/* the single-matrix form (matrix.cuh:120-146): pvInv = 1 / piv;  row == p : x *= pvInv
 *                                              else : x = fma(-(col_snapshot[row] * pvInv), row_snapshot[col], x)   */
static void invert_rcp(uint32_t dim, float *A)
{
    float *colv = (float *)malloc(sizeof(float) * (2 * dim + 1)), *rowv = colv + dim;
    for (uint32_t p = 0; p < dim; ++p) {
        const uint32_t off = p * dim;
        const float pv_inv = 1.0f / A[p + off];
        for (uint32_t i = 0; i < dim; ++i) colv[i] = A[i + off];
        for (uint32_t i = 0; i < dim + 1; ++i) rowv[i] = A[p + off + i * dim];
        for (uint32_t ind = 0; ind < dim * (dim + 1); ++ind) {
            const uint32_t row = ind % dim, col = ind / dim;
            if (row == p) A[off + ind] = A[off + ind] * pv_inv;
            else A[off + ind] = fmaf(-(colv[row] * pv_inv), rowv[col], A[off + ind]);
        }
    }
    free(colv);
}

static void store_block(uint32_t n, const float *src, float *dst, uint32_t col, uint32_t brow, float mult)
{
    float *d = dst + (size_t)brow * 3 * n * n + (size_t)col * n * n;
    for (uint32_t i = 0; i < n * n; ++i) d[i] = src[i] * mult;
}

EXPORT int schur_oracle_form_f32(uint32_t n, uint32_t m, uint32_t N, float *G, const float *C, const float *g, const float *c,
                                 float *S, float *Pinv, float *gamma, float rho)
{
    const uint32_t nn = n * n, mm = m * m, nm = n * m, Gset = nn + mm, Cset = nn + nm;
    float *buf = (float *)malloc(sizeof(float) * (12 * (size_t)nn + 4 * mm + 2 * nm + 8 * n + 2 * m));
    if (!buf) return -1;
    float *Gin = (float *)malloc(sizeof(float) * ((size_t)Gset * N));      /* all block rows read G before any of them */
    if (!Gin) { free(buf); return -1; }                                    /* overwrites it (co-resident CTAs)         */
    memcpy(Gin, G, sizeof(float) * ((size_t)Gset * (N - 1) + nn));
    float *phi = buf, *theta = phi + nn, *thetaInv = theta + nn, *gam = thetaInv + nn;
    float *Ak = gam + n, *Bk = Ak + nn, *Qk = Bk + nm, *Qk_i = Qk + nn, *Qkp1 = Qk_i + nn, *Qkp1_i = Qkp1 + nn;
    float *Rk = Qkp1_i + nn, *Rk_i = Rk + mm, *qk = Rk_i + mm, *qkp1 = qk + n, *rk = qkp1 + n, *extra = rk + m;
    float *tmpT = extra + 2 * n;                                           /* nn */

    /* ---------------- phase 1, block row 0 (linsys_setup.cuh:151-278) */
    {
        float *Q0 = Qk, *Q0_i = Qk_i;
        memcpy(Q0, Gin, sizeof(float) * nn);
        add_identity(Q0, n, rho);
        store_block(n, Q0, Pinv, 1, 0, -1.0f);                             /* -(Q_0 + rho I) in PhiInv spot 00 */
        load_identity(n, Q0_i);
        invert_div(n, Q0);                                                 /* two-matrix form: same arithmetic per matrix */
        store_block(n, Q0_i, S, 1, 0, -1.0f);
        mat_vec_prod(n, n, Q0_i, g, extra);
        for (uint32_t i = 0; i < n; ++i) gamma[i] = -extra[i];
    }
    /* ---------------- phase 1, block rows 1 .. N-1 (linsys_setup.cuh:280-560) */
    for (uint32_t b = 1; b < N; ++b) {
        memcpy(Ak, C + (size_t)(b - 1) * Cset, sizeof(float) * nn);
        memcpy(Bk, C + (size_t)(b - 1) * Cset + nn, sizeof(float) * nm);
        memcpy(Qk, Gin + (size_t)(b - 1) * Gset, sizeof(float) * nn);
        memcpy(Qkp1, Gin + (size_t)b * Gset, sizeof(float) * nn);
        memcpy(Rk, Gin + (size_t)(b - 1) * Gset + nn, sizeof(float) * mm);
        memcpy(qk, g + (size_t)(b - 1) * (n + m), sizeof(float) * n);
        memcpy(qkp1, g + (size_t)b * (n + m), sizeof(float) * n);
        memcpy(rk, g + (size_t)(b - 1) * (n + m) + n, sizeof(float) * m);
        add_identity(Qk, n, rho);
        add_identity(Qkp1, n, rho);
        add_identity(Rk, m, rho);
        load_identity(n, Qk_i);
        load_identity(n, Qkp1_i);
        load_identity(m, Rk_i);
        invert_div(n, Qk);
        invert_div(n, Qkp1);
        invert_div(m, Rk);
        memcpy(G + (size_t)(b - 1) * Gset, Qk_i, sizeof(float) * nn);
        memcpy(G + (size_t)(b - 1) * Gset + nn, Rk_i, sizeof(float) * mm);
        if (b == N - 1) memcpy(G + (size_t)b * Gset, Qkp1_i, sizeof(float) * nn);

        gemm(n, n, n, Ak, Qk_i, phi, 0);                                   /* A Q^-1 */
        gemm(n, m, m, Bk, Rk_i, Qkp1, 0);                                  /* B R^-1  (into the Qkp1 buffer) */
        mat_vec_prod(n, n, Qkp1_i, qkp1, gam);
        for (uint32_t i = 0; i < n; ++i) gam[i] = gam[i] - c[(size_t)b * n + i];
        mat_vec_prod(n, n, phi, qk, extra);
        mat_vec_prod(n, m, Qkp1, rk, extra + n);
        for (uint32_t i = 0; i < n; ++i) gam[i] = gam[i] + (extra[n + i] + extra[i]);
        gemm(n, n, n, phi, Ak, theta, 1);                                  /* A Q^-1 A^T */
        for (uint32_t i = 0; i < nn; ++i) theta[i] = theta[i] + Qkp1_i[i];
        gemm(n, m, n, Qkp1, Bk, Qkp1_i, 1);                                /* B R^-1 B^T (into the Qkp1_i buffer) */
        for (uint32_t i = 0; i < nn; ++i) theta[i] = theta[i] + Qkp1_i[i];
        store_block(n, phi, S, 0, b, -1.0f);
        store_block(n, theta, S, 1, b, -1.0f);
        load_identity(n, thetaInv);
        invert_rcp(n, theta);                                              /* theta | thetaInv are contiguous */
        store_block(n, thetaInv, Pinv, 1, b, -1.0f);
        for (uint32_t i = 0; i < n; ++i) gamma[(size_t)b * n + i] = gam[i] * -1.0f;
        /* phi^T via gemm<TRANSPOSE_B>(I, phi): sums of exact zeros around one exact product */
        load_identity(n, Ak);
        gemm(n, n, n, Ak, phi, tmpT, 1);
        store_block(n, tmpT, S, 2, b - 1, -1.0f);
    }
    /* ---------------- phase 2 (linsys_setup.cuh:9-137), on the STORED (negated) tiles */
    {
        float *scr = tmpT, *res = theta, *phiT = phi;
        float *Pout = (float *)malloc(sizeof(float) * 2 * (size_t)nn * N);  /* off-diagonal tiles read only diagonals */
        if (!Pout) { free(buf); free(Gin); return -1; }
        for (uint32_t b = 0; b < N; ++b) {
            const float *Tk = Pinv + (size_t)b * 3 * nn + nn;
            if (b != 0) {
                const float *phik = S + (size_t)b * 3 * nn, *Tkm1 = Pinv + (size_t)(b - 1) * 3 * nn + nn;
                gemm(n, n, n, Tk, phik, scr, 0);
                gemm(n, n, n, scr, Tkm1, res, 0);
                for (uint32_t i = 0; i < nn; ++i) Pout[(size_t)(2 * b) * nn + i] = res[i] * -1.0f;
            }
            if (b != N - 1) {
                const float *src = S + (size_t)(b + 1) * 3 * nn, *Tkp1 = Pinv + (size_t)(b + 1) * 3 * nn + nn;
                for (uint32_t ind = 0; ind < nn; ++ind) phiT[(ind % n) * n + ind / n] = src[ind];   /* transposed load */
                gemm(n, n, n, Tk, phiT, scr, 0);
                gemm(n, n, n, scr, Tkp1, res, 0);
                for (uint32_t i = 0; i < nn; ++i) Pout[(size_t)(2 * b + 1) * nn + i] = res[i] * -1.0f;
            }
        }
        for (uint32_t b = 0; b < N; ++b) {
            if (b != 0) memcpy(Pinv + (size_t)b * 3 * nn, Pout + (size_t)(2 * b) * nn, sizeof(float) * nn);
            if (b != N - 1) memcpy(Pinv + (size_t)b * 3 * nn + 2 * nn, Pout + (size_t)(2 * b + 1) * nn, sizeof(float) * nn);
        }
        free(Pout);
    }
    free(buf);
    free(Gin);
    return 0;
}

/* dz.cuh:3-136.  Ginv is G after schur_oracle_form_f32 (inverses); dz has the layout of g. */
EXPORT int schur_oracle_dz_f32(uint32_t n, uint32_t m, uint32_t N, const float *Ginv, const float *C, const float *g,
                               const float *lambda, float *dz)
{
    const uint32_t nn = n * n, mm = m * m, nm = n * m, Gset = nn + mm, Cset = nn + nm;
    float *scr = (float *)malloc(sizeof(float) * 2 * (n + m));
    if (!scr) return -1;
    for (uint32_t set = 0; set < N; ++set) {
        /* state row: dz_x = Q^-1 (q - (lambda_k + A^T lambda_{k+1})) */
        if (set != N - 1) {
            const float *A = C + (size_t)set * Cset;
            for (uint32_t ind = 0; ind < n; ++ind) {                       /* gato_ATx: out[ind] = sum_t mat[ind*n+t] vec[t] */
                float res = 0.0f;
                for (uint32_t t = 0; t < n; ++t) res = fmaf(A[ind * n + t], lambda[(size_t)(set + 1) * n + t], res);
                scr[ind] = res;
            }
        } else {
            for (uint32_t i = 0; i < n; ++i) scr[i] = 0.0f;
        }
        for (uint32_t i = 0; i < n; ++i) scr[i] = lambda[(size_t)set * n + i] + scr[i];
        for (uint32_t i = 0; i < n; ++i) scr[i] = g[(size_t)set * (n + m) + i] - scr[i];
        mat_vec_prod(n, n, Ginv + (size_t)set * Gset, scr, dz + (size_t)set * (n + m));
        /* control row: dz_u = R^-1 (r - B^T lambda_{k+1}) */
        if (set != N - 1) {
            const float *B = C + (size_t)set * Cset + nn;
            for (uint32_t ind = 0; ind < m; ++ind) {
                float res = 0.0f;
                for (uint32_t t = 0; t < n; ++t) res = fmaf(B[ind * n + t], lambda[(size_t)(set + 1) * n + t], res);
                scr[ind] = res;
            }
            for (uint32_t i = 0; i < m; ++i) scr[i] = g[(size_t)set * (n + m) + n + i] - scr[i];
            mat_vec_prod(m, m, Ginv + (size_t)set * Gset + nn, scr, dz + (size_t)set * (n + m) + n);
        }
    }
    free(scr);
    return 0;
}
