/*
 * pcg_oracle.c -- CPU restatement of the reference GBD-PCG solve.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  It is never linked into, imported by or
 * executed from the product (mpcgpu_b200/, include/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * It follows, operation by operation and in the same floating-point order, the reference kernel
 *   GBD-PCG/include/pcg.cuh:98-217      (algorithm, update order, exit rule, outputs)
 *   GBD-PCG/include/utils.cuh:9-85      (loadbdVec window, bdmv edge handling, column-major tiles,
 *                                        sequential accumulation over c)
 *   GLASS/src/L1/dot.cuh:52-63          (elementwise product, then tree)
 *   GLASS/src/L1/reduce.cuh:5-33,36-69  (halving tree with odd fix-up, serial tail of <=3)
 * nvcc contracts `val += a*b`, `x += alpha*p`, `r -= alpha*u`, `p = rt + beta*p` into single FMAs
 * (default -fmad=true; confirmed in the sm_100a SASS of the unmodified header), so the oracle uses
 * fmaf()/fma() at exactly those sites when `contract` != 0 and separate mul/add otherwise.
 * Build with -ffp-contract=off so that the compiler adds no fusions of its own.
 *
 * Parity pin: the reference has no golden vectors for PCG itself.  The oracle
 * is pinned (a) on the GLASS known answers (dot=656700, reduce=4950, GLASS/GTests/test.cu:157-171,
 * 275-282), (b) on the GBD-PCG demo system (GBD-PCG/examples/pcg_solve.cu:14-25) against an fp64
 * direct solve, and (c) bit-for-bit against the UNMODIFIED reference kernel run on a B200
 * (oracle/_ref/libref_gbdpcg.so; outputs committed under tests/golden/).
 *
 * Layout: S, Pinv are [N][3][n][n]; tile t of block row b at b*3n^2 + t*n^2,
 * column-major inside a tile (elem(r,c) at c*n + r); tile (0,0) and (N-1,2) are never read.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ fp32 */

/* GLASS/src/L1/reduce.cuh:5-33 : in-place halving tree, result in x[0] */
ORACLE_API float glass_reduce_f32(uint32_t n, float *x)
{
    uint32_t size_left = n;
    while (size_left > 3) {
        uint32_t odd = size_left % 2;
        size_left = (size_left - odd) / 2;
        for (uint32_t i = 0; i < size_left; i++) x[i] += x[i + size_left];
        if (odd) x[0] += x[2 * size_left];
    }
    for (uint32_t i = 1; i < size_left; i++) x[0] += x[i];
    return x[0];
}

/* GLASS/src/L1/dot.cuh:52-63 : out[i] = x[i]*y[i]; reduce(n,out) */
ORACLE_API float glass_dot_f32(uint32_t n, const float *x, const float *y, float *scratch)
{
    for (uint32_t i = 0; i < n; i++) scratch[i] = x[i] * y[i];
    return glass_reduce_f32(n, scratch);
}

/* GBD-PCG/include/utils.cuh:46-85 : one block row of the band matvec */
static void bdmv_row_f32(uint32_t n, uint32_t N, uint32_t b, const float *M, const float *x,
                         float *dst, int contract)
{
    const float *mat = M + (size_t)b * 3 * n * n;
    uint32_t c0, c1;             /* column range inside the [L|D|R] band row */
    const float *vec;            /* vec[c] pairs with column c0 + c */
    if (b == 0)          { c0 = n; c1 = 3 * n; vec = x; }                    /* utils.cuh:58-66 */
    else if (b == N - 1) { c0 = 0; c1 = 2 * n; vec = x + (size_t)(b - 1) * n; } /* :67-75 */
    else                 { c0 = 0; c1 = 3 * n; vec = x + (size_t)(b - 1) * n; } /* :76-84 */
    for (uint32_t r = 0; r < n; r++) {
        float val = 0.0f;
        for (uint32_t c = c0; c < c1; c++) {
            float m = mat[(size_t)n * c + r], v = vec[c - c0];
            if (contract) val = fmaf(m, v, val);
            else { volatile float prod = m * v; val = val + prod; }
        }
        dst[r] = val;
    }
}

ORACLE_API void bdmv_f32(uint32_t n, uint32_t N, const float *M, const float *x, float *y, int contract)
{
    for (uint32_t b = 0; b < N; b++) bdmv_row_f32(n, N, b, M, x, y + (size_t)b * n, contract);
}

/*
 * Full solve.  lambda is in/out.  r_out/p_out (nullable) receive what the reference leaves in
 * d_r/d_p.  eta_out (nullable) receives the last eta' evaluated.  Returns 0, or -1 on bad args.
 * pcg.cuh:98-217.
 */
ORACLE_API int pcg_oracle_f32(uint32_t n, uint32_t N, const float *S, const float *Pinv,
                              const float *gamma, float *lambda, uint32_t max_iter, float exit_tol,
                              int contract, uint32_t *iters_out, uint8_t *max_iter_exit_out,
                              float *r_out, float *p_out, float *eta_out)
{
    if (n == 0 || N < 2) return -1;
    const size_t len = (size_t)n * N;
    float *r = malloc(len * sizeof(float)), *p = malloc(len * sizeof(float));
    float *rt = malloc(len * sizeof(float)), *ups = malloc(len * sizeof(float));
    float *part = malloc((N > n ? N : n) * sizeof(float)), *scr = malloc(n * sizeof(float));
    uint32_t iter;
    uint8_t max_iter_exit = 1;
    float alpha, beta, eta, eta_new = 0.0f;

    /* r = gamma - S*lambda   (pcg.cuh:118-126) */
    bdmv_f32(n, N, S, lambda, r, contract);
    for (size_t i = 0; i < len; i++) r[i] = gamma[i] - r[i];
    /* r_tilde = Pinv*r ; p = r_tilde ; eta = r.r_tilde   (:130-149) */
    bdmv_f32(n, N, Pinv, r, rt, contract);
    memcpy(p, rt, len * sizeof(float));
    for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f32(n, r + (size_t)b * n, rt + (size_t)b * n, scr);
    eta = glass_reduce_f32(N, part);

    for (iter = 0; iter < max_iter; iter++) {
        /* upsilon = S*p ; alpha = eta / (p.upsilon)   (:156-169) */
        bdmv_f32(n, N, S, p, ups, contract);
        for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f32(n, p + (size_t)b * n, ups + (size_t)b * n, scr);
        alpha = eta / glass_reduce_f32(N, part);
        /* lambda += alpha p ; r -= alpha upsilon   (:172-176) */
        for (size_t i = 0; i < len; i++) {
            if (contract) {
                lambda[i] = fmaf(alpha, p[i], lambda[i]);
                r[i] = fmaf(-alpha, ups[i], r[i]);
            } else {
                volatile float a = alpha * p[i], c = alpha * ups[i];
                lambda[i] = lambda[i] + a;
                r[i] = r[i] - c;
            }
        }
        /* r_tilde = Pinv*r ; eta' = r.r_tilde   (:180-193) */
        bdmv_f32(n, N, Pinv, r, rt, contract);
        for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f32(n, r + (size_t)b * n, rt + (size_t)b * n, scr);
        eta_new = glass_reduce_f32(N, part);
        if (fabsf(eta_new) < exit_tol) { iter++; max_iter_exit = 0; break; }   /* :195 */
        beta = eta_new / eta;                                                   /* :199-200 */
        eta = eta_new;
        for (size_t i = 0; i < len; i++) {                                      /* :203-206 */
            if (contract) p[i] = fmaf(beta, p[i], rt[i]);
            else { volatile float a = beta * p[i]; p[i] = rt[i] + a; }
        }
    }
    if (iters_out) *iters_out = iter;                                           /* :212 */
    if (max_iter_exit_out) *max_iter_exit_out = max_iter_exit;
    if (r_out) memcpy(r_out, r, len * sizeof(float));
    if (p_out) memcpy(p_out, p, len * sizeof(float));
    if (eta_out) *eta_out = eta_new;
    free(r); free(p); free(rt); free(ups); free(part); free(scr);
    return 0;
}

/* ------------------------------------------------------------------ fp64 (USE_DOUBLES=1 instantiation) */

ORACLE_API double glass_reduce_f64(uint32_t n, double *x)
{
    uint32_t size_left = n;
    while (size_left > 3) {
        uint32_t odd = size_left % 2;
        size_left = (size_left - odd) / 2;
        for (uint32_t i = 0; i < size_left; i++) x[i] += x[i + size_left];
        if (odd) x[0] += x[2 * size_left];
    }
    for (uint32_t i = 1; i < size_left; i++) x[0] += x[i];
    return x[0];
}

ORACLE_API double glass_dot_f64(uint32_t n, const double *x, const double *y, double *scratch)
{
    for (uint32_t i = 0; i < n; i++) scratch[i] = x[i] * y[i];
    return glass_reduce_f64(n, scratch);
}

ORACLE_API void bdmv_f64(uint32_t n, uint32_t N, const double *M, const double *x, double *y, int contract)
{
    for (uint32_t b = 0; b < N; b++) {
        const double *mat = M + (size_t)b * 3 * n * n;
        uint32_t c0, c1;
        const double *vec;
        if (b == 0)          { c0 = n; c1 = 3 * n; vec = x; }
        else if (b == N - 1) { c0 = 0; c1 = 2 * n; vec = x + (size_t)(b - 1) * n; }
        else                 { c0 = 0; c1 = 3 * n; vec = x + (size_t)(b - 1) * n; }
        for (uint32_t r = 0; r < n; r++) {
            double val = 0.0;
            for (uint32_t c = c0; c < c1; c++) {
                double m = mat[(size_t)n * c + r], v = vec[c - c0];
                if (contract) val = fma(m, v, val);
                else { volatile double prod = m * v; val = val + prod; }
            }
            y[(size_t)b * n + r] = val;
        }
    }
}

ORACLE_API int pcg_oracle_f64(uint32_t n, uint32_t N, const double *S, const double *Pinv,
                              const double *gamma, double *lambda, uint32_t max_iter, double exit_tol,
                              int contract, uint32_t *iters_out, uint8_t *max_iter_exit_out,
                              double *r_out, double *p_out, double *eta_out)
{
    if (n == 0 || N < 2) return -1;
    const size_t len = (size_t)n * N;
    double *r = malloc(len * sizeof(double)), *p = malloc(len * sizeof(double));
    double *rt = malloc(len * sizeof(double)), *ups = malloc(len * sizeof(double));
    double *part = malloc((N > n ? N : n) * sizeof(double)), *scr = malloc(n * sizeof(double));
    uint32_t iter;
    uint8_t max_iter_exit = 1;
    double alpha, beta, eta, eta_new = 0.0;

    bdmv_f64(n, N, S, lambda, r, contract);
    for (size_t i = 0; i < len; i++) r[i] = gamma[i] - r[i];
    bdmv_f64(n, N, Pinv, r, rt, contract);
    memcpy(p, rt, len * sizeof(double));
    for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f64(n, r + (size_t)b * n, rt + (size_t)b * n, scr);
    eta = glass_reduce_f64(N, part);

    for (iter = 0; iter < max_iter; iter++) {
        bdmv_f64(n, N, S, p, ups, contract);
        for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f64(n, p + (size_t)b * n, ups + (size_t)b * n, scr);
        alpha = eta / glass_reduce_f64(N, part);
        for (size_t i = 0; i < len; i++) {
            if (contract) {
                lambda[i] = fma(alpha, p[i], lambda[i]);
                r[i] = fma(-alpha, ups[i], r[i]);
            } else {
                volatile double a = alpha * p[i], c = alpha * ups[i];
                lambda[i] = lambda[i] + a;
                r[i] = r[i] - c;
            }
        }
        bdmv_f64(n, N, Pinv, r, rt, contract);
        for (uint32_t b = 0; b < N; b++) part[b] = glass_dot_f64(n, r + (size_t)b * n, rt + (size_t)b * n, scr);
        eta_new = glass_reduce_f64(N, part);
        if (fabs(eta_new) < exit_tol) { iter++; max_iter_exit = 0; break; }
        beta = eta_new / eta;
        eta = eta_new;
        for (size_t i = 0; i < len; i++) {
            if (contract) p[i] = fma(beta, p[i], rt[i]);
            else { volatile double a = beta * p[i]; p[i] = rt[i] + a; }
        }
    }
    if (iters_out) *iters_out = iter;
    if (max_iter_exit_out) *max_iter_exit_out = max_iter_exit;
    if (r_out) memcpy(r_out, r, len * sizeof(double));
    if (p_out) memcpy(p_out, p, len * sizeof(double));
    if (eta_out) *eta_out = eta_new;
    free(r); free(p); free(rt); free(ups); free(part); free(scr);
    return 0;
}

/* ------------------------------------------------------------------ batched driver (threads = caller's choice) */

/* Solve `batch` independent systems back to back; used by the CPU baseline timing ("port" kind). */
ORACLE_API int pcg_oracle_batched_f32(uint32_t n, uint32_t N, uint32_t batch, const float *S,
                                      const float *Pinv, const float *gamma, float *lambda,
                                      uint32_t max_iter, float exit_tol, int contract,
                                      uint32_t *iters_out, uint8_t *max_iter_exit_out)
{
    const size_t ms = (size_t)3 * n * n * N, vs = (size_t)n * N;
    for (uint32_t i = 0; i < batch; i++) {
        int rc = pcg_oracle_f32(n, N, S + i * ms, Pinv + i * ms, gamma + i * vs, lambda + i * vs,
                                max_iter, exit_tol, contract, iters_out ? iters_out + i : NULL,
                                max_iter_exit_out ? max_iter_exit_out + i : NULL, NULL, NULL, NULL);
        if (rc) return rc;
    }
    return 0;
}
