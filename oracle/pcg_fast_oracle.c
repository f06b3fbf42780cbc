/*
 * pcg_fast_oracle.c -- CPU restatement of the TOLERANCE-PARITY ("fast") GBD-PCG kernels, operation by
 * operation in THEIR floating-point order.  TEST INFRASTRUCTURE ONLY (same rules as pcg_oracle.c: never
 * linked into or executed from the product).
 *
 * Purpose: the fast kernels (include/gbd/gbd_cluster_pcg_fast.cuh) give up the reference's summation
 * order, so against the reference (GBD-PCG/include/pcg.cuh:54-218, restated in pcg_oracle.c) they are
 * held to the tolerance policy of SURVEY.md 8(c)(ii).  A tolerance can hide a real bug, so the kernels
 * are ALSO checked bit for bit against this file, which restates exactly what they compute:
 *
 *   algorithm   Chronopoulos-Gear form of preconditioned CG (same iterates as pcg.cuh:154-208 in exact
 *               arithmetic, same exit rule |r.Pinv r| < tol after every update, same iteration count and
 *               max_iter_exit meaning, pcg.cuh:195,212)
 *   band row    three per-tile chains (first term a product, then one FMA per column, ascending), then
 *               (left + diag) + right; pad tiles and rows outside the system contribute exact zeros
 *   dots        per CTA of R = N/C own knot rows laid out in 16-lane groups (group 0 = the halo row a-1,
 *               contributing zeros): the NT per-thread products are added by eight lanes, lane l taking the
 *               pairs {16m + 2l, 16m + 2l + 1} in a balanced tree, then an XOR butterfly 4,2,1 over the eight;
 *               the C CTA partials are summed 4 per lane in ascending order and then by an XOR butterfly
 *               over C/4 lanes
 *   scalars     correctly rounded reciprocals (1.0f/x) times products, FMAs as written
 *
 * Layout as in pcg_oracle.c: S, Pinv = [N][3][n][n], column-major tiles, tiles (0,left), (N-1,right) unused.
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* one band row (knot row b, element j) of M times x; x is the global vector, rows outside [0,N) read as zero */
static float chain3(uint32_t n, uint32_t N, const float *M, const float *x, int b, uint32_t j)
{
    float acc[3];
    for (int blk = 0; blk < 3; blk++) {
        const int kb = b - 1 + blk;                 /* knot row the tile multiplies */
        const int zero = kb < 0 || kb >= (int)N;    /* pad tile / missing neighbour: the kernel multiplies zeros by zeros */
        const float *tile = M + ((size_t)b * 3 + blk) * n * n;
        float s = 0.0f;
        for (uint32_t c = 0; c < n; c++) {
            const float m = zero ? 0.0f : tile[(size_t)c * n + j];
            const float v = zero ? 0.0f : x[(size_t)kb * n + c];
            s = c == 0 ? m * v : fmaf(m, v, s);
        }
        acc[blk] = s;
    }
    return (acc[0] + acc[1]) + acc[2];
}

static void band(uint32_t n, uint32_t N, const float *M, const float *x, float *y)
{
    for (uint32_t b = 0; b < N; b++)
        for (uint32_t j = 0; j < n; j++) y[(size_t)b * n + j] = chain3(n, N, M, x, (int)b, j);
}

/* the kernels' reduction of per-element products a[i]*b[i] (i over N*n) for cluster size C */
static float dot_fast(uint32_t n, uint32_t N, uint32_t C, const float *a, const float *b)
{
    const uint32_t R = N / C, NG = R + 2, NT = (NG * 16 + 31) / 32 * 32, PPL = NT / 16;
    float part[16];
    float *prod = (float *)malloc(NT * sizeof(float));
    for (uint32_t cr = 0; cr < C; cr++) {
        /* thread t = 16 g + j parks its product; group 0 and R+1 (halo rows), lanes j >= n and padding threads park zeros */
        for (uint32_t t = 0; t < NT; t++) {
            const uint32_t g = t / 16, j = t % 16;
            const int own = g >= 1 && g <= R && j < n;
            const size_t i = ((size_t)cr * R + (g - 1)) * n + j;
            prod[t] = own ? a[i] * b[i] : 0.0f;
        }
        /* eight lanes: lane l adds the pairs {16 m + 2 l, 16 m + 2 l + 1}, m < PPL, in a balanced tree */
        float lanev[8];
        for (uint32_t l = 0; l < 8; l++) {
            float v[64];
            for (uint32_t m = 0; m < PPL; m++) v[m] = prod[16 * m + 2 * l] + prod[16 * m + 2 * l + 1];
            for (uint32_t cnt = PPL; cnt > 1; cnt = (cnt + 1) / 2) {
                for (uint32_t i = 0; i < cnt / 2; i++) v[i] = v[2 * i] + v[2 * i + 1];
                if (cnt & 1u) v[cnt / 2] = v[cnt - 1];
            }
            lanev[l] = v[0];
        }
        for (uint32_t s = 4; s >= 1; s >>= 1) {
            float nv[8];
            for (uint32_t l = 0; l < 8; l++) nv[l] = lanev[l] + lanev[l ^ s];
            memcpy(lanev, nv, sizeof nv);
        }
        part[cr] = lanev[0];
    }
    free(prod);
    const uint32_t PERQ = C < 4 ? C : 4, LQ = C / PERQ;
    float lanev[4];
    for (uint32_t l = 0; l < LQ; l++) {
        float s = part[l * PERQ];
        for (uint32_t m = 1; m < PERQ; m++) s = s + part[l * PERQ + m];
        lanev[l] = s;
    }
    for (uint32_t s = LQ / 2; s >= 1; s >>= 1) {
        float nv[4];
        for (uint32_t l = 0; l < LQ; l++) nv[l] = lanev[l] + lanev[l ^ s];
        memcpy(lanev, nv, sizeof nv);
    }
    return lanev[0];
}

/*
 * Returns 0, or -1 on bad arguments.  lambda is in/out; r_out / p_out (nullable) receive the final residual and
 * direction; eta_out (nullable) the last gamma = r.Pinv r.
 */
ORACLE_API int pcg_fast_oracle_f32(uint32_t n, uint32_t N, uint32_t C, const float *S, const float *Pinv, const float *gamma,
                                   float *lambda, uint32_t max_iter, float exit_tol, uint32_t *iters_out,
                                   uint8_t *max_iter_exit_out, float *r_out, float *p_out, float *eta_out)
{
    if (!S || !Pinv || !gamma || !lambda || n < 2 || n > 16 || C < 1 || C > 16 || N % C || N / C < 2) return -1;
    if (!(C < 4 || C % 4 == 0)) return -1;
    const size_t len = (size_t)n * N;
    float *buf = (float *)calloc(7 * len, sizeof(float));
    if (!buf) return -1;
    float *r = buf, *u = buf + len, *w = buf + 2 * len, *p = buf + 3 * len, *s = buf + 4 * len, *t = buf + 5 * len;

    band(n, N, S, lambda, t);
    for (size_t i = 0; i < len; i++) r[i] = gamma[i] - t[i];
    band(n, N, Pinv, r, u);
    band(n, N, S, u, w);
    float gam = dot_fast(n, N, C, r, u), del = dot_fast(n, N, C, w, u);
    float alpha = gam * (1.0f / del), beta = 0.0f;
    float rgam = 1.0f / gam, q = del * rgam;
    uint32_t iter = 0;
    uint8_t flag = 1;
    for (; iter < max_iter; iter++) {
        for (size_t i = 0; i < len; i++) {
            p[i] = fmaf(beta, p[i], u[i]);
            s[i] = fmaf(beta, s[i], w[i]);
            lambda[i] = fmaf(alpha, p[i], lambda[i]);
            r[i] = fmaf(-alpha, s[i], r[i]);
        }
        band(n, N, Pinv, r, u);
        band(n, N, S, u, w);
        const float gam_new = dot_fast(n, N, C, r, u), del_new = dot_fast(n, N, C, w, u);
        gam = gam_new;
        if (fabsf(gam_new) < exit_tol) { iter++; flag = 0; break; }
        beta = gam_new * rgam;
        const float bg = beta * gam_new;
        const float den = fmaf(-bg, q, del_new);
        alpha = gam_new * (1.0f / den);
        rgam = 1.0f / gam_new;
        q = den * rgam;
    }
    if (iters_out) *iters_out = iter;
    if (max_iter_exit_out) *max_iter_exit_out = flag;
    if (r_out) memcpy(r_out, r, len * sizeof(float));
    if (p_out) memcpy(p_out, p, len * sizeof(float));
    if (eta_out) *eta_out = gam;
    free(buf);
    return 0;
}
