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
 *   band row    six half-tile chains (packed FFMA2: columns 0 .. n/2-1 in the low half, n/2 .. n-1 in the high half;
 *               first term a product, then one FMA per column, ascending), combined as
 *               ((L.lo + D.lo) + R.lo) + ((L.hi + D.hi) + R.hi); pad tiles and rows outside the system
 *               contribute exact zeros
 *   dots        per CTA of R = N/C own knot rows laid out in groups of G lanes (G = 16, lanes n .. 15 idle, or G = n,
 *               packed): the G R per-thread products are added by LN lanes (8, or 16 when G R >= 384), lane l taking
 *               the products {2 LN m + 2l, 2 LN m + 2l + 1} and then a balanced tree over m, then an XOR butterfly
 *               LN/2 .. 1 over the LN lanes; the C CTA partials are summed in a balanced tree in ascending CTA order
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

/* one band row (knot row b, element j) of M times x; x is the global vector, rows outside [0,N) read as zero.
 * Per tile two half chains (columns 0 .. n/2-1 and n/2 .. n-1: the two halves of the kernel's packed FFMA2), each a product
 * followed by one FMA per column; combined as ((L.lo + D.lo) + R.lo) + ((L.hi + D.hi) + R.hi). */
static float chain_pairs(uint32_t n, uint32_t N, const float *M, const float *x, int b, uint32_t j)
{
    const uint32_t H = n / 2;
    float lo[3], hi[3];
    for (int blk = 0; blk < 3; blk++) {
        const int kb = b - 1 + blk;                 /* knot row the tile multiplies */
        const int zero = kb < 0 || kb >= (int)N;    /* pad tile / missing neighbour: the kernel multiplies zeros by zeros */
        const float *tile = M + ((size_t)b * 3 + blk) * n * n;
        float sl = 0.0f, sh = 0.0f;
        for (uint32_t c = 0; c < H; c++) {
            const float ml = zero ? 0.0f : tile[(size_t)c * n + j], vl = zero ? 0.0f : x[(size_t)kb * n + c];
            const float mh = zero ? 0.0f : tile[(size_t)(c + H) * n + j], vh = zero ? 0.0f : x[(size_t)kb * n + c + H];
            sl = c == 0 ? ml * vl : fmaf(ml, vl, sl);
            sh = c == 0 ? mh * vh : fmaf(mh, vh, sh);
        }
        lo[blk] = sl;
        hi[blk] = sh;
    }
    const float a = (lo[0] + lo[1]) + lo[2], c = (hi[0] + hi[1]) + hi[2];
    return a + c;
}

static void band(uint32_t n, uint32_t N, const float *M, const float *x, float *y)
{
    for (uint32_t b = 0; b < N; b++)
        for (uint32_t j = 0; j < n; j++) y[(size_t)b * n + j] = chain_pairs(n, N, M, x, (int)b, j);
}

/* balanced tree: (v0 + v1), (v2 + v3), ...; an odd element moves up unchanged */
static float tree_sum(float *v, uint32_t cnt)
{
    for (; cnt > 1; cnt = (cnt + 1) / 2) {
        for (uint32_t i = 0; i < cnt / 2; i++) v[i] = v[2 * i] + v[2 * i + 1];
        if (cnt & 1u) v[cnt / 2] = v[cnt - 1];
    }
    return v[0];
}

/* the batch kernel's reduction (gbd_cluster_pcg_fastb.cuh): four lanes per knot row, lane q holding the elements q, q + n/2, q + 4,
 * q + 4 + n/2 (those that exist); a lane adds its products as (p0 + p1) + (p2 + p3), missing ones as +0; a warp = eight knot rows
 * = 32 lanes adds them in an XOR butterfly 1, 2, 4, 8, 16 (a balanced tree in lane order); the R/8 warp sums of a CTA and then the
 * C CTA sums are added in balanced trees in ascending order */
static float dot_fast_b(uint32_t n, uint32_t N, uint32_t C, const float *a, const float *b)
{
    const uint32_t R = N / C, H = n / 2, W = R / 8;
    float part[16];
    for (uint32_t cr = 0; cr < C; cr++) {
        float wsum[128];
        for (uint32_t w = 0; w < W; w++) {
            float lanev[32];
            for (uint32_t l = 0; l < 32; l++) {
                const uint32_t g = w * 8 + l / 4, q = l % 4;
                const uint32_t e[4] = {q, q + H, q + 4, q + 4 + H};
                const int ok[4] = {q < H, q < H, q + 4 < H, q + 4 < H};
                float p[4];
                for (int k = 0; k < 4; k++) {
                    const size_t i = ((size_t)cr * R + g) * n + e[k];
                    p[k] = ok[k] ? a[i] * b[i] : 0.0f;
                }
                lanev[l] = (p[0] + p[1]) + (p[2] + p[3]);
            }
            wsum[w] = tree_sum(lanev, 32);
        }
        part[cr] = tree_sum(wsum, W);
    }
    return tree_sum(part, C);
}

/* the grid kernel's reduction (gbd_grid_pcg_fast.cuh; n a multiple of 32, C = number of CTAs, R = N/C knot rows each): one thread per
 * element in (row, element) order; every warp of 32 consecutive threads adds its products in an XOR butterfly (a balanced tree in
 * lane order); the R n / 32 warp sums of a CTA, and then the C CTA sums, are added in balanced trees in ascending order */
static float dot_fast_grid(uint32_t n, uint32_t N, uint32_t C, const float *a, const float *b)
{
    const uint32_t R = N / C, NW = R * n / 32;
    float *part = (float *)malloc(C * sizeof(float));
    for (uint32_t cr = 0; cr < C; cr++) {
        float wsum[64];
        for (uint32_t w = 0; w < NW; w++) {
            float lanev[32];
            for (uint32_t l = 0; l < 32; l++) {
                const size_t i = (size_t)cr * R * n + (size_t)w * 32 + l;
                lanev[l] = a[i] * b[i];
            }
            wsum[w] = tree_sum(lanev, 32);
        }
        part[cr] = tree_sum(wsum, NW);
    }
    const float tot = tree_sum(part, C);
    free(part);
    return tot;
}

/* the kernels' reduction of per-element products a[i]*b[i] (i over N*n) for cluster size C and G lanes per knot row
 * (G = 0: the batch kernel's order, G = 1: the grid kernel's order, above) */
static float dot_fast(uint32_t n, uint32_t N, uint32_t C, uint32_t G, const float *a, const float *b)
{
    if (G == 0) return dot_fast_b(n, N, C, a, b);
    if (G == 1) return dot_fast_grid(n, N, C, a, b);
    const uint32_t R = N / C, NOWN = R * G, LN = NOWN >= 384 ? 16 : 8, PPL = NOWN / (2 * LN);
    float part[16];
    float *prod = (float *)malloc(NOWN * sizeof(float));
    float *v = (float *)malloc(PPL * sizeof(float));
    for (uint32_t cr = 0; cr < C; cr++) {
        /* thread t = G g + j of the own-row warps parks its product; idle lanes j >= n hold zeros */
        for (uint32_t t = 0; t < NOWN; t++) {
            const uint32_t g = t / G, j = t % G;
            const size_t i = ((size_t)cr * R + g) * n + j;
            prod[t] = j < n ? a[i] * b[i] : 0.0f;
        }
        /* LN lanes: lane l adds the products {2 LN m + 2 l, 2 LN m + 2 l + 1}, m < PPL, then a balanced tree over m */
        float lanev[16];
        for (uint32_t l = 0; l < LN; l++) {
            for (uint32_t m = 0; m < PPL; m++) v[m] = prod[2 * LN * m + 2 * l] + prod[2 * LN * m + 2 * l + 1];
            lanev[l] = tree_sum(v, PPL);
        }
        for (uint32_t s = LN / 2; s >= 1; s >>= 1) {
            float nv[16];
            for (uint32_t l = 0; l < LN; l++) nv[l] = lanev[l] + lanev[l ^ s];
            memcpy(lanev, nv, sizeof nv);
        }
        part[cr] = lanev[0];
    }
    free(v);
    free(prod);
    return tree_sum(part, C);       /* every lane of the gathering warp adds the C pairs in the same balanced tree */
}

/*
 * Returns 0, or -1 on bad arguments.  lambda is in/out; r_out / p_out (nullable) receive the final residual and
 * direction; eta_out (nullable) the last gamma = r.Pinv r.
 */
ORACLE_API int pcg_fast_oracle_g_f32(uint32_t n, uint32_t N, uint32_t C, uint32_t G, const float *S, const float *Pinv, const float *gamma,
                                     float *lambda, uint32_t max_iter, float exit_tol, uint32_t *iters_out,
                                     uint8_t *max_iter_exit_out, float *r_out, float *p_out, float *eta_out)
{
    if (!S || !Pinv || !gamma || !lambda || n < 2 || C < 1 || N % C || N / C < 2 || n % 2) return -1;
    if (G == 1) {                                      /* grid kernel: whole warps per knot row, power-of-two CTA count */
        if (n % 32 || n > 64 || (C & (C - 1)) || (N / C) * n / 32 > 64) return -1;
    } else {
        if (n > 16 || C > 16) return -1;
        if ((N / C) % 2 || (G != 16 && G != n && G != 0) || (G && ((N / C) * G) % 16) || (!G && (N / C) % 8)) return -1;
    }
    const size_t len = (size_t)n * N;
    float *buf = (float *)calloc(7 * len, sizeof(float));
    if (!buf) return -1;
    float *r = buf, *u = buf + len, *w = buf + 2 * len, *p = buf + 3 * len, *s = buf + 4 * len, *t = buf + 5 * len;

    band(n, N, S, lambda, t);
    for (size_t i = 0; i < len; i++) r[i] = gamma[i] - t[i];
    band(n, N, Pinv, r, u);
    band(n, N, S, u, w);
    float gam = dot_fast(n, N, C, G, r, u), del = dot_fast(n, N, C, G, w, u);
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
        const float gam_new = dot_fast(n, N, C, G, r, u), del_new = dot_fast(n, N, C, G, w, u);
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

/* the 16-lanes-per-row kernels (single solves) */
ORACLE_API int pcg_fast_oracle_f32(uint32_t n, uint32_t N, uint32_t C, const float *S, const float *Pinv, const float *gamma,
                                   float *lambda, uint32_t max_iter, float exit_tol, uint32_t *iters_out,
                                   uint8_t *max_iter_exit_out, float *r_out, float *p_out, float *eta_out)
{
    return pcg_fast_oracle_g_f32(n, N, C, 16, S, Pinv, gamma, lambda, max_iter, exit_tol, iters_out, max_iter_exit_out, r_out, p_out, eta_out);
}
