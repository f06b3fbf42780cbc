/*
 * qdldl_driver.c -- host driver around the REFERENCE's own QDLDL (qdldl v0.1.7, compiled from
 * /root/reference/qdldl/src/qdldl.c by oracle/Makefile into oracle/_ref/).  TEST/BASELINE
 * INFRASTRUCTURE ONLY: used by tests/ and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Restates the glue the reference wraps around QDLDL:
 *   include/utils/csr.cuh:40-74     prep_csr            -> qdldl_ref_pattern()
 *   include/utils/csr.cuh:10-36     store_block_csr_lowertri (called for the left and diagonal
 *                                   tiles at include/qdldl/linsys_setup.cuh:106,315,321)
 *                                                        -> qdldl_ref_values()
 *   include/qdldl/sqp.cuh:148-198   workspace + QDLDL_etree once            -> qdldl_ref_create()
 *   include/qdldl/sqp.cuh:22-49     qdldl_solve_schur = factor + copy + solve -> qdldl_ref_solve()
 * The matrix is the stored (negated) Schur complement in upper-triangular CSC: column j=(b,row)
 * holds rows (b-1)n..(b-1)n+n-1 (left tile, row `row` of L_b) then rows bn..bn+row (diag tile).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "qdldl.h"

#define API __attribute__((visibility("default")))

typedef struct {
    QDLDL_int n, N, An, nnz, sumLnz;
    QDLDL_int *col_ptr, *row_ind, *etree, *Lnz, *Lp, *Li, *iwork;
    QDLDL_float *Lx, *D, *Dinv, *fwork;
    QDLDL_bool *bwork;
} qdldl_ref_ws;

API int qdldl_ref_nnz(int n, int N) { return (N - 1) * n * n + N * ((n + 1) * n / 2); }   /* sqp.cuh:148 */
API int qdldl_ref_float_bytes(void) { return (int)sizeof(QDLDL_float); }

/* csr.cuh:40-74 */
API void qdldl_ref_pattern(int n, int N, QDLDL_int *col_ptr, QDLDL_int *row_ind)
{
    const int brow_val_ct = n * n + ((n + 1) * n) / 2;
    col_ptr[0] = 0;
    for (int b = 0; b < N; b++)
        for (int row = 0; row < n; row++) {
            int tri = ((row + 1) * row) / 2;
            int off = (b > 0) * ((n + 1) * n) / 2 + (b > 0) * (b - 1) * brow_val_ct + (b > 0) * row * n + tri;
            int len = (b > 0) * n + row + 1;
            col_ptr[b * n + row + 1] = off + len;
            for (int col = 0; col < len; col++) row_ind[off + col] = (b > 0) * (b - 1) * n + col;
        }
}

/* csr.cuh:10-36 applied to tiles 0 (col1=false) and 1 (col1=true) of every block row of S */
API void qdldl_ref_values(int n, int N, const float *S, QDLDL_float *val)
{
    const int brow_val_ct = n * n + ((n + 1) * n) / 2;
    for (int b = 0; b < N; b++) {
        const float *L = S + (size_t)b * 3 * n * n, *Dg = L + n * n;
        for (int row = 0; row < n; row++) {
            int tri = ((row + 1) * row) / 2;
            int off = (b > 0) * ((n + 1) * n) / 2 + (b > 0) * (b - 1) * brow_val_ct + (b > 0) * row * n + tri;
            if (b > 0)
                for (int col = 0; col < n; col++) val[off + col] = (QDLDL_float)L[row + col * n];
            for (int col = 0; col <= row; col++) val[off + (b > 0) * n + col] = (QDLDL_float)Dg[row + col * n];
        }
    }
}

API void qdldl_ref_destroy(qdldl_ref_ws *w)
{
    if (!w) return;
    free(w->col_ptr); free(w->row_ind); free(w->etree); free(w->Lnz); free(w->Lp); free(w->Li);
    free(w->iwork); free(w->Lx); free(w->D); free(w->Dinv); free(w->fwork); free(w->bwork); free(w);
}

/* sqp.cuh:148-198 */
API qdldl_ref_ws *qdldl_ref_create(int n, int N)
{
    qdldl_ref_ws *w = calloc(1, sizeof(*w));
    w->n = n; w->N = N; w->An = n * N; w->nnz = qdldl_ref_nnz(n, N);
    w->col_ptr = malloc(sizeof(QDLDL_int) * (w->An + 1));
    w->row_ind = malloc(sizeof(QDLDL_int) * w->nnz);
    w->etree = malloc(sizeof(QDLDL_int) * w->An);
    w->Lnz = malloc(sizeof(QDLDL_int) * w->An);
    w->Lp = malloc(sizeof(QDLDL_int) * (w->An + 1));
    w->D = malloc(sizeof(QDLDL_float) * w->An);
    w->Dinv = malloc(sizeof(QDLDL_float) * w->An);
    w->iwork = malloc(sizeof(QDLDL_int) * 3 * w->An);
    w->bwork = malloc(sizeof(QDLDL_bool) * w->An);
    w->fwork = malloc(sizeof(QDLDL_float) * w->An);
    qdldl_ref_pattern(n, N, w->col_ptr, w->row_ind);
    w->sumLnz = QDLDL_etree(w->An, w->col_ptr, w->row_ind, w->iwork, w->Lnz, w->etree);
    if (w->sumLnz < 0) { qdldl_ref_destroy(w); return NULL; }
    w->Li = malloc(sizeof(QDLDL_int) * w->sumLnz);
    w->Lx = malloc(sizeof(QDLDL_float) * w->sumLnz);
    return w;
}

API int qdldl_ref_sum_lnz(const qdldl_ref_ws *w) { return (int)w->sumLnz; }

/* sqp.cuh:22-49 ; returns the number of positive pivots (>=0) or -1 on factorisation failure */
API int qdldl_ref_solve(qdldl_ref_ws *w, const QDLDL_float *val, const QDLDL_float *b, QDLDL_float *x)
{
    QDLDL_int rc = QDLDL_factor(w->An, w->col_ptr, w->row_ind, val, w->Lp, w->Li, w->Lx, w->D, w->Dinv,
                                w->Lnz, w->etree, w->bwork, w->iwork, w->fwork);
    for (QDLDL_int i = 0; i < w->An; i++) x[i] = b[i];
    QDLDL_solve(w->An, w->Lp, w->Li, w->Lx, w->Dinv, x);
    return (int)rc;
}

/* ---- batched: one system per thread over `nthreads` host threads (QDLDL itself is sequential) */
typedef struct {
    int n, N, tid, nthreads, batch, reps, rc;
    const QDLDL_float *val, *b;
    QDLDL_float *x;
} job_t;

static void *worker(void *arg)
{
    job_t *j = arg;
    qdldl_ref_ws *w = qdldl_ref_create(j->n, j->N);
    if (!w) { j->rc = -1; return NULL; }
    const size_t nnz = (size_t)w->nnz, An = (size_t)w->An;
    for (int rep = 0; rep < j->reps; rep++)
        for (int i = j->tid; i < j->batch; i += j->nthreads)
            if (qdldl_ref_solve(w, j->val + i * nnz, j->b + i * An, j->x + i * An) < 0) j->rc = -2;
    qdldl_ref_destroy(w);
    return NULL;
}

/* Returns elapsed wall seconds for reps x batch factor+solve pairs (workspace creation and the
 * one-off etree are inside each thread but outside nothing: they are included, once per thread,
 * which is negligible for reps*batch/nthreads >= 8).  <0 on failure. */
API double qdldl_ref_solve_batched(int n, int N, int batch, int reps, int nthreads,
                                   const QDLDL_float *val, const QDLDL_float *b, QDLDL_float *x)
{
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = malloc(sizeof(job_t) * nthreads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){n, N, t, nthreads, batch, reps, 0, val, b, x};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int bad = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); bad |= jobs[t].rc; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs);
    if (bad) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
