/*
 * gbd_pcg.h -- flat C ABI of the B200-native GBD-PCG solver (libgbdpcg.so).
 *
 * This is the drop-in boundary for the one hot path of A2R-Lab/MPCGPU: the block-tridiagonal
 * preconditioned-conjugate-gradient solve of  S * lambda = gamma  with preconditioner Pinv.
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * tree).  Plain pointers and sizes only; no C++ or torch types.  All functions return
 * GBD_PCG_OK (0) or a negative gbd_pcg_status; none of them ever calls exit() (the reference
 * aborts through gpuErrchk, GBD-PCG/include/gpuassert.cuh:5-14).
 *
 * Data layout (identical to the reference, GBD-PCG/include/utils.cuh:58-84,98-161):
 *   S, Pinv : [N][3][n][n]   tile t in {0:left, 1:diag, 2:right} of block row b at
 *                            b*3n^2 + t*n^2, COLUMN-major inside a tile (elem(r,c) at c*n + r).
 *                            Tiles (0,0) and (N-1,2) are padding: never read, may be uninitialised.
 *   gamma, lambda, r, p : [N*n]
 * Batched variants prepend a [batch] dimension to every array.
 *
 * Numerics (gbd_pcg_set_numerics): two kernel families solve the same problem behind the same entry points.
 *   GBD_PCG_NUMERICS_FAST (default)  tolerance parity with the reference kernel pcg<T,n,N>: same algorithm, exit rule and
 *       iteration-count meaning, but its own summation order and a single-reduction form of the CG recurrence
 *       (include/gbd/gbd_cluster_pcg_fast.cuh).  Stated tolerance (tests/test_gpu_fast.py): iteration
 *       count within +-2 of the reference kernel's, max_iter_exit identical unless within 2 of the cap,
 *       max|lambda - lambda_ref| / max|lambda_ref| <= 1e-3, fp64 relative residual <= 1.1 x the reference kernel's.
 *       Results are deterministic (no atomics: same input, same bits, whichever cluster solves a system).
 *   GBD_PCG_NUMERICS_BITEXACT  bit-identical to the reference kernel (same floating-point operation order: sequential
 *       FMA over the band row, GLASS halving trees for the dots, IEEE division); about 2x slower.
 *   Shapes without a fast kernel (fp64, N = 512) are solved by the bit-exact family under either setting.
 *
 * Process model: one CUDA device per process (the multi-GPU path is one process per GPU): kernel
 * attributes, cluster occupancy, work counters and packet workspaces are created once, for the device that
 * is current at the first compute call; a compute call made with another device current returns
 * GBD_PCG_ERR_DEVICE.  Entry points taking a stream may be called from several host threads on different
 * streams; gbd_pcg_linsys_f32 brackets the legacy default stream like the reference's stopwatch window and
 * is one caller at a time.
 */
#ifndef GBD_PCG_H
#define GBD_PCG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gbd_pcg_status {
    GBD_PCG_OK = 0,
    GBD_PCG_ERR_UNSUPPORTED = -1, /* (n, N, dtype) has no compiled kernel; see gbd_pcg_supported() */
    GBD_PCG_ERR_BADARG = -2,      /* null pointer, N < 2, batch == 0, ... */
    GBD_PCG_ERR_CUDA = -3,        /* a CUDA call failed; code in gbd_pcg_last_cuda_error() */
    GBD_PCG_ERR_NODEVICE = -4,    /* no CUDA device / driver: there is NO CPU fallback */
    GBD_PCG_ERR_DEVICE = -5       /* the current CUDA device is not the one this process first used the library on */
} gbd_pcg_status;

#define GBD_PCG_ABI_VERSION 2

int gbd_pcg_abi_version(void);
const char *gbd_pcg_strerror(int status);
int gbd_pcg_last_cuda_error(void); /* cudaError_t of the last failing CUDA call on this thread */

/* Which (state_size n, knot_points N) pairs are compiled in.  The reference fixes (n, N) at compile
 * time through the STATE_SIZE / KNOT_POINTS macros (GBD-PCG/include/constants.cuh:5-11,
 * interface.cuh:110); here the pairs are enumerated at run time. */
int gbd_pcg_supported(uint32_t n, uint32_t N, int is_f64);
int gbd_pcg_num_variants(void);
int gbd_pcg_variant_at(int i, uint32_t *n, uint32_t *N, uint32_t *cluster, int *mode, int *is_f64,
                       uint32_t *threads, size_t *smem_bytes);

/* Kernel family used when no tuning is set: GBD_PCG_NUMERICS_FAST (default; environment GBD_PCG_NUMERICS=exact selects the
 * other one at load time) or GBD_PCG_NUMERICS_BITEXACT.  Process-wide; takes effect for subsequent launches. */
#define GBD_PCG_NUMERICS_BITEXACT 0
#define GBD_PCG_NUMERICS_FAST 1
int gbd_pcg_set_numerics(int numerics);
int gbd_pcg_get_numerics(void);

/* The variant a launch of this shape would run right now (tuning, numerics and what the device can place taken into
 * account): cluster size (CTA count for the whole-GPU kernels), mode as in gbd_pcg_set_tuning, threads per CTA, dynamic
 * shared memory, and the kernel's name as profilers print it (without template arguments).  Needs a CUDA device. */
int gbd_pcg_resolved_variant(uint32_t n, uint32_t N, int is_f64, int batched, uint32_t *cluster, int *mode, uint32_t *threads,
                             size_t *smem_bytes, char *kernel_name, size_t kernel_name_len);

/* Tuning knob: pick the cluster size (CTAs per system) and kernel build used for (n, N).
 * mode: 0 = v1 kernel, tiles in shared memory; 1 = v1, tiles in registers; 2 = v2 kernel (st.async +
 * mbarrier signalling), 1 CTA/SM register budget; 3 = v2, 2 CTAs/SM budget; 4 = grid kernel (whole GPU on
 * one system, L2 packets); 5 = v3 kernel (two matrix rows per thread), 1 CTA/SM budget; 6 = v3, 2 CTAs/SM;
 * 7 = v4 kernel (self-validating {value, epoch} packets polled in shared memory), 1 CTA/SM; 8 = v4, 2 CTAs/SM;
 * 10, 14 = v4 / v2 timeline builds; 11, 12 = v5 kernel (v3's mapping + v4's packets).  Modes 0 .. 14 give bit-identical results.
 * 20 = fast cluster kernel (tolerance parity, see Numerics above), 21 = its 2 CTAs/SM build, 22 = its timeline build,
 * 24 = fast whole-GPU kernel, 26 = fast batched kernel.  A tuning overrides gbd_pcg_set_numerics for its shape.
 * cluster = 0 and mode = -1 restore the built-in default. */
int gbd_pcg_set_tuning(uint32_t n, uint32_t N, int is_f64, uint32_t cluster, int mode);

/*
 * One solve, device pointers, asynchronous on `stream` (a cudaStream_t, or NULL).
 * Replaces the kernel launch at include/pcg/sqp.cuh:129-151,230
 *   cudaLaunchCooperativeKernel(pcg<T,STATE_SIZE,KNOT_POINTS>, knot_points, PCG_NUM_THREADS, args, smem)
 * with the kernel's own 12-argument list (GBD-PCG/include/pcg.cuh:56-68), plus n, N and the stream.
 *   d_lambda  in: initial guess, out: solution
 *   d_r, d_p  caller-owned scratch of N*n; on return hold the final residual / direction, as the
 *             reference leaves them (either may be NULL)
 *   d_v_temp, d_eta_new_temp  caller-owned scratch of N in the reference; unused here, may be NULL
 *   d_iters   1 x uint32: iterations run (k+1 when iteration k passed the exit test, max_iter otherwise)
 *   d_max_iter_exit  1 byte (the reference's bool): 1 when the iteration cap was hit
 */
int gbd_pcg_solve_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_Pinv, const float *d_gamma,
                      float *d_lambda, float *d_r, float *d_p, float *d_v_temp, float *d_eta_new_temp,
                      uint32_t *d_iters, uint8_t *d_max_iter_exit, uint32_t max_iter, float exit_tol,
                      void *stream);
int gbd_pcg_solve_f64(uint32_t n, uint32_t N, const double *d_S, const double *d_Pinv, const double *d_gamma,
                      double *d_lambda, double *d_r, double *d_p, double *d_v_temp, double *d_eta_new_temp,
                      uint32_t *d_iters, uint8_t *d_max_iter_exit, uint32_t max_iter, double exit_tol,
                      void *stream);

/*
 * The reference's "SQP linsys" window, include/pcg/sqp.cuh:224-241: device sync, launch, blocking
 * read-back of the iteration count and the exit flag, device sync.  Returns the two host values and
 * (if elapsed_us != NULL) the wall time of the window measured the way the reference measures it
 * (clock_gettime(CLOCK_MONOTONIC) around it).
 */
int gbd_pcg_linsys_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_Pinv, const float *d_gamma,
                       float *d_lambda, float *d_r, float *d_p, uint32_t *d_iters, uint8_t *d_max_iter_exit,
                       uint32_t max_iter, float exit_tol, uint32_t *h_iters, uint8_t *h_max_iter_exit,
                       double *elapsed_us);

/*
 * Many independent systems in one launch (new capability; the reference has no batched path).
 * Arrays carry a leading [batch] dimension; d_iters / d_max_iter_exit have `batch` entries.
 * System i is solved exactly as gbd_pcg_solve_f32 would solve it alone (bit-identical).
 * d_r / d_p may be NULL.  With more systems than the GPU keeps clusters resident, clusters take the systems first come,
 * first served (a 4-byte device counter zeroed on `stream` ahead of the launch); the environment variable
 * GBD_PCG_STATIC_BATCH=1 restores a fixed stride.  Which cluster solves a system never changes its result.
 */
int gbd_pcg_solve_batched_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_Pinv,
                              const float *d_gamma, float *d_lambda, float *d_r, float *d_p,
                              uint32_t *d_iters, uint8_t *d_max_iter_exit, uint32_t max_iter, float exit_tol,
                              void *stream);

/*
 * Host-buffer path.  Replaces solvePCG<T>(h_S, h_gamma, h_lambda, stateSize, knotPoints, config)
 * (GBD-PCG/include/interface.cuh:24-89): copies S, Pinv, gamma, lambda to the device, solves,
 * copies lambda back.  Unlike the reference it (a) takes the preconditioner explicitly (the
 * reference leaves d_Pinv uninitialised, interface.cuh:45-46), (b) returns the real iteration
 * count and exit flag (the reference returns the constant 1, interface.cuh:88), and (c) keeps its
 * device buffers and stream in a reusable plan instead of cudaMalloc/cudaFree per call.
 * Host buffers may be pageable or pinned; `batch` systems are solved per call.
 * Pinned / registered buffers are read and written in place by the kernel (zero-copy); the plan remembers the device alias
 * of every such buffer it has seen.  A buffer that is unregistered or freed while the plan is alive must be forgotten with
 * gbd_pcg_plan_invalidate before its address can be reused for another allocation.
 */
typedef struct gbd_pcg_plan gbd_pcg_plan;
int gbd_pcg_plan_create(uint32_t n, uint32_t N, uint32_t batch, int is_f64, gbd_pcg_plan **out);
int gbd_pcg_plan_destroy(gbd_pcg_plan *plan);
int gbd_pcg_plan_invalidate(gbd_pcg_plan *plan);   /* forget the cached device aliases of caller buffers */
int gbd_pcg_plan_solve_host_f32(gbd_pcg_plan *plan, const float *h_S, const float *h_Pinv, const float *h_gamma,
                                float *h_lambda, uint32_t max_iter, float exit_tol, uint32_t *h_iters,
                                uint8_t *h_max_iter_exit);
int gbd_pcg_plan_solve_host_f64(gbd_pcg_plan *plan, const double *h_S, const double *h_Pinv,
                                const double *h_gamma, double *h_lambda, uint32_t max_iter, double exit_tol,
                                uint32_t *h_iters, uint8_t *h_max_iter_exit);

/*
 * The two steps either side of the solve inside one SQP iteration, same layouts and
 * bit-identical results:
 *   gbd_form_schur_system_f32  replaces form_schur_system<float>(state_size, control_size, knot_points, d_G_dense,
 *       d_C_dense, d_g, d_c, d_S, d_Pinv, d_gamma, rho)  (include/pcg/linsys_setup.cuh:621-657, called at
 *       include/pcg/sqp.cuh:207): KKT blocks -> S, Pinv, gamma; d_G_dense is overwritten with the block inverses.
 *       Two ordinary launches on `stream` (no cooperative launch, no co-residency requirement).
 *   gbd_compute_dz_f32         replaces compute_dz<float>(state_size, control_size, knot_points, d_G_dense, d_C_dense,
 *       d_g_val, d_lambda, d_dz)  (include/common/dz.cuh:125-136, called at include/pcg/sqp.cuh:250).
 * Sizes: G (n*n+m*m)*(N-1)+n*n, C (n*n+n*m)*(N-1), g and dz (n+m)*(N-1)+n, c n*N.  (n, m) must be a compiled pair
 * (gbd_schur_supported); GBD_PCG_ERR_UNSUPPORTED otherwise.
 */
int gbd_schur_supported(uint32_t n, uint32_t m);
/* How the assembly's first phase is mapped: 0 = one CTA per block row (lowest latency: the launch for one trajectory), 1 = one
 * warp per block row (highest throughput: no CTA barriers, the launch for batches), -1 (default) = by size: warp-per-row when
 * batch * N >= 4096 block rows.  Both run the same operations per element (bit-identical results).  Returns the previous mode. */
int gbd_schur_set_team(int mode);
int gbd_form_schur_system_f32(uint32_t n, uint32_t m, uint32_t N, float *d_G, const float *d_C, const float *d_g,
                              const float *d_c, float *d_S, float *d_Pinv, float *d_gamma, float rho, void *stream);
int gbd_compute_dz_f32(uint32_t n, uint32_t m, uint32_t N, const float *d_Ginv, const float *d_C, const float *d_g,
                       const float *d_lambda, float *d_dz, void *stream);

/*
 * One SQP linear-system step for `batch` trajectories: what include/pcg/sqp.cuh:207-258 does per SQP
 * iteration -- form_schur_system, the pcg<> launch with two blocking read-backs, compute_dz -- enqueued on `stream` as
 * assembly -> solve (warm-started from d_lambda, in/out) -> dz with no host round trip in between.  The plan owns S, Pinv,
 * gamma and the result slots (the reference cudaMallocs them on every sqpSolvePcg call, sqp.cuh:94-135).  All arrays carry a
 * leading [batch] dimension; d_G is overwritten with the block inverses as form_schur_system does.
 * gbd_step_results blocks on `stream` and returns the per-trajectory iteration counts and max_iter_exit flags of the last
 * run; gbd_step_device_flags is the device copy of the flags (what the multi-GPU driver all-gathers per outer step).
 */
typedef struct gbd_step_plan gbd_step_plan;
int gbd_step_plan_create(uint32_t n, uint32_t m, uint32_t N, uint32_t batch, gbd_step_plan **out);
int gbd_step_plan_destroy(gbd_step_plan *plan);
int gbd_step_run_f32(gbd_step_plan *plan, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                     float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream);
/* gbd_step_run_f32 with a direct fallback: trajectories whose PCG solve hit max_iter are re-solved by gbd_bcr_solve_flagged_f32
 * before dz is formed (their max_iter_exit flags stay set, so the caller still sees which ones they were). */
int gbd_step_run_fallback_f32(gbd_step_plan *plan, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                              float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream);
int gbd_step_results(gbd_step_plan *plan, uint32_t *h_iters, uint8_t *h_max_iter_exit, void *stream);
const uint8_t *gbd_step_device_flags(gbd_step_plan *plan);

/*
 * The QDLDL wire format of the band matrix (include/utils/csr.cuh:10-74): upper-triangular CSC of the
 * symmetric block-tridiagonal S, what the reference's CPU path (include/qdldl/sqp.cuh:148-198,268-273) factorises.
 * gbd_schur_csr_pattern_i32 replaces prep_csr (col_ptr: n*N+1 ints, row_ind: nnz ints); gbd_schur_csr_values_f32 packs the
 * left and diagonal tiles of d_S ([N][3][n][n], the pcg<> layout) as store_block_csr_lowertri does (nnz floats).
 */
uint32_t gbd_schur_csr_nnz(uint32_t n, uint32_t N);
int gbd_schur_csr_pattern_i32(uint32_t n, uint32_t N, int32_t *d_col_ptr, int32_t *d_row_ind, void *stream);
int gbd_schur_csr_values_f32(uint32_t n, uint32_t N, const float *d_S, float *d_val, void *stream);

/*
 * Direct solve of S lambda = gamma by block cyclic reduction (the GPU alternative to the reference's
 * CPU QDLDL path, include/qdldl/sqp.cuh:22-49, for systems on which PCG runs into its cap).  Same band layout as
 * gbd_pcg_solve_f32; no preconditioner, no initial guess, no iteration count.  A different algorithm from the reference's
 * two solvers: results agree with them to fp32 solver tolerance, not bit for bit.  Asynchronous on `stream`.
 */
int gbd_bcr_supported(uint32_t n, uint32_t N);
int gbd_bcr_solve_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_gamma, float *d_lambda, void *stream);
int gbd_bcr_solve_batched_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_gamma, float *d_lambda,
                              void *stream);
/* Same, but only for the systems i with d_only_if[i] != 0 -- e.g. the max_iter_exit flags of a batched PCG solve: the
 * trajectories on which PCG ran into its cap get the direct solution, the others keep their PCG solution. */
int gbd_bcr_solve_flagged_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_gamma, float *d_lambda,
                              const uint8_t *d_only_if, void *stream);

/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
uint64_t gbd_pcg_launch_count(void);

/* Diagnostics: device buffer that the timeline builds (mode 10) fill with per-thread %clock stamps of PCG
 * iterations 8..11, uint32 [4][12][cluster * threads]; NULL (the default) disables it (tools/timeline.py). */
void gbd_pcg_set_debug_buffer(void *d_buf);

#ifdef __cplusplus
}
#endif
#endif /* GBD_PCG_H */
