// gbd_capi.cu -- C ABI of libgbdpcg.so (declared in include/gbd_pcg.h): variant table, launch
// configuration (cluster dims, opt-in shared memory), device-pointer / linsys-window / batched /
// host-buffer entry points.  There is no CPU fallback anywhere in this file: without a CUDA
// device every compute entry returns GBD_PCG_ERR_NODEVICE / GBD_PCG_ERR_CUDA.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/gbd_pcg.h"
#include "../../include/gbd/gbd_cluster_pcg.cuh"      // PcgArgs
#include "../../include/gbd/gbd_grid_pcg.cuh"         // GridArgs
#include "../../include/gbd/gbd_schur.cuh"
#include "../../include/gbd/gbd_bcr.cuh"
#include "gbd_variants.h"

namespace {

using namespace gbd;
using gbdlib::Variant;
using gbdlib::mode_is_fast;
using gbdlib::mode_is_grid;

// The table of compiled (n, N, cluster size, kernel family) variants, filled by the registrars of variants_*.cu.
std::vector<Variant> &variants()
{
    static std::vector<Variant> v = [] {
        std::vector<Variant> t;
        gbdlib::register_exact_v4(t);
        gbdlib::register_exact_v1v2(t);
        gbdlib::register_exact_v3v5(t);
        gbdlib::register_fast(t);
        gbdlib::register_fastb(t);
        gbdlib::register_grid(t);
        return t;
    }();
    return v;
}

// Measured defaults (B200; profiles/r01c_ab_bench.json for the bit-exact family, profiles/r02_* for the fast one): per
// (n, N, numerics, single / batched) the preferred (cluster size, mode) in order; the first one the device can place is
// used.  Shapes not listed take the first usable variant of the requested numerics in table order, then any usable variant
// (a shape with no fast kernel is solved by the bit-exact one: stronger, never weaker).
struct Pref { uint32_t n, N; bool fast, batched; uint32_t C; int mode; };
const Pref g_prefs[] = {
    // bit-exact, single solve
    {14, 32, false, false, 4, 7},    {14, 64, false, false, 8, 7},    {14, 128, false, false, 16, 2},
    {14, 128, false, false, 8, 2},   {14, 256, false, false, 16, 2},  {14, 256, false, false, 8, 2},
    {14, 512, false, false, 16, 5},  {14, 512, false, false, 16, 2},
    // bit-exact, batched
    {14, 128, false, true, 4, 11},   {14, 32, false, true, 2, 11},    {14, 64, false, true, 2, 6},
    {14, 128, false, true, 8, 6},    {14, 32, false, true, 1, 6},
    // tolerance parity, single solve
    {14, 128, true, false, 16, 20},  {14, 128, true, false, 8, 20},   {14, 32, true, false, 4, 20},
    {14, 64, true, false, 8, 20},    {14, 256, true, false, 16, 20},
    // tolerance parity, batched: the packed-row kernels (few CTAs per system, 16-32 knot rows each: many systems in flight), then
    // the bit-exact v5 kernel (the single-solve fast kernels keep one system per 8-16 SMs and lose to it on throughput:
    // 131 K vs 210 K systems/s at 1024 x N = 128, profiles/r02_ab_batched.log)
    {14, 128, true, true, 4, 27},    {14, 32, true, true, 1, 27},     {14, 64, true, true, 2, 27},
    {14, 256, true, true, 8, 27},    {14, 512, true, true, 16, 27},
    {14, 128, true, true, 4, 11},    {14, 32, true, true, 2, 11},
};

struct Tuning { uint32_t n, N; bool f64; uint32_t C; int mode; };
std::vector<Tuning> &tunings() { static std::vector<Tuning> t; return t; }
std::mutex g_mu;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_numerics{-1};   // -1: not chosen yet (environment, else fast)
uint32_t *g_dbg = nullptr;   // timeline builds write %clock stamps here
thread_local int tl_cuda_err = 0;

int numerics()
{
    int m = g_numerics.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("GBD_PCG_NUMERICS");
        m = (e && (!strcmp(e, "exact") || !strcmp(e, "bitexact") || !strcmp(e, "0"))) ? GBD_PCG_NUMERICS_BITEXACT : GBD_PCG_NUMERICS_FAST;
        g_numerics.store(m, std::memory_order_relaxed);
    }
    return m;
}

int cuda_fail(cudaError_t e)
{
    tl_cuda_err = (int)e;
    cudaGetLastError();   // clear the sticky-less error state
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return GBD_PCG_ERR_NODEVICE;
    return GBD_PCG_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_); } while (0)

// One CUDA device per process (gbd_pcg.h): kernel attributes, cluster occupancy, work counters, packet workspaces and the
// pinned result mirror are created once and belong to the device that was current at the first compute call.  A call made
// with another device current would launch unprepared kernels on foreign buffers, so it is refused instead.
int g_device = -1;
int check_device()
{
    int dev = -1;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) g_device = dev;
    return dev == g_device ? GBD_PCG_OK : GBD_PCG_ERR_DEVICE;
}

// Wait for a stream by polling it: the wake-up latency of cudaStreamSynchronize is several microseconds, which is a few
// per cent of a 100 us solve.  Solves that run long fall back to the blocking wait so a core is not burnt for nothing.
cudaError_t spin_sync(cudaStream_t st)
{
    timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (uint32_t spins = 0;; ++spins) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
        if ((spins & 255u) == 255u) {
            clock_gettime(CLOCK_MONOTONIC, &b);
            if ((b.tv_sec - a.tv_sec) * 1000000000L + (b.tv_nsec - a.tv_nsec) > 2000000L) return cudaStreamSynchronize(st);
        }
    }
}

Variant *lookup(uint32_t n, uint32_t N, bool f64, uint32_t C, int mode)
{
    for (auto &v : variants())
        if (v.n == n && v.N == N && v.f64 == f64 && (C == 0 || v.C == C) && (mode < 0 || v.mode == mode)) return &v;
    return nullptr;
}

Variant *find_variant(uint32_t n, uint32_t N, bool f64, bool batched)
{
    {
        uint32_t wantC = 0; int wantMode = -1;
        for (auto &t : tunings()) if (t.n == n && t.N == N && t.f64 == f64) { wantC = t.C; wantMode = t.mode; }
        if (wantC || wantMode >= 0) return lookup(n, N, f64, wantC, wantMode);
    }
    const bool fast = !f64 && numerics() == GBD_PCG_NUMERICS_FAST;
    for (int pass = fast ? 0 : 1; pass < 2; ++pass) {        // pass 0: fast kernels, pass 1: bit-exact kernels
        const bool want_fast = pass == 0;
        if (!f64)
            for (const Pref &p : g_prefs)
                if (p.n == n && p.N == N && p.fast == want_fast && p.batched == batched) {
                    Variant *v = lookup(n, N, false, p.C, p.mode);
                    if (v && !v->unusable) return v;
                }
        if (want_fast && batched) continue;                 // batches: only the measured preferences above, else bit-exact
        Variant *grid = nullptr;
        for (auto &v : variants()) {
            if (v.n != n || v.N != N || v.f64 != f64 || v.unusable || mode_is_fast(v.mode) != want_fast) continue;
            if (v.mode == 10 || v.mode == 14 || v.mode == gbdlib::MODE_FAST_PROF || v.mode == gbdlib::MODE_FAST_B_PROF) continue;      // timeline builds are never a default
            if (gbdlib::mode_is_packed(v.mode)) continue;                                           // batch kernels: only by preference
            if (v.mode == gbdlib::MODE_FAST_GRID2) continue;                                        // measured slower than the flat exchange: only by tuning                                       // batch kernels: only by preference
            if (mode_is_grid(v.mode)) { if (!grid) grid = &v; continue; }                          // whole-GPU kernels last
            return &v;
        }
        if (grid) return grid;
    }
    return nullptr;
}

// how many clusters of this variant the device can hold at once (persistent-grid size for batches); 0 when the device
// cannot place a cluster of this size at all (a 16-CTA cluster needs 16 free SMs inside one GPC)
int max_clusters(Variant &v, int *out)
{
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(v.C * 1024);
    cfg.blockDim = dim3(v.nt);
    cfg.dynamicSmemBytes = v.smem;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = v.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, v.kernel, &cfg) != cudaSuccess) {
        cudaGetLastError();
        nc = 0;
    }
    *out = nc;
    return GBD_PCG_OK;
}

int prepare(Variant &v)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (v.prepared) return GBD_PCG_OK;
    if (v.smem > 48 * 1024) CK(cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
    if (v.C > 8 && !mode_is_grid(v.mode)) CK(cudaFuncSetAttribute(v.kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (!mode_is_grid(v.mode)) {
        // GBD_PCG_MAX_CLUSTER=c pretends the device cannot place clusters larger than c (tests of the fallback chain)
        static const int cap = [] { const char *e = getenv("GBD_PCG_MAX_CLUSTER"); return e ? atoi(e) : 0; }();
        max_clusters(v, &v.resident);
        v.unusable = v.resident < 1 || (cap > 0 && (int)v.C > cap);
    } else if (v.grid_cluster > 1) {
        // a clustered grid kernel needs ALL its clusters co-resident (they poll each other): usable only if the device can hold them
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        cfg.gridDim = dim3(v.C);
        cfg.blockDim = dim3(v.nt);
        cfg.dynamicSmemBytes = v.smem;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = v.grid_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, v.kernel, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
        v.unusable = (uint32_t)nc * v.grid_cluster < v.C;
    }
    v.prepared = true;
    return GBD_PCG_OK;
}

// the variant a launch will use: the tuned or default one, or -- when this device cannot place its cluster size -- the
// next usable default for the shape (every variant computes the same bits, so only speed changes)
Variant *resolve_variant(uint32_t n, uint32_t N, bool f64, bool batched, int *rc)
{
    *rc = GBD_PCG_ERR_UNSUPPORTED;
    for (;;) {
        Variant *v = find_variant(n, N, f64, batched);
        if (!v) return nullptr;
        const int r = prepare(*v);
        if (r) { *rc = r; return nullptr; }
        if (!v->unusable) { *rc = GBD_PCG_OK; return v; }
        if (find_variant(n, N, f64, batched) == v) return nullptr;      // explicitly tuned to a shape this device cannot run
    }
}

// grid kernel: one cooperative launch per system; packet workspace per (kernel, stream), zeroed once
struct GridWs { unsigned long long *ws; uint32_t epoch; };
std::map<std::pair<const void *, cudaStream_t>, GridWs> &grid_ws() { static std::map<std::pair<const void *, cudaStream_t>, GridWs> m; return m; }

template <typename T>
int launch_grid(Variant &v, uint32_t batch, const T *S, const T *P, const T *g, T *lam, T *r, T *p, uint32_t *iters,
                uint8_t *flag, uint32_t max_iter, T tol, cudaStream_t st, bool no_tma)
{
    GridWs *w;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto key = std::make_pair(v.kernel, st);
        auto it = grid_ws().find(key);
        if (it == grid_ws().end()) {
            GridWs nw{nullptr, 0};
            CK(cudaMalloc((void **)&nw.ws, v.ws_words * sizeof(unsigned long long)));
            CK(cudaMemsetAsync(nw.ws, 0, v.ws_words * sizeof(unsigned long long), st));
            it = grid_ws().emplace(key, nw).first;
        }
        w = &it->second;
    }
    const size_t ms = (size_t)3 * v.n * v.n * v.N, vs = (size_t)v.n * v.N;
    for (uint32_t i = 0; i < batch; ++i) {
        GridArgs<T> ga;
        ga.a.S = S + i * ms; ga.a.Pinv = P + i * ms; ga.a.gamma = g + i * vs; ga.a.lambda = lam + i * vs;
        ga.a.r_out = r ? r + i * vs : nullptr; ga.a.p_out = p ? p + i * vs : nullptr;
        ga.a.iters = iters + i; ga.a.max_iter_exit = flag + i; ga.a.batch = 1; ga.a.max_iter = max_iter; ga.a.exit_tol = tol;
        ga.a.use_tma = (!no_tma && (((uintptr_t)ga.a.S | (uintptr_t)ga.a.Pinv) & 15u) == 0) ? 1u : 0u;
        ga.ws = w->ws;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            ga.epoch_base = w->epoch;
            w->epoch += 2u * max_iter + 4u;                  // upper bound on the phases one launch can use
        }
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[2];
        cfg.gridDim = dim3(v.C);
        cfg.blockDim = dim3(v.nt);
        cfg.dynamicSmemBytes = v.smem;
        cfg.stream = st;
        at[0].id = cudaLaunchAttributeCooperative;           // co-residency of all CTAs (they poll each other)
        at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (v.grid_cluster > 1) {                            // two-level exchange: DSMEM inside clusters, L2 between their leaders
            at[1].id = cudaLaunchAttributeClusterDimension;
            at[1].val.clusterDim.x = v.grid_cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
            cfg.numAttrs = 2;
        }
        void *args[] = {&ga};
        CK(cudaLaunchKernelExC(&cfg, v.kernel, args));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return GBD_PCG_OK;
}

template <typename T>
int launch(uint32_t n, uint32_t N, uint32_t batch, const T *S, const T *P, const T *g, T *lam, T *r, T *p,
           uint32_t *iters, uint8_t *flag, uint32_t max_iter, T tol, cudaStream_t st, bool no_tma = false,
           uint32_t *host_result = nullptr)
{
    if (!S || !P || !g || !lam || !iters || !flag || batch == 0 || N < 2 || n == 0) return GBD_PCG_ERR_BADARG;
    if (!lookup(n, N, sizeof(T) == 8, 0, -1)) return GBD_PCG_ERR_UNSUPPORTED;
    int rc = check_device();
    if (rc) return rc;
    Variant *v = resolve_variant(n, N, sizeof(T) == 8, batch > 1, &rc);
    if (!v) return rc;

    if (mode_is_grid(v->mode)) return launch_grid<T>(*v, batch, S, P, g, lam, r, p, iters, flag, max_iter, tol, st, no_tma);

    uint32_t nclusters = batch;
    if (batch > 1 && nclusters > (uint32_t)v->resident) nclusters = (uint32_t)v->resident;   // persistent clusters loop over the batch
    PcgArgs<T> a;
    a.S = S; a.Pinv = P; a.gamma = g; a.lambda = lam; a.r_out = r; a.p_out = p;
    a.iters = iters; a.max_iter_exit = flag; a.batch = batch; a.max_iter = max_iter; a.exit_tol = tol;
    a.use_tma = (!no_tma && (((uintptr_t)S | (uintptr_t)P) & 15u) == 0) ? 1u : 0u;
    a.dbg = g_dbg;
    a.host_result = host_result;
    if constexpr (sizeof(T) == 4) {
        // v5 with more systems than resident clusters: clusters draw their next system from a counter (zeroed on this
        // stream ahead of the launch) -- GBD_PCG_STATIC_BATCH=1 keeps the fixed stride (A/B)
        static const bool static_batch = [] { const char *e = getenv("GBD_PCG_STATIC_BATCH"); return e && atoi(e) != 0; }();
        const bool draws = v->mode == 11 || v->mode == 12 || mode_is_fast(v->mode);
        if (draws && batch > nclusters && !static_batch) {
            // one counter per stream: the memset and the kernel that draws from it are stream-ordered, so the next launch on
            // the same stream cannot zero the counter under a kernel that is still running, however many launches are queued
            static std::map<cudaStream_t, uint32_t *> counters;
            uint32_t *ctr;
            {
                std::lock_guard<std::mutex> lk(g_mu);
                auto it = counters.find(st);
                if (it == counters.end()) {
                    uint32_t *c = nullptr;
                    CK(cudaMalloc((void **)&c, sizeof(uint32_t)));
                    it = counters.emplace(st, c).first;
                }
                ctr = it->second;
            }
            CK(cudaMemsetAsync(ctr, 0, sizeof(uint32_t), st));
            a.work_counter = ctr;
        }
    }

    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(nclusters * v->C);
    cfg.blockDim = dim3(v->nt);
    cfg.dynamicSmemBytes = v->smem;
    cfg.stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = v->C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void *args[] = {&a};
    CK(cudaLaunchKernelExC(&cfg, v->kernel, args));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GBD_PCG_OK;
}

}  // namespace

struct gbd_pcg_plan {
    uint32_t n, N, batch;
    bool f64;
    size_t esz;
    void *dS, *dP, *dg, *dl;
    uint32_t *d_iters;
    uint8_t *d_flag;
    uint32_t *h_iters_pin;   // pinned + mapped result slots for the zero-copy path
    uint8_t *h_flag_pin;
    uint32_t *z_iters;       // their device aliases
    uint8_t *z_flag;
    cudaStream_t st;
    // device aliases of caller buffers this plan has already seen (a pinned buffer keeps its alias for as long as it stays
    // registered; unregistering a buffer while a plan that has used it is alive is not supported)
    struct Alias { const void *host; const void *dev; } alias[512];
};

extern "C" {

int gbd_pcg_abi_version(void) { return GBD_PCG_ABI_VERSION; }

const char *gbd_pcg_strerror(int s)
{
    switch (s) {
        case GBD_PCG_OK: return "ok";
        case GBD_PCG_ERR_UNSUPPORTED: return "no kernel compiled for this (state_size, knot_points, dtype)";
        case GBD_PCG_ERR_BADARG: return "bad argument";
        case GBD_PCG_ERR_CUDA: return "CUDA error (see gbd_pcg_last_cuda_error)";
        case GBD_PCG_ERR_NODEVICE: return "no CUDA device (this library has no CPU fallback)";
        case GBD_PCG_ERR_DEVICE: return "the current CUDA device is not the one this process first used the library on";
        default: return "unknown status";
    }
}

int gbd_pcg_last_cuda_error(void) { return tl_cuda_err; }

int gbd_pcg_supported(uint32_t n, uint32_t N, int is_f64)
{
    for (auto &v : variants()) if (v.n == n && v.N == N && v.f64 == (is_f64 != 0)) return 1;
    return 0;
}

int gbd_pcg_num_variants(void) { return (int)variants().size(); }

int gbd_pcg_variant_at(int i, uint32_t *n, uint32_t *N, uint32_t *cluster, int *regs, int *is_f64,
                       uint32_t *threads, size_t *smem_bytes)
{
    if (i < 0 || i >= (int)variants().size()) return GBD_PCG_ERR_BADARG;
    const Variant &v = variants()[i];
    if (n) *n = v.n;
    if (N) *N = v.N;
    if (cluster) *cluster = v.C;
    if (regs) *regs = v.mode;
    if (is_f64) *is_f64 = v.f64;
    if (threads) *threads = v.nt;
    if (smem_bytes) *smem_bytes = v.smem;
    return GBD_PCG_OK;
}

int gbd_pcg_set_numerics(int m)
{
    if (m != GBD_PCG_NUMERICS_BITEXACT && m != GBD_PCG_NUMERICS_FAST) return GBD_PCG_ERR_BADARG;
    g_numerics.store(m, std::memory_order_relaxed);
    return GBD_PCG_OK;
}

int gbd_pcg_get_numerics(void) { return numerics(); }

int gbd_pcg_resolved_variant(uint32_t n, uint32_t N, int is_f64, int batched, uint32_t *cluster, int *mode, uint32_t *threads,
                             size_t *smem_bytes, char *kernel_name, size_t kernel_name_len)
{
    int rc = check_device();
    if (rc) return rc;
    Variant *v = resolve_variant(n, N, is_f64 != 0, batched != 0, &rc);
    if (!v) return rc;
    if (cluster) *cluster = v->C;
    if (mode) *mode = v->mode;
    if (threads) *threads = v->nt;
    if (smem_bytes) *smem_bytes = v->smem;
    if (kernel_name && kernel_name_len) snprintf(kernel_name, kernel_name_len, "%s", v->name);
    return GBD_PCG_OK;
}

int gbd_pcg_set_tuning(uint32_t n, uint32_t N, int is_f64, uint32_t cluster, int regs)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto &t = tunings();
    for (size_t i = 0; i < t.size(); ++i)
        if (t[i].n == n && t[i].N == N && t[i].f64 == (is_f64 != 0)) { t.erase(t.begin() + i); break; }
    if (cluster == 0 && regs < 0) return GBD_PCG_OK;
    bool ok = false;
    for (auto &v : variants())
        if (v.n == n && v.N == N && v.f64 == (is_f64 != 0) && (cluster == 0 || v.C == cluster) && (regs < 0 || v.mode == regs)) ok = true;
    if (!ok) return GBD_PCG_ERR_UNSUPPORTED;
    t.push_back(Tuning{n, N, is_f64 != 0, cluster, regs});
    return GBD_PCG_OK;
}

int gbd_pcg_solve_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_Pinv, const float *d_gamma,
                      float *d_lambda, float *d_r, float *d_p, float *, float *, uint32_t *d_iters,
                      uint8_t *d_max_iter_exit, uint32_t max_iter, float exit_tol, void *stream)
{
    return launch<float>(n, N, 1, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_max_iter_exit, max_iter,
                         exit_tol, (cudaStream_t)stream);
}

int gbd_pcg_solve_f64(uint32_t n, uint32_t N, const double *d_S, const double *d_Pinv, const double *d_gamma,
                      double *d_lambda, double *d_r, double *d_p, double *, double *, uint32_t *d_iters,
                      uint8_t *d_max_iter_exit, uint32_t max_iter, double exit_tol, void *stream)
{
    return launch<double>(n, N, 1, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_max_iter_exit, max_iter,
                          exit_tol, (cudaStream_t)stream);
}

int gbd_pcg_linsys_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_Pinv, const float *d_gamma,
                       float *d_lambda, float *d_r, float *d_p, uint32_t *d_iters, uint8_t *d_max_iter_exit,
                       uint32_t max_iter, float exit_tol, uint32_t *h_iters, uint8_t *h_max_iter_exit,
                       double *elapsed_us)
{
    if (!h_iters || !h_max_iter_exit) return GBD_PCG_ERR_BADARG;
    { const int drc = check_device(); if (drc) return drc; }
    // same observable window as include/pcg/sqp.cuh:224-241 (device idle before, both results on the host and the device
    // idle after), but the two blocking cudaMemcpy + cudaDeviceSynchronize become two async copies into pinned slots and
    // one spin on the stream: ~15 us less host-side latency per SQP iteration
    static uint32_t *pin = nullptr, *pin_dev = nullptr;     // [0] iters, [1] flag: pinned + mapped, written by the kernel itself
    {   // (the window itself is one caller at a time by construction: it brackets the legacy default stream, like the reference's)
        std::lock_guard<std::mutex> lk(g_mu);
        if (!pin) {
            CK(cudaHostAlloc((void **)&pin, 2 * sizeof(uint32_t), cudaHostAllocMapped));
            CK(cudaHostGetDevicePointer((void **)&pin_dev, pin, 0));
        }
    }
    timespec t0, t1;
    CK(cudaDeviceSynchronize());
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int vrc;
    Variant *v = resolve_variant(n, N, false, false, &vrc);
    if (!v) return vrc;
    const bool mirror = !mode_is_grid(v->mode);              // the cluster kernels write the two results into the mapped slots
    int rc = launch<float>(n, N, 1, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_max_iter_exit, max_iter,
                           exit_tol, (cudaStream_t)0, false, mirror ? pin_dev : nullptr);
    if (rc) return rc;
    if (!mirror) {
        CK(cudaMemcpyAsync(&pin[0], d_iters, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)0));
        CK(cudaMemcpyAsync(&pin[1], d_max_iter_exit, sizeof(uint8_t), cudaMemcpyDeviceToHost, (cudaStream_t)0));
    }
    CK(spin_sync((cudaStream_t)0));
    *h_iters = pin[0];
    *h_max_iter_exit = (uint8_t)pin[1];
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (elapsed_us) *elapsed_us = 1e6 * (double)(t1.tv_sec - t0.tv_sec) + 1e-3 * (double)(t1.tv_nsec - t0.tv_nsec);
    return GBD_PCG_OK;
}

int gbd_pcg_solve_batched_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_Pinv,
                              const float *d_gamma, float *d_lambda, float *d_r, float *d_p, uint32_t *d_iters,
                              uint8_t *d_max_iter_exit, uint32_t max_iter, float exit_tol, void *stream)
{
    return launch<float>(n, N, batch, d_S, d_Pinv, d_gamma, d_lambda, d_r, d_p, d_iters, d_max_iter_exit, max_iter,
                         exit_tol, (cudaStream_t)stream);
}

int gbd_pcg_plan_create(uint32_t n, uint32_t N, uint32_t batch, int is_f64, gbd_pcg_plan **out)
{
    if (!out || batch == 0) return GBD_PCG_ERR_BADARG;
    if (!gbd_pcg_supported(n, N, is_f64)) return GBD_PCG_ERR_UNSUPPORTED;
    { const int drc = check_device(); if (drc) return drc; }
    gbd_pcg_plan *p = (gbd_pcg_plan *)calloc(1, sizeof(gbd_pcg_plan));
    p->n = n; p->N = N; p->batch = batch; p->f64 = is_f64 != 0; p->esz = is_f64 ? 8 : 4;
    const size_t mat = (size_t)3 * n * n * N * batch * p->esz, vec = (size_t)n * N * batch * p->esz;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaMalloc(&p->dS, mat)) != cudaSuccess || (e = cudaMalloc(&p->dP, mat)) != cudaSuccess ||
        (e = cudaMalloc(&p->dg, vec)) != cudaSuccess || (e = cudaMalloc(&p->dl, vec)) != cudaSuccess ||
        (e = cudaMalloc((void **)&p->d_iters, sizeof(uint32_t) * batch)) != cudaSuccess ||
        (e = cudaMalloc((void **)&p->d_flag, batch)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&p->h_iters_pin, sizeof(uint32_t) * batch, cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&p->h_flag_pin, batch, cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer((void **)&p->z_iters, p->h_iters_pin, 0)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer((void **)&p->z_flag, p->h_flag_pin, 0)) != cudaSuccess) {
        gbd_pcg_plan_destroy(p);
        return cuda_fail(e);
    }
    *out = p;
    return GBD_PCG_OK;
}

int gbd_pcg_plan_invalidate(gbd_pcg_plan *p)
{
    if (!p) return GBD_PCG_ERR_BADARG;
    memset(p->alias, 0, sizeof(p->alias));
    return GBD_PCG_OK;
}

int gbd_pcg_plan_destroy(gbd_pcg_plan *p)
{
    if (!p) return GBD_PCG_OK;
    cudaFree(p->dS); cudaFree(p->dP); cudaFree(p->dg); cudaFree(p->dl); cudaFree(p->d_iters); cudaFree(p->d_flag);
    if (p->h_iters_pin) cudaFreeHost(p->h_iters_pin);
    if (p->h_flag_pin) cudaFreeHost(p->h_flag_pin);
    if (p->st) cudaStreamDestroy(p->st);
    free(p);
    return GBD_PCG_OK;
}

}  // extern "C"

namespace {
// device alias of a pinned / registered host pointer, or nullptr for pageable memory
const void *device_alias(const void *h)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// GBD_PCG_ZEROCOPY: 1 (default) = pinned host buffers are read/written by the kernel itself over PCIe
// (TMA bulk loads straight from host memory, lambda written straight back), 2 = same with plain
// loads instead of TMA, 0 = always stage through device buffers with cudaMemcpyAsync.
int zerocopy_mode()
{
    static int mode = [] { const char *e = getenv("GBD_PCG_ZEROCOPY"); return e ? atoi(e) : 1; }();
    return mode;
}

template <typename T>
int plan_solve_host(gbd_pcg_plan *p, const T *hS, const T *hP, const T *hg, T *hl, uint32_t max_iter, T tol,
                    uint32_t *h_iters, uint8_t *h_flag)
{
    if (!p || !hS || !hP || !hg || !hl || !h_iters || !h_flag) return GBD_PCG_ERR_BADARG;
    if (p->f64 != (sizeof(T) == 8)) return GBD_PCG_ERR_BADARG;
    if (zerocopy_mode() != 0) {
        auto alias_of = [&](const void *h) -> const void * {
            gbd_pcg_plan::Alias &slot = p->alias[((uintptr_t)h >> 8) & 511u];
            if (slot.host == h) return slot.dev;
            const void *d = device_alias(h);
            if (d) { slot.host = h; slot.dev = d; }
            return d;
        };
        const T *zS = (const T *)alias_of(hS), *zP = (const T *)alias_of(hP), *zg = (const T *)alias_of(hg);
        T *zl = (T *)alias_of(hl);
        if (zS && zP && zg && zl) {
            uint32_t *zi = p->z_iters; uint8_t *zf = p->z_flag;
            int rc = launch<T>(p->n, p->N, p->batch, zS, zP, zg, zl, (T *)nullptr, (T *)nullptr, zi, zf, max_iter, tol, p->st,
                               zerocopy_mode() == 2);
            if (rc) return rc;
            CK(spin_sync(p->st));
            memcpy(h_iters, p->h_iters_pin, sizeof(uint32_t) * p->batch);
            memcpy(h_flag, p->h_flag_pin, p->batch);
            return GBD_PCG_OK;
        }
    }
    const size_t mat = (size_t)3 * p->n * p->n * p->N * p->batch * sizeof(T), vec = (size_t)p->n * p->N * p->batch * sizeof(T);
    CK(cudaMemcpyAsync(p->dS, hS, mat, cudaMemcpyHostToDevice, p->st));
    CK(cudaMemcpyAsync(p->dP, hP, mat, cudaMemcpyHostToDevice, p->st));
    CK(cudaMemcpyAsync(p->dg, hg, vec, cudaMemcpyHostToDevice, p->st));
    CK(cudaMemcpyAsync(p->dl, hl, vec, cudaMemcpyHostToDevice, p->st));
    int rc = launch<T>(p->n, p->N, p->batch, (const T *)p->dS, (const T *)p->dP, (const T *)p->dg, (T *)p->dl,
                       (T *)nullptr, (T *)nullptr, p->d_iters, p->d_flag, max_iter, tol, p->st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(hl, p->dl, vec, cudaMemcpyDeviceToHost, p->st));
    CK(cudaMemcpyAsync(h_iters, p->d_iters, sizeof(uint32_t) * p->batch, cudaMemcpyDeviceToHost, p->st));
    CK(cudaMemcpyAsync(h_flag, p->d_flag, p->batch, cudaMemcpyDeviceToHost, p->st));
    CK(cudaStreamSynchronize(p->st));
    return GBD_PCG_OK;
}
}  // namespace

extern "C" {

int gbd_pcg_plan_solve_host_f32(gbd_pcg_plan *plan, const float *h_S, const float *h_Pinv, const float *h_gamma,
                                float *h_lambda, uint32_t max_iter, float exit_tol, uint32_t *h_iters,
                                uint8_t *h_max_iter_exit)
{
    return plan_solve_host<float>(plan, h_S, h_Pinv, h_gamma, h_lambda, max_iter, exit_tol, h_iters, h_max_iter_exit);
}

int gbd_pcg_plan_solve_host_f64(gbd_pcg_plan *plan, const double *h_S, const double *h_Pinv, const double *h_gamma,
                                double *h_lambda, uint32_t max_iter, double exit_tol, uint32_t *h_iters,
                                uint8_t *h_max_iter_exit)
{
    return plan_solve_host<double>(plan, h_S, h_Pinv, h_gamma, h_lambda, max_iter, exit_tol, h_iters, h_max_iter_exit);
}

}  // extern "C"

namespace {
std::atomic<int> g_schur_team{-1};          // gbd_schur_set_team
template <uint32_t n, uint32_t m>
int schur_launch(uint32_t N, uint32_t batch, float *G, const float *C, const float *g, const float *c, float *S, float *P, float *gam,
                 float rho, cudaStream_t st)
{
    using K = gbd::SchurShape<n, m>;
    { const int drc = check_device(); if (drc) return drc; }
    const int team = g_schur_team.load(std::memory_order_relaxed);
    const bool warp_rows = team == 1 || (team < 0 && (uint64_t)batch * N >= 4096u);
    if (warp_rows) {                                                      // batches: one warp per block row, four rows per CTA
        constexpr uint32_t RPC = K::NT / 32;
        gbd::schur_phase1_kernel<n, m, true><<<dim3((N + RPC - 1) / RPC, batch), K::NT, RPC * K::P1_STRIDE * sizeof(float), st>>>(
            N, G, C, g, c, S, P, gam, rho);
    } else {
        gbd::schur_phase1_kernel<n, m, false><<<dim3(N, batch), K::NT, K::P1_FLOATS * sizeof(float), st>>>(N, G, C, g, c, S, P, gam, rho);
    }
    {   // phase 2 with programmatic stream serialization: its launch overlaps phase 1, griddepcontrol.wait orders the data
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        cfg.gridDim = warp_rows ? dim3((N + K::NT / 32 - 1) / (K::NT / 32), batch) : dim3(N, batch);
        cfg.blockDim = dim3(K::NT);
        cfg.dynamicSmemBytes = (warp_rows ? K::P2W_FLOATS : K::P2_FLOATS) * sizeof(float);
        cfg.stream = st;
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const float *Sc = S;
        if (warp_rows) CK(cudaLaunchKernelEx(&cfg, gbd::schur_phase2_warp_kernel<n, m>, N, G, Sc, P));
        else CK(cudaLaunchKernelEx(&cfg, gbd::schur_phase2_kernel<n, m>, N, G, Sc, P));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return GBD_PCG_OK;
}
template <uint32_t n, uint32_t m>
int dz_launch(uint32_t N, uint32_t batch, const float *Gi, const float *C, const float *g, const float *lam, float *dz, cudaStream_t st)
{
    { const int drc = check_device(); if (drc) return drc; }
    gbd::compute_dz_kernel<n, m><<<dim3(N, batch), 64, 0, st>>>(N, Gi, C, g, lam, dz);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GBD_PCG_OK;
}
}  // namespace

// (state_size, control_size) pairs compiled in: IIWA (14, 7) and small shapes for tests
#define GBD_SCHUR_SHAPES(X) X(14, 7) X(6, 3) X(4, 2) X(2, 1)

extern "C" {

int gbd_schur_supported(uint32_t n, uint32_t m)
{
#define X(a, b) if (n == a && m == b) return 1;
    GBD_SCHUR_SHAPES(X)
#undef X
    return 0;
}

int gbd_schur_set_team(int mode) { return g_schur_team.exchange(mode < 0 ? -1 : (mode ? 1 : 0)); }

int gbd_form_schur_system_f32(uint32_t n, uint32_t m, uint32_t N, float *d_G, const float *d_C, const float *d_g,
                              const float *d_c, float *d_S, float *d_Pinv, float *d_gamma, float rho, void *stream)
{
    if (!d_G || !d_C || !d_g || !d_c || !d_S || !d_Pinv || !d_gamma || N < 2) return GBD_PCG_ERR_BADARG;
#define X(a, b) if (n == a && m == b) return schur_launch<a, b>(N, 1, d_G, d_C, d_g, d_c, d_S, d_Pinv, d_gamma, rho, (cudaStream_t)stream);
    GBD_SCHUR_SHAPES(X)
#undef X
    return GBD_PCG_ERR_UNSUPPORTED;
}

int gbd_compute_dz_f32(uint32_t n, uint32_t m, uint32_t N, const float *d_Ginv, const float *d_C, const float *d_g,
                       const float *d_lambda, float *d_dz, void *stream)
{
    if (!d_Ginv || !d_C || !d_g || !d_lambda || !d_dz || N < 2) return GBD_PCG_ERR_BADARG;
#define X(a, b) if (n == a && m == b) return dz_launch<a, b>(N, 1, d_Ginv, d_C, d_g, d_lambda, d_dz, (cudaStream_t)stream);
    GBD_SCHUR_SHAPES(X)
#undef X
    return GBD_PCG_ERR_UNSUPPORTED;
}

}  // extern "C"

// ---- f3: one SQP linear-system step (assembly -> PCG -> dz) for `batch` trajectories, enqueued without a host round trip
struct gbd_step_plan {
    uint32_t n, m, N, batch;
    float *dS, *dP, *dgam;
    uint32_t *d_iters;
    uint8_t *d_flag;
    uint32_t *h_iters;     // pinned
    uint8_t *h_flag;       // pinned
};

extern "C" {

int gbd_step_plan_create(uint32_t n, uint32_t m, uint32_t N, uint32_t batch, gbd_step_plan **out)
{
    if (!out || batch == 0 || N < 2) return GBD_PCG_ERR_BADARG;
    if (!gbd_schur_supported(n, m) || !gbd_pcg_supported(n, N, 0)) return GBD_PCG_ERR_UNSUPPORTED;
    { const int drc = check_device(); if (drc) return drc; }
    gbd_step_plan *p = (gbd_step_plan *)calloc(1, sizeof(gbd_step_plan));
    p->n = n; p->m = m; p->N = N; p->batch = batch;
    const size_t mat = (size_t)3 * n * n * N * batch * sizeof(float), vec = (size_t)n * N * batch * sizeof(float);
    cudaError_t e;
    if ((e = cudaMalloc((void **)&p->dS, mat)) != cudaSuccess || (e = cudaMalloc((void **)&p->dP, mat)) != cudaSuccess ||
        (e = cudaMalloc((void **)&p->dgam, vec)) != cudaSuccess ||
        (e = cudaMalloc((void **)&p->d_iters, sizeof(uint32_t) * batch)) != cudaSuccess ||
        (e = cudaMalloc((void **)&p->d_flag, batch)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&p->h_iters, sizeof(uint32_t) * batch, cudaHostAllocDefault)) != cudaSuccess ||
        (e = cudaHostAlloc((void **)&p->h_flag, batch, cudaHostAllocDefault)) != cudaSuccess) {
        gbd_step_plan_destroy(p);
        return cuda_fail(e);
    }
    // the pad tiles are never written by the assembly and never read by the solver; keep them defined
    cudaMemset(p->dS, 0, mat);
    cudaMemset(p->dP, 0, mat);
    *out = p;
    return GBD_PCG_OK;
}

int gbd_step_plan_destroy(gbd_step_plan *p)
{
    if (!p) return GBD_PCG_OK;
    cudaFree(p->dS); cudaFree(p->dP); cudaFree(p->dgam); cudaFree(p->d_iters); cudaFree(p->d_flag);
    if (p->h_iters) cudaFreeHost(p->h_iters);
    if (p->h_flag) cudaFreeHost(p->h_flag);
    free(p);
    return GBD_PCG_OK;
}

static int step_run(gbd_step_plan *p, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                    float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream, bool direct_fallback);

int gbd_step_run_f32(gbd_step_plan *p, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                     float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream)
{
    return step_run(p, d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol, stream, false);
}

int gbd_step_run_fallback_f32(gbd_step_plan *p, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                              float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream)
{
    return step_run(p, d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol, stream, true);
}

static int step_run(gbd_step_plan *p, float *d_G, const float *d_C, const float *d_g, const float *d_c, float rho,
                    float *d_lambda, float *d_dz, uint32_t max_iter, float exit_tol, void *stream, bool direct_fallback)
{
    if (!p || !d_G || !d_C || !d_g || !d_c || !d_lambda || !d_dz) return GBD_PCG_ERR_BADARG;
    // nothing is enqueued unless every stage of the step exists for this shape (d_G and d_lambda are overwritten by it)
    if (direct_fallback && !gbd_bcr_supported(p->n, p->N)) return GBD_PCG_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_device();
    if (rc) return rc;
    rc = GBD_PCG_ERR_UNSUPPORTED;
#define X(a, b) if (p->n == a && p->m == b) rc = schur_launch<a, b>(p->N, p->batch, d_G, d_C, d_g, d_c, p->dS, p->dP, p->dgam, rho, st);
    GBD_SCHUR_SHAPES(X)
#undef X
    if (rc) return rc;
    rc = launch<float>(p->n, p->N, p->batch, p->dS, p->dP, p->dgam, d_lambda, (float *)nullptr, (float *)nullptr, p->d_iters, p->d_flag,
                       max_iter, exit_tol, st);
    if (rc) return rc;
    if (direct_fallback) {      // trajectories on which PCG ran into its cap are solved again, directly (their flags stay set)
        rc = gbd_bcr_solve_flagged_f32(p->n, p->N, p->batch, p->dS, p->dgam, d_lambda, p->d_flag, stream);
        if (rc) return rc;
    }
    rc = GBD_PCG_ERR_UNSUPPORTED;
#define X(a, b) if (p->n == a && p->m == b) rc = dz_launch<a, b>(p->N, p->batch, d_G, d_C, d_g, d_lambda, d_dz, st);
    GBD_SCHUR_SHAPES(X)
#undef X
    if (rc) return rc;
    CK(cudaMemcpyAsync(p->h_iters, p->d_iters, sizeof(uint32_t) * p->batch, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p->h_flag, p->d_flag, p->batch, cudaMemcpyDeviceToHost, st));
    return GBD_PCG_OK;
}

int gbd_step_results(gbd_step_plan *p, uint32_t *h_iters, uint8_t *h_flags, void *stream)
{
    if (!p) return GBD_PCG_ERR_BADARG;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h_iters) memcpy(h_iters, p->h_iters, sizeof(uint32_t) * p->batch);
    if (h_flags) memcpy(h_flags, p->h_flag, p->batch);
    return GBD_PCG_OK;
}

const uint8_t *gbd_step_device_flags(gbd_step_plan *p) { return p ? p->d_flag : nullptr; }

}  // extern "C"

// ---- f4: direct solve by block cyclic reduction in one cluster (include/gbd/gbd_bcr.cuh)
namespace {
// MINB = 1: all registers to one CTA per SM (lowest latency, single solves); MINB = 4: 128-register build so that four
// CTAs share an SM (more systems in flight, batches)
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
int bcr_launch(uint32_t batch, const float *S, const float *g, float *lam, cudaStream_t st, const uint8_t *only_if = nullptr)
{
    using K = gbd::BcrShape<n, N, C>;
    auto kern = gbd::bcr_cluster_kernel<n, N, C, MINB, PROF>;
    static bool prepared = false;
    static int max_clusters = 0;
    { const int drc = check_device(); if (drc) return drc; }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!prepared) {
            if (K::SMEM_BYTES > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
            if (C > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t q = {};
            cudaLaunchAttribute qa[1];
            q.gridDim = dim3(C * 1024); q.blockDim = dim3(K::NT); q.dynamicSmemBytes = K::SMEM_BYTES;
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = C; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            CK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q));
            if (max_clusters < 1) max_clusters = 1;
            prepared = true;
        }
    }
    gbd::BcrArgs a{S, g, lam, batch, only_if, PROF ? g_dbg : nullptr};
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(C * (batch < (uint32_t)max_clusters ? batch : (uint32_t)max_clusters));
    cfg.blockDim = dim3(K::NT);
    cfg.dynamicSmemBytes = K::SMEM_BYTES;
    cfg.stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, a));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GBD_PCG_OK;
}
}  // namespace

// (state_size, knot_points, CTAs per system) compiled in for the direct solver
#define GBD_BCR_SHAPES(X) X(14, 8, 1) X(14, 16, 2) X(14, 32, 4) X(14, 64, 8) X(14, 128, 16) X(14, 256, 16) X(14, 512, 16) X(6, 16, 2) X(2, 4, 1)

extern "C" {

int gbd_bcr_supported(uint32_t n, uint32_t N)
{
#define X(a, b, c) if (n == a && N == b) return 1;
    GBD_BCR_SHAPES(X)
#undef X
    return 0;
}

int gbd_bcr_solve_batched_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_gamma, float *d_lambda,
                              void *stream)
{
    if (!d_S || !d_gamma || !d_lambda || batch == 0) return GBD_PCG_ERR_BADARG;
#define X(a, b, c) if (n == a && N == b) return batch > 1 ? bcr_launch<a, b, c, 4>(batch, d_S, d_gamma, d_lambda, (cudaStream_t)stream) \
                                                        : bcr_launch<a, b, c, 1>(batch, d_S, d_gamma, d_lambda, (cudaStream_t)stream);
    GBD_BCR_SHAPES(X)
#undef X
    return GBD_PCG_ERR_UNSUPPORTED;
}

int gbd_bcr_solve_flagged_f32(uint32_t n, uint32_t N, uint32_t batch, const float *d_S, const float *d_gamma, float *d_lambda,
                              const uint8_t *d_only_if, void *stream)
{
    if (!d_S || !d_gamma || !d_lambda || !d_only_if || batch == 0) return GBD_PCG_ERR_BADARG;
#define X(a, b, c) if (n == a && N == b) return bcr_launch<a, b, c, 4>(batch, d_S, d_gamma, d_lambda, (cudaStream_t)stream, d_only_if);
    GBD_BCR_SHAPES(X)
#undef X
    return GBD_PCG_ERR_UNSUPPORTED;
}

int gbd_bcr_solve_f32(uint32_t n, uint32_t N, const float *d_S, const float *d_gamma, float *d_lambda, void *stream)
{
    // timeline build (diagnostics): with a debug buffer set, the IIWA N = 128 / 32 shapes run the stamped kernel
    if (g_dbg && n == 14 && N == 128) return bcr_launch<14, 128, 16, 1, true>(1, d_S, d_gamma, d_lambda, (cudaStream_t)stream);
    if (g_dbg && n == 14 && N == 32) return bcr_launch<14, 32, 4, 1, true>(1, d_S, d_gamma, d_lambda, (cudaStream_t)stream);
    return gbd_bcr_solve_batched_f32(n, N, 1, d_S, d_gamma, d_lambda, stream);
}

// ---- f4: the QDLDL wire format of the band matrix (include/utils/csr.cuh)
uint32_t gbd_schur_csr_nnz(uint32_t n, uint32_t N) { return gbd::csr_nnz(n, N); }

int gbd_schur_csr_pattern_i32(uint32_t n, uint32_t N, int32_t *d_col_ptr, int32_t *d_row_ind, void *stream)
{
    if (!d_col_ptr || !d_row_ind || N < 1 || n < 1) return GBD_PCG_ERR_BADARG;
    gbd::csr_pattern_kernel<<<N, 64, 0, (cudaStream_t)stream>>>(n, N, d_col_ptr, d_row_ind);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GBD_PCG_OK;
}

int gbd_schur_csr_values_f32(uint32_t n, uint32_t N, const float *d_S, float *d_val, void *stream)
{
    if (!d_S || !d_val || N < 1 || n < 1) return GBD_PCG_ERR_BADARG;
    gbd::csr_values_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(n, N, d_S, d_val);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GBD_PCG_OK;
}

void gbd_pcg_set_debug_buffer(void *d_buf) { g_dbg = (uint32_t *)d_buf; }

uint64_t gbd_pcg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
