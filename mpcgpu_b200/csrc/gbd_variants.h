// gbd_variants.h -- internal to libgbdpcg.so: the table of compiled kernel variants.  The kernels are instantiated in several
// translation units (variants_*.cu, compiled in parallel by mpcgpu_b200/build.py); each one appends its variants to the table
// through a registrar, gbd_capi.cu owns the table, the defaults and the launch code.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace gbdlib {

// mode (what gbd_pcg_set_tuning / gbd_pcg_variant_at call `mode`):
//   bit-exact family (same floating-point operation order as the reference kernel):
//       0 = v1 kernel, tiles in shared memory; 1 = v1 kernel, tiles in registers;
//       2 = v2 kernel (st.async/mbarrier signalling), 1 CTA/SM register budget; 3 = v2, 2 CTAs/SM budget;
//       4 = grid kernel (whole GPU on one system, packets through L2; C then holds the CTA count)
//       5 = v3 kernel (two matrix rows per thread, 8-lane knot rows), 1 CTA/SM register budget; 6 = v3, 2 CTAs/SM
//       7 = v4 kernel (self-validating packets polled in shared memory, register N-way tree), 1 CTA/SM; 8 = v4, 2 CTAs/SM
//      11 = v5 kernel (v3's two-rows-per-thread mapping + v4's packet exchange), 1 CTA/SM; 12 = v5, 2 CTAs/SM
//      10 = v4 timeline build, 14 = v2 timeline build (%clock stamps of iterations 8..11 into gbd_pcg_set_debug_buffer())
//   tolerance-parity family (GBD_PCG_NUMERICS_FAST; include/gbd/gbd_cluster_pcg_fast.cuh):
//      20 = fast cluster kernel (single-exchange recurrence, per-CTA reductions), 1 CTA/SM; 21 = 2 CTAs/SM budget;
//      22 = its timeline build; 24 = fast grid kernel (n = 64: whole GPU on one system); 25 = the same with a two-level exchange
//           (4-CTA clusters: DSMEM inside, L2 between the cluster leaders)
//      26 = fast cluster kernel with packed knot rows (n lanes per row, 16 or 32 rows per CTA): the batch kernels
//      27 = fast batch kernel (gbd_cluster_pcg_fastb.cuh: four rows of one matrix per thread, P-threads / S-threads), 1 CTA/SM;
//      28 = the same, 2 CTAs/SM; 29 = its timeline build; 31 = the same with the near-halo rows of u travelling instead of recomputed
constexpr int MODE_GRID = 4, MODE_FAST = 20, MODE_FAST2 = 21, MODE_FAST_PROF = 22, MODE_FAST_GRID = 24, MODE_FAST_GRID2 = 25, MODE_FAST_BATCH = 26,
              MODE_FAST_B = 27, MODE_FAST_B2 = 28, MODE_FAST_B_PROF = 29, MODE_FAST_BX = 31;
inline bool mode_is_packed(int mode) { return (mode >= 26 && mode <= 28) || mode == 31; }   // per-CTA products parked n per knot row (oracle: lanes = n)
inline bool mode_is_fast(int mode) { return mode >= 20; }
inline bool mode_is_grid(int mode) { return mode == MODE_GRID || mode == MODE_FAST_GRID || mode == MODE_FAST_GRID2; }

struct Variant {
    uint32_t n, N, C;
    int mode;
    bool f64;
    uint32_t nt;
    size_t smem;
    const void *kernel;
    const char *name;        // kernel family as it appears in profiles (ncu prints the full template name)
    size_t ws_words = 0;     // grid kernels: u64 words of packet workspace
    uint32_t grid_cluster = 0;   // grid kernels: CTAs per thread-block cluster (0 / 1: launched without clusters)
    bool prepared = false;
    int resident = -1;       // cluster kernels: clusters of this shape the device can hold (set by prepare)
    bool unusable = false;   // the device cannot place even one cluster of this size: defaults skip the variant
};

void register_exact_v1v2(std::vector<Variant> &v);
void register_exact_v3v5(std::vector<Variant> &v);
void register_exact_v4(std::vector<Variant> &v);
void register_grid(std::vector<Variant> &v);
void register_fast(std::vector<Variant> &v);
void register_fastb(std::vector<Variant> &v);

}  // namespace gbdlib
