// variants_fast.cu -- tolerance-parity family (include/gbd/gbd_cluster_pcg_fast.cuh).
#include "gbd_variants.h"
#include "../../include/gbd/gbd_cluster_pcg_fast.cuh"

namespace gbdlib {
using namespace gbd;

template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
static Variant make_fast()
{
    using K = ClusterPcgFast<n, N, C>;
    return Variant{n, N, C, PROF ? MODE_FAST_PROF : (MINB == 1 ? MODE_FAST : MODE_FAST2), false, K::NT, K::SMEM_BYTES,
                   (const void *)pcg_cluster_kernel_fast<n, N, C, MINB, PROF>, "gbd::pcg_cluster_kernel_fast"};
}
// batch kernels: knot rows packed n lanes each (no idle lanes), 16 or 32 knot rows per CTA -- few CTAs per system, so many
// systems in flight; for n = 14 and 32 rows a CTA is 15 warps, at most 4 per scheduler partition = 128 registers per thread
template <uint32_t n, uint32_t N, uint32_t C>
static Variant make_fast_packed()
{
    using K = ClusterPcgFast<n, N, C, n>;
    return Variant{n, N, C, MODE_FAST_BATCH, false, K::NT, K::SMEM_BYTES,
                   (const void *)pcg_cluster_kernel_fast<n, N, C, 1, false, n>, "gbd::pcg_cluster_kernel_fast(packed)"};
}

void register_fast(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_fast<14, 128, 16, 1>(),  make_fast<14, 128, 8, 1>(),   make_fast<14, 32, 4, 1>(),
        make_fast<14, 32, 2, 1>(),    make_fast<14, 64, 8, 1>(),    make_fast<14, 64, 4, 1>(),
        make_fast<14, 256, 16, 1>(),  make_fast<14, 16, 4, 1>(),    make_fast<14, 8, 2, 1>(),
        make_fast<6, 12, 3, 1>(),     make_fast<6, 12, 2, 1>(),
        make_fast<14, 128, 16, 1, true>(), make_fast<14, 128, 8, 1, true>(), make_fast<14, 32, 4, 1, true>(),
        // packed rows (n lanes per knot row, GL = n): measured slower than the batch kernel of gbd_cluster_pcg_fastb.cuh on every shape
        // (profiles/r02c_ab.log: 182 K vs 230 K systems/s at 1024 x N = 128); one build kept so the code path stays tested
        make_fast_packed<14, 32, 2>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
