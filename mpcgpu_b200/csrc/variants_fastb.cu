// variants_fastb.cu -- tolerance-parity family, throughput (batch) kernels (include/gbd/gbd_cluster_pcg_fastb.cuh).
#include "gbd_variants.h"
#include "../../include/gbd/gbd_cluster_pcg_fastb.cuh"

namespace gbdlib {
using namespace gbd;

template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
static Variant make_fastb()
{
    using K = ClusterPcgFastB<n, N, C>;
    return Variant{n, N, C, PROF ? MODE_FAST_B_PROF : (MINB == 1 ? MODE_FAST_B : MODE_FAST_B2), false, K::NT, K::SMEM_BYTES,
                   (const void *)pcg_cluster_kernel_fastb<n, N, C, MINB, PROF>, "gbd::pcg_cluster_kernel_fastb"};
}

// the same kernel with the near-halo rows of u travelling from the neighbours instead of being recomputed (UX)
template <uint32_t n, uint32_t N, uint32_t C>
static Variant make_fastbx()
{
    using K = ClusterPcgFastB<n, N, C, true>;
    return Variant{n, N, C, MODE_FAST_BX, false, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel_fastb<n, N, C, 1, false, true>,
                   "gbd::pcg_cluster_kernel_fastb(u travels)"};
}

void register_fastb(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_fastb<14, 128, 4, 1>(), make_fastb<14, 128, 8, 2>(), make_fastb<14, 128, 8, 1>(), make_fastb<14, 32, 1, 1>(),
        make_fastb<14, 32, 2, 2>(),  make_fastb<14, 64, 2, 1>(),  make_fastb<14, 64, 4, 2>(),  make_fastb<14, 256, 8, 1>(),
        make_fastb<14, 512, 16, 1>(), make_fastb<6, 16, 2, 1>(),
        make_fastb<14, 128, 4, 1, true>(), make_fastb<14, 32, 1, 1, true>(),
        // u travels (UX): measured 10 % slower than recomputing it on every shape (profiles/r02_ab_fastb_u_travels.log); one build kept
        make_fastbx<14, 128, 4>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
