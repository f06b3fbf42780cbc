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

void register_fastb(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_fastb<14, 128, 4, 1>(), make_fastb<14, 128, 8, 2>(), make_fastb<14, 128, 8, 1>(), make_fastb<14, 32, 1, 1>(),
        make_fastb<14, 32, 2, 2>(),  make_fastb<14, 64, 2, 1>(),  make_fastb<14, 64, 4, 2>(),  make_fastb<14, 256, 8, 1>(),
        make_fastb<14, 512, 16, 1>(), make_fastb<6, 16, 2, 1>(),
        make_fastb<14, 128, 4, 1, true>(), make_fastb<14, 32, 1, 1, true>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
