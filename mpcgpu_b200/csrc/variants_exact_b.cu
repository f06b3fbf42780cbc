// variants_exact_b.cu -- bit-exact family, generations v3 (two matrix rows per thread) and v5 (v3's mapping + packets).
#include "gbd_variants.h"
#include "../../include/gbd/gbd_cluster_pcg_v5.cuh"

namespace gbdlib {
using namespace gbd;

template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
static Variant make_v3()
{
    using K = ClusterPcg3<n, N, C, true>;
    return Variant{n, N, C, MINB == 1 ? 5 : 6, false, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel_v3<n, N, C, MINB>,
                   "gbd::pcg_cluster_kernel_v3"};
}
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
static Variant make_v5()
{
    using K = ClusterPcg5<n, N, C, true>;
    return Variant{n, N, C, MINB == 1 ? 11 : 12, false, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel_v5<n, N, C, MINB>,
                   "gbd::pcg_cluster_kernel_v5"};
}

void register_exact_v3v5(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_v3<14, 512, 16, 1>(),   make_v3<14, 128, 8, 2>(),   make_v3<14, 128, 8, 1>(),
        make_v3<14, 32, 1, 2>(),     make_v3<14, 64, 2, 2>(),    make_v3<14, 256, 16, 1>(),
        make_v3<14, 16, 2, 1>(),     make_v3<6, 12, 3, 1>(),
        make_v5<14, 128, 4, 1>(),    make_v5<14, 128, 8, 1>(),   make_v5<14, 32, 2, 1>(),
        make_v5<14, 64, 4, 1>(),     make_v5<14, 256, 8, 1>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
