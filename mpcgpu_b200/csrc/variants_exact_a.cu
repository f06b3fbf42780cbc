// variants_exact_a.cu -- bit-exact family, generations v1 (barrier.cluster, generic in n and dtype) and v2 (st.async + mbarrier).
#include "gbd_variants.h"
#include "../../include/gbd/gbd_cluster_pcg_v2.cuh"

namespace gbdlib {
using namespace gbd;

template <typename T, uint32_t n, uint32_t N, uint32_t C, bool REGS>
static Variant make_v1()
{
    using K = ClusterPcg<T, n, N, C, REGS>;
    return Variant{n, N, C, REGS ? 1 : 0, sizeof(T) == 8, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel<T, n, N, C, REGS>,
                   "gbd::pcg_cluster_kernel"};
}
template <typename T, uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
static Variant make_v2()
{
    using K = ClusterPcg2<T, n, N, C>;
    return Variant{n, N, C, MINB == 1 ? 2 : 3, sizeof(T) == 8, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel_v2<T, n, N, C, MINB>,
                   "gbd::pcg_cluster_kernel_v2"};
}
template <uint32_t n, uint32_t N, uint32_t C>
static Variant make_v2_prof()
{
    using K = ClusterPcg2<float, n, N, C>;
    return Variant{n, N, C, 14, false, K::NT, K::SMEM_BYTES, (const void *)pcg_cluster_kernel_v2<float, n, N, C, 1, true>,
                   "gbd::pcg_cluster_kernel_v2"};
}

// IIWA (n = 14) at the reference's horizons (include/common/settings.cuh:123-138) plus small systems for tests
// (n = 2, N = 3 is the GBD-PCG demo, GBD-PCG/examples/pcg_solve.cu:14-25)
void register_exact_v1v2(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_v2<float, 14, 128, 16, 1>(),          make_v2<float, 14, 128, 8, 1>(),
        make_v2<float, 14, 128, 8, 2>(),           make_v2<float, 14, 32, 4, 1>(),
        make_v2<float, 14, 64, 8, 1>(),            make_v2<float, 14, 256, 16, 1>(),
        make_v2<float, 14, 256, 8, 1>(),           make_v2<float, 14, 512, 16, 1>(),
        make_v2<float, 14, 16, 4, 1>(),            make_v2<float, 14, 8, 8, 1>(),
        make_v2<float, 6, 12, 4, 1>(),             make_v2<float, 6, 12, 1, 2>(),
        make_v2<float, 2, 3, 1, 1>(),              make_v2<float, 2, 3, 3, 1>(),
        make_v2_prof<14, 128, 16>(),
        // v1 (gbd_cluster_pcg.cuh) is kept for fp64 only, where it is the one kernel; its fp32 builds were retired in round 2
        make_v1<double, 14, 128, 8, false>(),      make_v1<double, 14, 32, 8, false>(),
        make_v1<double, 6, 12, 4, false>(),        make_v1<double, 2, 3, 1, false>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
