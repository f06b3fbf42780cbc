// variants_exact_c.cu -- bit-exact family, generation v4 (self-validating packets polled in shared memory).
#include "gbd_variants.h"
#include "../../include/gbd/gbd_cluster_pcg_v4.cuh"

namespace gbdlib {
using namespace gbd;

template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool PROF = false>
static Variant make_v4()
{
    using K = ClusterPcg4<n, N, C, 0>;
    return Variant{n, N, C, PROF ? 10 : (MINB == 1 ? 7 : 8), false, K::NT, K::SMEM_BYTES,
                   (const void *)pcg_cluster_kernel_v4<n, N, C, MINB, false, PROF, 0>, "gbd::pcg_cluster_kernel_v4"};
}

void register_exact_v4(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_v4<14, 32, 4, 1>(),     make_v4<14, 64, 8, 1>(),     make_v4<14, 128, 8, 1>(),
        make_v4<14, 128, 16, 1>(),   make_v4<14, 256, 16, 1>(),   make_v4<14, 32, 4, 2>(),
        make_v4<14, 32, 4, 1, true>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
