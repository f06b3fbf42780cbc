// variants_grid.cu -- the whole GPU on one system (n = 64, N = 256 and the fallback behind every shape): packets through L2.
#include "gbd_variants.h"
#include "../../include/gbd/gbd_grid_pcg.cuh"
#include "../../include/gbd/gbd_grid_pcg_fast.cuh"

namespace gbdlib {
using namespace gbd;

template <typename T, uint32_t n, uint32_t N, uint32_t R>
static Variant make_grid()
{
    using K = GridPcg<T, n, N, R>;
    Variant v{n, N, K::CTAS, MODE_GRID, sizeof(T) == 8, K::NT_MIN < 128 ? 128 : K::NT_MIN, K::SMEM_BYTES,
              (const void *)pcg_grid_kernel<T, n, N, R>, "gbd::pcg_grid_kernel"};
    v.ws_words = K::WS_WORDS;
    return v;
}

// tolerance-parity grid kernel (n a multiple of 32: BASELINE config 5 and small shapes of the same kind for the tests)
template <uint32_t n, uint32_t N, uint32_t R, uint32_t CL = 1>
static Variant make_grid_fast()
{
    using K = GridPcgFast<n, N, R, CL>;
    Variant v{n, N, K::CTAS, CL > 1 ? MODE_FAST_GRID2 : MODE_FAST_GRID, false, K::NT, K::SMEM_BYTES, (const void *)pcg_grid_kernel_fast<n, N, R, CL>,
              CL > 1 ? "gbd::pcg_grid_kernel_fast(2-level)" : "gbd::pcg_grid_kernel_fast"};
    v.ws_words = K::WS_WORDS;
    v.grid_cluster = CL;
    return v;
}

void register_grid(std::vector<Variant> &v)
{
    const Variant list[] = {
        make_grid<float, 64, 256, 2>(), make_grid<float, 14, 512, 4>(), make_grid<float, 14, 128, 1>(),
        make_grid<float, 14, 32, 1>(),  make_grid<float, 14, 256, 2>(), make_grid<float, 6, 12, 1>(),
        make_grid<float, 2, 3, 1>(),    make_grid<double, 14, 32, 1>(),
        make_grid<float, 32, 8, 2>(),   make_grid<float, 64, 16, 2>(),  make_grid<float, 32, 32, 2>(),
        make_grid_fast<64, 256, 2, 4>(), make_grid_fast<32, 32, 2, 4>(),
        make_grid_fast<64, 256, 2>(),   make_grid_fast<32, 8, 2>(),     make_grid_fast<64, 16, 2>(),
    };
    for (const Variant &x : list) v.push_back(x);
}
}  // namespace gbdlib
