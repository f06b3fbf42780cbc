// ref_schur_wrapper.cu -- TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE's own Schur-complement /
// preconditioner assembly (include/pcg/linsys_setup.cuh:621-657, form_schur_system<T>) and step recovery
// (include/common/dz.cuh:125-136, compute_dz<T>), compiled by oracle/Makefile from the reference headers where
// they lie into oracle/_ref/libref_schur.so.  Used by tests/ and tools/ for A/B parity and timing against
// mpcgpu_b200's own kernels; never linked into the product.
#include <cstdint>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include "gpu_pcg.cuh"
#include "settings.cuh"
#include "pcg/linsys_setup.cuh"
#include "common/dz.cuh"

extern "C" {

// exactly the reference call (cooperative launch, grid = knot_points, SCHUR_THREADS threads, default stream)
int ref_form_schur_system_f32(uint32_t n, uint32_t m, uint32_t N, float *d_G, float *d_C, float *d_g, float *d_c, float *d_S,
                              float *d_Pinv, float *d_gamma, float rho)
{
    form_schur_system<float>(n, m, N, d_G, d_C, d_g, d_c, d_S, d_Pinv, d_gamma, rho);
    return (int)cudaGetLastError();
}

int ref_compute_dz_f32(uint32_t n, uint32_t m, uint32_t N, float *d_Ginv, float *d_C, float *d_g, float *d_lambda, float *d_dz)
{
    compute_dz<float>(n, m, N, d_Ginv, d_C, d_g, d_lambda, d_dz);
    return (int)cudaGetLastError();
}

int ref_schur_threads(void) { return SCHUR_THREADS; }
int ref_dz_threads(void) { return DZ_THREADS; }

}  // extern "C"
