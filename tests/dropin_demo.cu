// dropin_demo.cu -- exercises the header-only drop-in set (include/gbd_dropin) exactly the way the
// reference's SQP loop uses GBD-PCG (include/pcg/sqp.cuh:116-151,230-232): kernel pointer
// pcg<T,STATE_SIZE,KNOT_POINTS>, positional 12-argument array, pcgSharedMemSize<T>(), cooperative launch
// with grid = knot_points and a caller-chosen block size, two blocking D2H reads.
//   dropin_demo <in.bin> <out.bin> <max_iter> <tol> <block> [repeat]
// in.bin : float32 S[3n^2N] Pinv[3n^2N] gamma[nN] lambda0[nN];  out.bin : float32 lambda[nN] r[nN] p[nN], u32 iters, u32 flag
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gpu_pcg.cuh"

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    const uint32_t n = STATE_SIZE, N = KNOT_POINTS;
    const size_t mat = (size_t)3 * n * n * N, vec = (size_t)n * N;
    std::vector<float> h(2 * mat + 2 * vec);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(h.data(), sizeof(float), h.size(), f) != h.size()) return 3;
    fclose(f);
    pcg_config<float> config;
    config.pcg_max_iter = (uint32_t)atoi(argv[3]);
    config.pcg_exit_tol = (float)atof(argv[4]);
    const unsigned block = (unsigned)atoi(argv[5]);
    const int repeat = argc > 6 ? atoi(argv[6]) : 1;

    if (!checkPcgOccupancy<float>((void *)pcg<float, STATE_SIZE, KNOT_POINTS>, dim3(block), n, N)) return 4;
    float *d_S, *d_Pinv, *d_gamma, *d_lambda, *d_r, *d_p, *d_v_temp, *d_eta_new_temp;
    gpuErrchk(cudaMalloc(&d_S, mat * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_Pinv, mat * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_gamma, vec * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_lambda, vec * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_r, vec * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_p, vec * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_v_temp, N * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_eta_new_temp, N * sizeof(float)));
    gpuErrchk(cudaMemcpy(d_S, h.data(), mat * sizeof(float), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_Pinv, h.data() + mat, mat * sizeof(float), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_gamma, h.data() + 2 * mat, vec * sizeof(float), cudaMemcpyHostToDevice));

    void *pcg_kernel = (void *)pcg<float, STATE_SIZE, KNOT_POINTS>;
    uint32_t pcg_iters = 0, *d_pcg_iters;
    bool pcg_exit = false, *d_pcg_exit;
    gpuErrchk(cudaMalloc(&d_pcg_iters, sizeof(uint32_t)));
    gpuErrchk(cudaMalloc(&d_pcg_exit, sizeof(bool)));
    void *pcgKernelArgs[] = {(void *)&d_S, (void *)&d_Pinv, (void *)&d_gamma, (void *)&d_lambda, (void *)&d_r, (void *)&d_p,
                             (void *)&d_v_temp, (void *)&d_eta_new_temp, (void *)&d_pcg_iters, (void *)&d_pcg_exit,
                             (void *)&config.pcg_max_iter, (void *)&config.pcg_exit_tol};
    const size_t smem = pcgSharedMemSize<float>(n, N);
    if (smem != gbd_dropin::Shape<float, STATE_SIZE, KNOT_POINTS>::SMEM_BYTES) { fprintf(stderr, "smem size mismatch\n"); return 5; }

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total_ms = 0.f;
    for (int rep = 0; rep < repeat; ++rep) {
        gpuErrchk(cudaMemcpy(d_lambda, h.data() + 2 * mat + vec, vec * sizeof(float), cudaMemcpyHostToDevice));
        cudaEventRecord(e0);
        gpuErrchk(cudaLaunchCooperativeKernel(pcg_kernel, N, block, pcgKernelArgs, smem));
        cudaEventRecord(e1);
        gpuErrchk(cudaMemcpy(&pcg_iters, d_pcg_iters, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        gpuErrchk(cudaMemcpy(&pcg_exit, d_pcg_exit, sizeof(bool), cudaMemcpyDeviceToHost));
        gpuErrchk(cudaDeviceSynchronize());
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= repeat / 2) total_ms += ms;
    }
    std::vector<float> out(3 * vec);
    gpuErrchk(cudaMemcpy(out.data(), d_lambda, vec * sizeof(float), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaMemcpy(out.data() + vec, d_r, vec * sizeof(float), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaMemcpy(out.data() + 2 * vec, d_p, vec * sizeof(float), cudaMemcpyDeviceToHost));
    f = fopen(argv[2], "wb");
    fwrite(out.data(), sizeof(float), out.size(), f);
    uint32_t tail[2] = {pcg_iters, (uint32_t)pcg_exit};
    fwrite(tail, sizeof(uint32_t), 2, f);
    fclose(f);
    printf("iters %u exit %d kernel_us %.2f\n", pcg_iters, (int)pcg_exit, 1e3f * total_ms / (repeat - repeat / 2));
    return 0;
}
