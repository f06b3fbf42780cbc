// ref_capture.cu -- dumps real IIWA Schur systems produced by the REFERENCE's own assembly kernels.
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile against the reference headers where they lie
// (one binary per KNOT_POINTS, like every reference build) into oracle/_ref/.
//
// It reproduces the first steps of sqpSolvePcg (include/pcg/sqp.cuh:94-219) on the reference trajectory
// (examples/trajfiles/0_0_*, loaded as examples/track_iiwa_pcg.cu:84-112 does): generate_kkt_submatrices
// then form_schur_system at rho = 1e-3, and writes (S, Pinv, gamma) of each requested system as raw
// float32.  d_S / d_Pinv are pre-filled with 0xFF bytes (NaN) so the pad tiles the reference never
// writes are visibly garbage.
//   ref_capture <traj.csv> <eepos.traj> <out.bin> <knot offset> <count> <perturb 0|1> [mode bits]
// mode bits (debug aid): 1 = do not pre-fill S/Pinv with 0xFF; 2 = run compute_merit first as sqp.cuh:173 does;
// 4 = create the cuBLAS handle and 8 streams first as sqp.cuh:64-72 does; 8 = run the PCG kernel after the assembly
// exactly as sqp.cuh:230 launches it and append lambda (n*N floats), iters, flag to the record;
// 16 = also write <out.bin>.kkt: G, C, g, c as generate_kkt_submatrices left them (the inputs of form_schur_system),
// then G after form_schur_system (the block inverses), and -- with bit 8 -- dz from the reference's compute_dz on the
// PCG solution (golden vectors for rows f1 / f2, tools/make_golden_schur.py).
// count > 1 with perturb = 1 builds BASELINE config 4's batch: system i = the window at <offset> plus
// N(0, 0.05^2) on q, N(0, 0.01^2) on qd, N(0, 1) on u, std::mt19937_64(1234 + i).
// perturb = 2 slides the window instead: system i = the window at <offset> + i * <stride> (argv[8], default 1), no noise --
// the ring of distinct real systems bench.py solves (what successive MPC steps of examples/track_iiwa_pcg.cu assemble).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "mpcsim.cuh"
#include "dynamics/rbd_plant.cuh"
#include "settings.cuh"
#include "utils/experiment.cuh"
#include "gpu_pcg.cuh"

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: ref_capture traj.csv eepos.traj out.bin offset count perturb\n"); return 2; }
    constexpr uint32_t state_size = grid::NUM_JOINTS * 2, control_size = grid::NUM_JOINTS, knot_points = KNOT_POINTS;
    const linsys_t timestep = .015625;
    const uint32_t offset = atoi(argv[4]), count = atoi(argv[5]);
    const bool perturb = atoi(argv[6]) == 1, slide = atoi(argv[6]) == 2;
    const uint32_t stride = argc > 8 ? atoi(argv[8]) : 1;
    auto xu2d = readCSVToVecVec<linsys_t>(argv[1]);
    auto ee2d = readCSVToVecVec<linsys_t>(argv[2]);
    const uint32_t last_offset = offset + (slide ? (count - 1) * stride : 0);
    if (xu2d.size() < last_offset + knot_points || ee2d.size() < last_offset + knot_points) { fprintf(stderr, "trajectory too short\n"); return 3; }
    std::vector<linsys_t> xu0, ee;
    auto load_window = [&](uint32_t off) {
        xu0.clear();
        ee.clear();
        for (uint32_t k = 0; k < knot_points; k++) {
            const auto &row = xu2d[off + k];
            const uint32_t take = (k + 1 < knot_points) ? state_size + control_size : state_size;
            xu0.insert(xu0.end(), row.begin(), row.begin() + take);
            ee.insert(ee.end(), ee2d[off + k].begin(), ee2d[off + k].begin() + 6);
        }
    };
    load_window(offset);
    const uint32_t traj_len = (state_size + control_size) * knot_points - control_size;
    const uint32_t states_sq = state_size * state_size, controls_sq = control_size * control_size;
    const size_t G_bytes = ((states_sq + controls_sq) * knot_points - controls_sq) * sizeof(linsys_t);
    const size_t C_bytes = (states_sq + state_size * control_size) * (knot_points - 1) * sizeof(linsys_t);
    const size_t g_bytes = ((state_size + control_size) * knot_points - control_size) * sizeof(linsys_t);
    const size_t c_bytes = state_size * knot_points * sizeof(linsys_t);
    const size_t mat = 3 * (size_t)states_sq * knot_points, vec = (size_t)state_size * knot_points;

    linsys_t *d_G, *d_C, *d_g, *d_c, *d_S, *d_Pinv, *d_gamma, *d_xu, *d_xs, *d_ee;
    gpuErrchk(cudaMalloc(&d_G, G_bytes)); gpuErrchk(cudaMalloc(&d_C, C_bytes));
    gpuErrchk(cudaMalloc(&d_g, g_bytes)); gpuErrchk(cudaMalloc(&d_c, c_bytes));
    gpuErrchk(cudaMalloc(&d_S, mat * sizeof(linsys_t))); gpuErrchk(cudaMalloc(&d_Pinv, mat * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_gamma, vec * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_xu, traj_len * sizeof(linsys_t))); gpuErrchk(cudaMalloc(&d_xs, state_size * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_ee, 6 * knot_points * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpy(d_ee, ee.data(), 6 * knot_points * sizeof(linsys_t), cudaMemcpyHostToDevice));
    void *d_dynmem = gato_plant::initializeDynamicsConstMem<linsys_t>();
    linsys_t rho = 1e-3;
    const int mode = argc > 7 ? atoi(argv[7]) : 0;
    cudaStream_t streams[8];
    cublasHandle_t handle;
    if (mode & 4) {
        for (int i = 0; i < 8; i++) cudaStreamCreate(&streams[i]);
        if (cublasCreate(&handle) != CUBLAS_STATUS_SUCCESS) return 13;
    }
    linsys_t *d_merit, *d_lambda, *d_r, *d_p, *d_v, *d_e;
    uint32_t *d_it;
    bool *d_fl;
    gpuErrchk(cudaMalloc(&d_merit, sizeof(linsys_t)));
    gpuErrchk(cudaMemset(d_merit, 0, sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_lambda, vec * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_r, vec * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_p, vec * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_v, knot_points * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_e, knot_points * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_it, sizeof(uint32_t)));
    gpuErrchk(cudaMalloc(&d_fl, sizeof(bool)));

    FILE *f = fopen(argv[3], "wb");
    FILE *fk = (mode & 16) ? fopen((std::string(argv[3]) + ".kkt").c_str(), "wb") : nullptr;
    auto dump_dev = [&](FILE *fp, const linsys_t *d, size_t bytes) {
        std::vector<char> h(bytes);
        gpuErrchk(cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost));
        fwrite(h.data(), 1, bytes, fp);
    };
    linsys_t *d_dz = nullptr;
    if (fk) gpuErrchk(cudaMalloc(&d_dz, g_bytes));
    std::vector<linsys_t> hS(mat), hP(mat), hg(vec);
    for (uint32_t i = 0; i < count; i++) {
        if (slide && i > 0) {
            load_window(offset + i * stride);
            gpuErrchk(cudaMemcpy(d_ee, ee.data(), 6 * knot_points * sizeof(linsys_t), cudaMemcpyHostToDevice));
        }
        std::vector<linsys_t> xu = xu0;
        if (perturb) {
            std::mt19937_64 gen(1234 + i);
            std::normal_distribution<double> nq(0.0, 0.05), nqd(0.0, 0.01), nu(0.0, 1.0);
            for (uint32_t k = 0; k < knot_points; k++) {
                linsys_t *row = xu.data() + (size_t)k * (state_size + control_size);
                for (uint32_t e = 0; e < control_size; e++) row[e] += (linsys_t)nq(gen);
                for (uint32_t e = control_size; e < state_size; e++) row[e] += (linsys_t)nqd(gen);
                if (k + 1 < knot_points)
                    for (uint32_t e = 0; e < control_size; e++) row[state_size + e] += (linsys_t)nu(gen);
            }
        }
        gpuErrchk(cudaMemcpy(d_xu, xu.data(), traj_len * sizeof(linsys_t), cudaMemcpyHostToDevice));
        gpuErrchk(cudaMemcpy(d_xs, xu.data(), state_size * sizeof(linsys_t), cudaMemcpyHostToDevice));
        if (!(mode & 1)) {
            gpuErrchk(cudaMemset(d_S, 0xFF, mat * sizeof(linsys_t)));
            gpuErrchk(cudaMemset(d_Pinv, 0xFF, mat * sizeof(linsys_t)));
        }
        if (mode & 2) {
            compute_merit<linsys_t><<<knot_points, MERIT_THREADS, get_merit_smem_size<linsys_t>(state_size, control_size)>>>(
                state_size, control_size, knot_points, d_xu, d_ee, static_cast<linsys_t>(10), timestep, d_dynmem, d_merit);
            gpuErrchk(cudaPeekAtLastError());
        }
        generate_kkt_submatrices<linsys_t><<<knot_points, KKT_THREADS, 2 * get_kkt_smem_size<linsys_t>(state_size, control_size)>>>(
            state_size, control_size, knot_points, d_G, d_C, d_g, d_c, d_dynmem, timestep, d_ee, d_xs, d_xu);
        gpuErrchk(cudaPeekAtLastError());
        if (fk) {
            gpuErrchk(cudaDeviceSynchronize());
            dump_dev(fk, d_G, G_bytes); dump_dev(fk, d_C, C_bytes); dump_dev(fk, d_g, g_bytes); dump_dev(fk, d_c, c_bytes);
        }
        form_schur_system<linsys_t>(state_size, control_size, knot_points, d_G, d_C, d_g, d_c, d_S, d_Pinv, d_gamma, rho);
        gpuErrchk(cudaPeekAtLastError());
        gpuErrchk(cudaDeviceSynchronize());
        gpuErrchk(cudaMemcpy(hS.data(), d_S, mat * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        gpuErrchk(cudaMemcpy(hP.data(), d_Pinv, mat * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        gpuErrchk(cudaMemcpy(hg.data(), d_gamma, vec * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        fwrite(hS.data(), sizeof(linsys_t), mat, f);
        fwrite(hP.data(), sizeof(linsys_t), mat, f);
        fwrite(hg.data(), sizeof(linsys_t), vec, f);
        if (fk) dump_dev(fk, d_G, G_bytes);
        if (mode & 8) {
            pcg_config<linsys_t> config;
            config.pcg_exit_tol = 1e-4;
            config.pcg_max_iter = PCG_MAX_ITER;
            gpuErrchk(cudaMemset(d_lambda, 0, vec * sizeof(linsys_t)));
            void *pcg_kernel = (void *)pcg<linsys_t, STATE_SIZE, KNOT_POINTS>;
            void *args[] = {(void *)&d_S, (void *)&d_Pinv, (void *)&d_gamma, (void *)&d_lambda, (void *)&d_r, (void *)&d_p, (void *)&d_v,
                            (void *)&d_e, (void *)&d_it, (void *)&d_fl, (void *)&config.pcg_max_iter, (void *)&config.pcg_exit_tol};
            gpuErrchk(cudaLaunchCooperativeKernel(pcg_kernel, knot_points, PCG_NUM_THREADS, args, pcgSharedMemSize<linsys_t>(state_size, knot_points)));
            uint32_t it = 0;
            bool fl = false;
            gpuErrchk(cudaMemcpy(&it, d_it, sizeof(uint32_t), cudaMemcpyDeviceToHost));
            gpuErrchk(cudaMemcpy(&fl, d_fl, sizeof(bool), cudaMemcpyDeviceToHost));
            gpuErrchk(cudaMemcpy(hg.data(), d_lambda, vec * sizeof(linsys_t), cudaMemcpyDeviceToHost));
            fwrite(hg.data(), sizeof(linsys_t), vec, f);
            uint32_t tail[2] = {it, (uint32_t)fl};
            fwrite(tail, sizeof(uint32_t), 2, f);
            printf("  pcg: iters %u max_iter_exit %d\n", it, (int)fl);
            if (fk) {
                compute_dz<linsys_t>(state_size, control_size, knot_points, d_G, d_C, d_g, d_lambda, d_dz);
                gpuErrchk(cudaDeviceSynchronize());
                dump_dev(fk, d_dz, g_bytes);
            }
        }
    }
    fclose(f);
    if (fk) fclose(fk);
    printf("captured %u system(s): n=%u N=%u offset=%u perturb=%d slide=%d stride=%u -> %s\n", count, state_size, knot_points, offset,
           (int)perturb, (int)slide, stride, argv[3]);
    return 0;
}
