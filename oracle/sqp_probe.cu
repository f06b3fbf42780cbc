// sqp_probe.cu -- TEST INFRASTRUCTURE ONLY.  Calls the REFERENCE's sqpSolvePcg (include/pcg/sqp.cuh:22) a few
// times on the reference trajectory and prints what the linear solver returned in every SQP iteration
// (pcg_iters, max_iter_exit) plus checksums of lambda and xu.  Built twice by oracle/Makefile: against the
// reference's GBD-PCG headers and against include/gbd_dropin -- the two printouts must agree.
//   sqp_probe <traj.csv> <eepos.traj> <tol> <calls> <warm calls at tol 1e-11 / 10000 iterations, as mpcsim.cuh:224-233>
// Build with -DSQP_MAX_TIME_US=1000000000 so that the wall-clock time-box never cuts the SQP loop (determinism).
#include <cstdio>
#include <vector>
#include "mpcsim.cuh"
#include "dynamics/rbd_plant.cuh"
#include "settings.cuh"
#include "utils/experiment.cuh"
#include "gpu_pcg.cuh"

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    setvbuf(stdout, NULL, _IOLBF, 0);
    constexpr uint32_t state_size = grid::NUM_JOINTS * 2, control_size = grid::NUM_JOINTS, knot_points = KNOT_POINTS;
    const linsys_t timestep = .015625;
    auto xu2d = readCSVToVecVec<linsys_t>(argv[1]);
    auto ee2d = readCSVToVecVec<linsys_t>(argv[2]);
    const float tol = atof(argv[3]);
    const int calls = atoi(argv[4]);
    std::vector<linsys_t> xu, ee;
    for (uint32_t e = 0; e < control_size; e++) xu2d[0][e] += 0.05f * (e % 2 ? -1.f : 1.f);   // start off the reference path
    for (uint32_t k = 0; k < knot_points; k++) {
        const uint32_t take = (k + 1 < knot_points) ? state_size + control_size : state_size;
        xu.insert(xu.end(), xu2d[k].begin(), xu2d[k].begin() + take);
        ee.insert(ee.end(), ee2d[k].begin(), ee2d[k].begin() + 6);
    }
    const uint32_t traj_len = (state_size + control_size) * knot_points - control_size;
    linsys_t *d_xu, *d_ee, *d_lambda;
    gpuErrchk(cudaMalloc(&d_xu, traj_len * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_ee, 6 * knot_points * sizeof(linsys_t)));
    gpuErrchk(cudaMalloc(&d_lambda, state_size * knot_points * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpy(d_xu, xu.data(), traj_len * sizeof(linsys_t), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_ee, ee.data(), 6 * knot_points * sizeof(linsys_t), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemset(d_lambda, 0, state_size * knot_points * sizeof(linsys_t)));
    void *d_dynmem = gato_plant::initializeDynamicsConstMem<linsys_t>();
    checkPcgOccupancy<linsys_t>((void *)pcg<linsys_t, state_size, knot_points>, PCG_NUM_THREADS, state_size, knot_points);
    pcg_config<linsys_t> config;
    config.pcg_block = PCG_NUM_THREADS;
    config.pcg_exit_tol = tol;
    config.pcg_max_iter = PCG_MAX_ITER;
    linsys_t rho = 1e-3;
    std::vector<linsys_t> hl(state_size * knot_points), hx(traj_len);
    const int warm = argc > 5 ? atoi(argv[5]) : 0;
#ifdef GBD_GRID_DEBUG
    void *d_dbg;
    gpuErrchk(cudaMalloc(&d_dbg, (2 * 3 * state_size * state_size * knot_points + 2 * state_size * knot_points) * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpyToSymbol(gbd_dropin::g_dbg_ptr, &d_dbg, sizeof(void *)));
    const unsigned delay_us = argc > 6 ? atoi(argv[6]) : 0;
    gpuErrchk(cudaMemcpyToSymbol(gbd_dropin::g_dbg_delay_us, &delay_us, sizeof(unsigned)));
#endif
    config.pcg_exit_tol = 1e-11;
    config.pcg_max_iter = 10000;
    for (int c = 0; c < warm; c++) {
        auto res = sqpSolvePcg<linsys_t>(state_size, control_size, knot_points, timestep, d_ee, d_lambda, d_xu, d_dynmem, config, rho, 1e-3);
        auto iters = std::get<0>(res);
        auto exits = std::get<5>(res);
        gpuErrchk(cudaMemcpy(hl.data(), d_lambda, hl.size() * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        double sl = 0;
        for (auto v : hl) sl += (double)v * v;
        printf("warm %d |lambda|^2 %.9g pcg_iters:", c, sl);
        for (size_t i = 0; i < iters.size(); i++) printf(" %d%s", iters[i], exits[i] ? "*" : "");
        printf("\n");
#ifdef GBD_GRID_DEBUG
        if (c == 0) {
            std::vector<linsys_t> sys(2 * 3 * state_size * state_size * knot_points + 2 * state_size * knot_points);
            gpuErrchk(cudaMemcpy(sys.data(), d_dbg, sys.size() * sizeof(linsys_t), cudaMemcpyDeviceToHost));
            FILE *f = fopen("dbg_first_system.bin", "wb");
            fwrite(sys.data(), sizeof(linsys_t), sys.size(), f);
            fclose(f);
        }
#endif
        gpuErrchk(cudaMemcpy(d_xu, xu.data(), traj_len * sizeof(linsys_t), cudaMemcpyHostToDevice));
    }
    rho = 1e-3;
    config.pcg_exit_tol = tol;
    config.pcg_max_iter = PCG_MAX_ITER;
    for (int c = 0; c < calls; c++) {
        auto res = sqpSolvePcg<linsys_t>(state_size, control_size, knot_points, timestep, d_ee, d_lambda, d_xu, d_dynmem, config, rho, 1e-3);
        auto iters = std::get<0>(res);
        auto times = std::get<1>(res);
        auto exits = std::get<5>(res);
        gpuErrchk(cudaMemcpy(hl.data(), d_lambda, hl.size() * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        gpuErrchk(cudaMemcpy(hx.data(), d_xu, hx.size() * sizeof(linsys_t), cudaMemcpyDeviceToHost));
        double sl = 0, sx = 0;
        for (auto v : hl) sl += (double)v * v;
        for (auto v : hx) sx += (double)v * v;
        printf("call %d sqp_iters %u rho %.6g |lambda|^2 %.9g |xu|^2 %.9g\n  pcg_iters:", c, (unsigned)std::get<3>(res), (double)rho, sl, sx);
        for (size_t i = 0; i < iters.size(); i++) printf(" %d%s", iters[i], exits[i] ? "*" : "");
        double tsum = 0;
        for (auto t : times) tsum += t;
        printf("\n  linsys_us mean %.1f\n", times.empty() ? 0.0 : tsum / times.size());
    }
    return 0;
}
