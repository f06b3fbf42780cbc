// closed_loop.cu -- TEST INFRASTRUCTURE ONLY.  One closed-loop MPC experiment of the REFERENCE, exactly what the inner loop of
// examples/track_iiwa_pcg.cu:84-117 does for one pcg_exit_tol: load examples/trajfiles/0_0_*, call simulateMPC
// (include/mpcsim.cuh:146-149: the reference's SQP loop sqpSolvePcg around the pcg<> kernel, its plant simulation, shifting and
// tracking-error bookkeeping, all unchanged) and report what it returns.  Built several times by oracle/Makefile (target
// closed_loop): against the reference's GBD-PCG headers, against include/gbd_dropin (bit-exact bodies) and against
// include/gbd_dropin with -DGBD_DROPIN_FAST=1 (tolerance-parity body); with -DTIME_LINSYS=0 for behaviour (SQP iterations per control
// step) and =1 for timing (the reference's own linsys stopwatch).  -DSQP_MAX_TIME_US=1000000000 takes the wall clock out of the SQP
// exit rule (include/pcg/sqp.cuh:161-166) and CONST_UPDATE_FREQ=1 keeps the simulated period fixed, so runs are deterministic.
//   closed_loop <traj.csv> <eepos.traj> <pcg_exit_tol> <trajectory knots to track (<= rows of the file)> <out.bin>
// out.bin: u32 count_a, count_b; then count_a doubles (linsys us) or u32 (SQP iterations per control step), then count_b float tracking errors
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "mpcsim.cuh"
#include "dynamics/rbd_plant.cuh"
#include "settings.cuh"
#include "utils/experiment.cuh"
#include "gpu_pcg.cuh"

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: closed_loop traj.csv eepos.traj tol knots out.bin\n"); return 2; }
    constexpr uint32_t state_size = grid::NUM_JOINTS * 2, control_size = grid::NUM_JOINTS, knot_points = KNOT_POINTS;
    const linsys_t timestep = .015625;
    const float tol = (float)atof(argv[3]);
    auto xu2d = readCSVToVecVec<linsys_t>(argv[1]);
    auto ee2d = readCSVToVecVec<linsys_t>(argv[2]);
    size_t rows = (size_t)atoi(argv[4]);
    if (rows > ee2d.size()) rows = ee2d.size();
    if (rows < knot_points) { fprintf(stderr, "need at least %u knots\n", knot_points); return 3; }
    xu2d.resize(rows);
    ee2d.resize(rows);
    checkPcgOccupancy<linsys_t>((void *)pcg<linsys_t, state_size, knot_points>, PCG_NUM_THREADS, state_size, knot_points);
    std::vector<linsys_t> h_ee, h_xu;
    for (const auto &v : ee2d) h_ee.insert(h_ee.end(), v.begin(), v.end());
    for (const auto &v : xu2d) h_xu.insert(h_xu.end(), v.begin(), v.end());
    linsys_t *d_ee, *d_xu, *d_xs;
    gpuErrchk(cudaMalloc(&d_ee, h_ee.size() * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpy(d_ee, h_ee.data(), h_ee.size() * sizeof(linsys_t), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMalloc(&d_xu, h_xu.size() * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpy(d_xu, h_xu.data(), h_xu.size() * sizeof(linsys_t), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMalloc(&d_xs, state_size * sizeof(linsys_t)));
    gpuErrchk(cudaMemcpy(d_xs, h_xu.data(), state_size * sizeof(linsys_t), cudaMemcpyHostToDevice));

    auto stats = simulateMPC<linsys_t, toplevel_return_type>(state_size, control_size, knot_points, (uint32_t)rows, timestep, d_ee, d_xu, d_xs,
                                                             0, 0, 0, tol, std::string("tmp/results/closed_loop"));
    std::vector<toplevel_return_type> top = std::get<0>(stats);
    std::vector<linsys_t> err = std::get<1>(stats);
    const double err_mean = err.empty() ? 0.0 : std::accumulate(err.begin(), err.end(), 0.0) / err.size();
    double top_sum = 0;
    for (auto v : top) top_sum += (double)v;
    printf("knots %u tol %g tracked %zu | tracking error mean %.9g final %.9g | %s count %zu sum %.9g mean %.6g\n", knot_points, (double)tol, err.size(),
           err_mean, (double)std::get<2>(stats), TIME_LINSYS ? "linsys_us" : "sqp_iters", top.size(), top_sum, top.empty() ? 0.0 : top_sum / top.size());
    FILE *f = fopen(argv[5], "wb");
    if (!f) return 4;
    uint32_t cnt[2] = {(uint32_t)top.size(), (uint32_t)err.size()};
    fwrite(cnt, sizeof(uint32_t), 2, f);
    fwrite(top.data(), sizeof(toplevel_return_type), top.size(), f);
    fwrite(err.data(), sizeof(linsys_t), err.size(), f);
    fclose(f);
    return 0;
}
