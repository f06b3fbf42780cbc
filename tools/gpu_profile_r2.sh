#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one full capture each of the default single-solve kernel (N = 128 and
# N = 32) and of the default batch kernel (256 systems, N = 128).  Raw pages are exported HERE afterwards by tools/ncu_summarize.py.
mkdir -p gpurun_out
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv \
   --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 100 --warmup 5 --no-cpu --prewarm 0.02 --batched-steps 2 --ring 64 \
   --no-refgpu --no-configs > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster_kernel_fast -s 20 -c 1 \
   -f -o gpurun_out/prof_r02_single128 env BATCH=1 KNOTS=128 python tools/one_solve.py > gpurun_out/prof_r02_single128.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster_kernel_fast -s 20 -c 1 \
   -f -o gpurun_out/prof_r02_single32 env BATCH=1 KNOTS=32 CAP=173 python tools/one_solve.py > gpurun_out/prof_r02_single32.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:fastb -s 1 -c 1 \
   -f -o gpurun_out/prof_r02_batched256 env BATCH=256 SINGLES=0 KNOTS=128 python tools/one_solve.py > gpurun_out/prof_r02_batched256.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_grid_kernel_fast -s 3 -c 1 \
   -f -o gpurun_out/prof_r02_cfg5 env BATCH=1 STATE=64 KNOTS=256 CAP=200 TOL=1e-6 SINGLES=6 python tools/one_solve.py > gpurun_out/prof_r02_cfg5.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench.csv
for f in single128 single32 batched256 cfg5; do tail -n 2 gpurun_out/prof_r02_$f.log; done
