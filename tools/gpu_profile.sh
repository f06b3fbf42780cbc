#!/bin/bash
# ncu evidence: launch list of the bench command + one full capture each of the single and batched kernel.
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 100 --warmup 5 --no-cpu --prewarm 0.02 --batched-steps 2 --ring 64 \
   > gpurun_out/bench_under_ncu.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster -s 20 -c 1 \
   -f -o gpurun_out/prof_single env BATCH=1 python tools/one_solve.py > gpurun_out/prof_single.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster -s 25 -c 1 \
   -f -o gpurun_out/prof_batched env BATCH=256 python tools/one_solve.py > gpurun_out/prof_batched.log 2>&1
ls -la gpurun_out | tail -8
tail -3 gpurun_out/prof_single.log gpurun_out/prof_batched.log
