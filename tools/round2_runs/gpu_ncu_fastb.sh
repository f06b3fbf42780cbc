#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:fastb -s 6 -c 1 \
   -f -o gpurun_out/prof_fastb_32_1 env BATCH=1 SINGLES=10 KNOTS=32 TUNE_C=1 TUNE_MODE=27 TOL=1e-30 python tools/one_solve.py > gpurun_out/prof_fastb_32.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:fastb -s 6 -c 1 \
   -f -o gpurun_out/prof_fastb_128_4 env BATCH=1 SINGLES=10 KNOTS=128 TUNE_C=4 TUNE_MODE=27 TOL=1e-30 python tools/one_solve.py > gpurun_out/prof_fastb_128.log 2>&1
tail -3 gpurun_out/prof_fastb_32.log gpurun_out/prof_fastb_128.log; ls -la gpurun_out/*.ncu-rep
