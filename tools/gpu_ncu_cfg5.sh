#!/bin/bash
# one full ncu capture of the config-5 kernel (n=64, N=256: the bandwidth-relevant single solve)
mkdir -p gpurun_out
timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:pcg_grid -s 3 -c 1 \
   -f -o gpurun_out/prof_cfg5 env BATCH=1 STATE=64 KNOTS=256 CAP=200 TOL=1e-6 SINGLES=6 python tools/one_solve.py > gpurun_out/prof_cfg5.log 2>&1
tail -3 gpurun_out/prof_cfg5.log
ncu -i gpurun_out/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/ncu_cfg5_raw.csv 2>/dev/null
rm -f gpurun_out/prof_cfg5.ncu-rep
ls -la gpurun_out | tail -4
