#!/bin/bash
mkdir -p gpurun_out
echo "== fast tests"; timeout -k 5 1200 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/r2c_pytest_fast.log
echo "== sweep debug"; timeout -k 5 300 python tools/debug_sweep.py 2>&1 | tail -12 | tee gpurun_out/r2c_debug_sweep.log
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128,14x512,14x256 AB_MODES=5,11,20,26 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2c_ab.log 2>&1; tail -30 gpurun_out/r2c_ab.log
echo "== closed loop"; KNOTS_LIST="128" ARMS="fast refp" bash tools/gpu_closed_loop.sh 2>&1 | tee gpurun_out/r2c_closed_loop.log
KNOTS_LIST="32" ARMS="refp" bash tools/gpu_closed_loop.sh 2>&1 | tee -a gpurun_out/r2c_closed_loop.log
