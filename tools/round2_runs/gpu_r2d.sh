#!/bin/bash
mkdir -p gpurun_out
echo "== fast tests"; timeout -k 5 1200 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/r2d_pytest_fast.log
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128,14x512 AB_MODES=5,11,20,27,28 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2d_ab.log 2>&1; tail -30 gpurun_out/r2d_ab.log
