#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -8 | tee gpurun_out/r2g_pytest_gpu.log
echo "== A/B batch shapes"; AB_NOREF=1 AB_ALLBATCH=1 AB_SHAPES=14x64 AB_MODES=6,11,20,27,28 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2g_ab.log 2>&1; grep batched gpurun_out/r2g_ab.log | cut -c1-200
bash tools/gpu_bench.sh 2>&1 | tee gpurun_out/r2g_bench.log
