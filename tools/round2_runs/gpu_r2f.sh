#!/bin/bash
mkdir -p gpurun_out
echo "== timeline fastb"; timeout -k 5 300 python tools/timeline_fastb.py 2>&1 | grep -v "^  [0-9]" | tee gpurun_out/r2f_timeline_fastb.log
echo "== fast tests"; timeout -k 5 1200 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2f_pytest_fast.log
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128 AB_MODES=11,20,27,28 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2f_ab.log 2>&1; tail -16 gpurun_out/r2f_ab.log
