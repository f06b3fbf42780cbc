#!/bin/bash
mkdir -p gpurun_out
echo "== fast tests"; timeout -k 5 900 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/r2b_pytest_fast.log
echo "== timeline"; timeout -k 5 300 python tools/timeline_fast.py > gpurun_out/r2b_timeline_fast.log 2>&1; cat gpurun_out/r2b_timeline_fast.log | head -90
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128,14x32,14x64,14x256 AB_MODES=2,7,20 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2b_ab.log 2>&1; grep -v batched gpurun_out/r2b_ab.log | tail -30
