#!/bin/bash
# round 2, first GPU pass: fast-kernel tests first (fail fast), then timeline + A/B, then the rest of the GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2a_gpu.log
echo "== fast tests"; timeout -k 5 900 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -25 | tee gpurun_out/r2a_pytest_fast.log
echo "== timeline"; timeout -k 5 300 python tools/timeline_fast.py > gpurun_out/r2a_timeline_fast.log 2>&1; cat gpurun_out/r2a_timeline_fast.log | head -60
echo "== A/B"; AB_SHAPES=14x128,14x32,14x64,14x256 AB_MODES=2,7,11,20 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/r2a_ab.log 2>&1; tail -40 gpurun_out/r2a_ab.log
cp gpurun_out/ab_bench.json gpurun_out/r2a_ab_bench.json 2>/dev/null
echo "== full gpu suite"; timeout -k 5 1500 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -8 | tee gpurun_out/r2a_pytest_gpu.log
