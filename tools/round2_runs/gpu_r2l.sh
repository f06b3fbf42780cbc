#!/bin/bash
mkdir -p gpurun_out
echo "== fast tests"; timeout -k 5 1500 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=1200 2>&1 | tail -4
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128,14x32,14x64,14x256 AB_MODES=20 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py 2>&1 | grep -v batched | cut -c1-175 | tee gpurun_out/r2l_ab.log
