#!/bin/bash
mkdir -p gpurun_out
echo "== fast variants vs oracle (incl. grid fast)"; timeout -k 5 1500 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=1200 -k "every_fast_variant" 2>&1 | tail -6
echo "== A/B vs reference incl. 64x256"; timeout -k 5 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=900 -k "ab_against_unmodified" 2>&1 | tail -4
echo "== A/B config 5"; AB_NOREF=1 AB_SHAPES=64x256 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py 2>&1 | grep -v batched | tee gpurun_out/r2j_ab_cfg5.log
