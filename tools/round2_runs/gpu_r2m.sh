#!/bin/bash
mkdir -p gpurun_out
echo "== fast variant test"; timeout -k 5 900 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=600 -k "every_fast_variant" 2>&1 | tail -5
echo "== A/B"; AB_NOREF=1 AB_ALLBATCH=1 AB_SHAPES=14x128 AB_MODES=27,31 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py 2>&1 | cut -c1-180 | tee gpurun_out/r2m_ab.log
