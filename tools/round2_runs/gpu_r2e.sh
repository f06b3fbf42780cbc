#!/bin/bash
mkdir -p gpurun_out
echo "== timeline fastb"; timeout -k 5 300 python tools/timeline_fastb.py 2>&1 | tee gpurun_out/r2e_timeline_fastb.log
echo "== fast tests"; timeout -k 5 1200 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/r2e_pytest_fast.log
