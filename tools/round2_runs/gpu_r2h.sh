#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout -k 5 1800 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -8 | tee gpurun_out/r2h_pytest_gpu.log
echo "== profiles"; bash tools/gpu_profile_r2.sh 2>&1 | tail -14
