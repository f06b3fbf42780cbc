#!/bin/bash
mkdir -p gpurun_out
echo "== direct tests"; timeout -k 5 900 python -m pytest tests/test_gpu_direct.py -m gpu -q -x --timeout=600 2>&1 | tail -4
echo "== closed loop, direct arm"; KNOTS_LIST="32 128" ARMS="direct" bash tools/gpu_closed_loop.sh 2>&1 | tee gpurun_out/r2k_closed_loop_direct.log
