#!/bin/bash
# round-2 end check: smoke, the GPU suite, both bench arms (what the driver runs), then the ncu evidence of the same tree
mkdir -p gpurun_out
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gpu tests"; timeout -k 5 1800 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -5 | tee gpurun_out/r02_pytest_gpu.log
bash tools/gpu_bench.sh 2>&1 | tee gpurun_out/r02_bench.log | cut -c1-700
echo "== profiles"; bash tools/gpu_profile_r2.sh 2>&1 | tail -12
