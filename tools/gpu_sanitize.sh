#!/bin/bash
# compute-sanitizer passes over the new kernels (memcheck everywhere, racecheck on the intra-CTA shared-memory protocols)
mkdir -p gpurun_out
echo "== memcheck: schur/dz/step"; timeout -k 5 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_schur.py -m gpu -q -x --timeout=600 -k "vs_oracle and (14-7-32 or 6-3-12) or batched_step and 32" 2>&1 | tail -6
echo "== memcheck: pcg default variants N=32/64/128 + grid 64x256 (smoke-sized)"; timeout -k 5 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=900 -k "reference_configs and (14-32 or 14-64 or 14-128-167-0.0001) or exit_semantics or batched" 2>&1 | tail -6
echo "== racecheck: schur kernels"; timeout -k 5 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_schur.py -m gpu -q -x --timeout=600 -k "vs_oracle and 6-3-12" 2>&1 | tail -8
echo "== synccheck: v4/v5 + schur"; timeout -k 5 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_schur.py -m gpu -q -x --timeout=600 -k "(reference_configs and 14-32) or (vs_oracle and 6-3-12)" 2>&1 | tail -6
