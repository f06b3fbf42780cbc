#!/bin/bash
# third sanitizer pass of round 2: the reworked assembly kernels (register-blocked products, 128-bit pivot snapshots, the
# warp-per-row launch with the paired Gauss-Jordan pass) under memcheck, racecheck and synccheck, both team mappings
mkdir -p gpurun_out
{
for tool in memcheck racecheck synccheck; do
echo "== $tool: tests/test_gpu_schur.py (form_schur vs oracle, both teams; batched step plan)"
timeout -k 5 900 compute-sanitizer --tool $tool --print-limit 6 python -m pytest tests/test_gpu_schur.py -m gpu -q -x --timeout=900 -k "bit_exact_vs_oracle or batched_step_plan" 2>&1 | grep -v "Host Frame\|Device Frame\|^=========\s*$" | tail -8
done
} 2>&1 | tee gpurun_out/r02_compute_sanitizer_c.log
