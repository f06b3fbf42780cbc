#!/bin/bash
# second sanitizer pass of round 2: synccheck on the batch kernel after reconverging warps ahead of its named barriers, and the
# racecheck report of the packet protocol (every hazard it lists is a packet store against a packet poll: a race by construction,
# made safe by the epoch carried in every packet; no hazard on the vector windows, the parked sums or the published scalars)
mkdir -p gpurun_out
{
echo "== synccheck: batch kernel (one CTA per system, and 2-CTA clusters, systems drawn through the work counter)"
timeout -k 5 900 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=900 -k "batched_equals and (1-27 or 2-28)" 2>&1 | grep -v "Host Frame\|Device Frame\|^=========\s*$" | tail -8
echo "== synccheck: every fast variant"
timeout -k 5 900 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=900 -k "every_fast_variant" 2>&1 | grep -v "Host Frame\|Device Frame\|^=========\s*$" | tail -8
} 2>&1 | tee gpurun_out/r02_compute_sanitizer_b.log
echo "== A/B"; AB_NOREF=1 AB_SHAPES=14x128 AB_MODES=20,27 AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py 2>&1 | grep "mode': 27" | cut -c1-170
