#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (the tolerance-parity family): memcheck on every fast variant and on batches
# drawn through the work counter, racecheck on the intra-CTA shared-memory protocols (windows, parked sums, scalars), synccheck on
# the named-barrier choreography.  racecheck does not see DSMEM stores from peers, only this CTA's own accesses.
mkdir -p gpurun_out
{
echo "== memcheck: every fast variant vs its oracle + batches through the work counter"
timeout -k 5 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=1200 -k "every_fast_variant or batched_equals" 2>&1 | tail -6
echo "== racecheck: single-solve fast kernel (N=32) and batch kernel (N=32 one CTA per system, N=64 two CTAs)"
timeout -k 5 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=1200 -k "(exit_semantics and 14-32) or (batched_equals and 1-27)" 2>&1 | tail -8
echo "== synccheck: fast + batch kernels"
timeout -k 5 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_fast.py -m gpu -q -x --timeout=900 -k "(exit_semantics and 14-32) or (batched_equals and 1-27)" 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r02_compute_sanitizer.log
