#!/bin/bash
mkdir -p gpurun_out/dbg
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== probe"; ( cd oracle/_ref/run; for W in ref dropin; do echo "-- $W"; timeout 60 ./sqp_probe_${W}_32 examples/trajfiles/0_0_traj.csv examples/trajfiles/0_0_eepos.traj 1e-4 3 2; done ) 2>&1 | tee gpurun_out/probe.log
echo "== capture+pcg in one process"; R=oracle/_ref; for W in "" dropin_; do timeout 30 $R/ref_capture_${W}32 $R/0_0_traj.csv $R/0_0_eepos.traj gpurun_out/dbg/cap_${W}m9.bin 0 2 1 9 2>&1 | tail -3; done; cmp gpurun_out/dbg/cap_m9.bin gpurun_out/dbg/cap_dropin_m9.bin && echo "ref/drop-in capture+solve IDENTICAL"; rm -f gpurun_out/dbg/cap_*m9.bin
echo "== ab"; AB_QUICK=1 timeout -k 5 600 python tools/ab_bench.py 2>&1 | tail -120 | tee gpurun_out/ab_bench.log
