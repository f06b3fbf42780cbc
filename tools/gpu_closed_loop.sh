#!/bin/bash
# closed-loop MPC (the reference's simulateMPC, unchanged) with the reference's GBD-PCG headers, the drop-in headers (bit-exact) and
# the drop-in headers with the tolerance-parity body; behaviour builds (b) and timing builds (t)
mkdir -p gpurun_out
cd oracle/_ref/run
KNOTS_LIST=${KNOTS_LIST:-"32 128"}
for K in $KNOTS_LIST; do
  if [ "$K" = "32" ]; then TOL=5e-6; ROWS=${ROWS32:-140}; else TOL=1e-4; ROWS=${ROWS128:-200}; fi
  for m in b t; do
    for v in ref dropin fast; do
      exe=./closed_loop_${v}_${m}_${K}
      [ -x $exe ] || continue
      t0=$SECONDS
      timeout 600 $exe examples/trajfiles/0_0_traj.csv examples/trajfiles/0_0_eepos.traj $TOL $ROWS ../../../gpurun_out/cl_${v}_${m}_${K}.bin 2>&1 | grep -E "knots|rror|GPUassert" | sed "s/^/[$v $m N=$K] /"
      echo "  [$v $m N=$K] wall $((SECONDS - t0)) s"
    done
  done
done
