#!/bin/bash
# closed-loop MPC (the reference's simulateMPC, unchanged) with the reference's GBD-PCG headers (ref), the drop-in headers (dropin,
# bit-exact bodies), the drop-in headers with the tolerance-parity body (fast), and -- as the noise floor of the experiment -- the
# reference headers again with pcg_exit_tol scaled by 1.001 (refp: a perturbation far below anything the solver promises);
# behaviour builds (b: SQP iterations per control step) and timing builds (t: the reference's linsys stopwatch)
mkdir -p gpurun_out
cd oracle/_ref/run
KNOTS_LIST=${KNOTS_LIST:-"32 128"}
ARMS=${ARMS:-"ref dropin fast direct refp"}
MODES=${MODES:-"b t"}
for K in $KNOTS_LIST; do
  if [ "$K" = "32" ]; then TOL=5e-6; ROWS=${ROWS32:-140}; else TOL=1e-4; ROWS=${ROWS128:-200}; fi
  for m in $MODES; do
    for v in $ARMS; do
      bin=$v; tol=$TOL
      if [ "$v" = "refp" ]; then bin=ref; tol=$(python3 -c "print(repr($TOL*1.001))"); fi
      exe=./closed_loop_${bin}_${m}_${K}
      [ -x $exe ] || continue
      t0=$SECONDS
      timeout 600 $exe examples/trajfiles/0_0_traj.csv examples/trajfiles/0_0_eepos.traj $tol $ROWS ../../../gpurun_out/cl_${v}_${m}_${K}.bin 2>&1 | grep -E "knots|rror|GPUassert|Too many" | sed "s/^/[$v $m N=$K] /"
      echo "  [$v $m N=$K] wall $((SECONDS - t0)) s"
    done
  done
done
