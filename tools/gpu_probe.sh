#!/bin/bash
# closed-loop check: the reference's own sqpSolvePcg driven on the reference trajectory, built against the reference's
# GBD-PCG headers and against include/gbd_dropin -- every printed quantity (pcg iteration counts per SQP iteration,
# rho, |lambda|^2, |xu|^2) must agree; linsys_us shows the reference's own stopwatch for both
mkdir -p gpurun_out
for K in 32 128; do
  ( cd oracle/_ref/run; for W in ref dropin; do echo "-- $W N=$K"; timeout 120 ./sqp_probe_${W}_$K examples/trajfiles/0_0_traj.csv examples/trajfiles/0_0_eepos.traj 1e-4 3 2; done ) > gpurun_out/probe_$K.log 2>&1
  python - <<PY
import re
t=open("gpurun_out/probe_$K.log").read()
a,b=t.split("-- dropin N=$K")
strip=lambda s: re.sub(r"linsys_us mean [0-9.]+","",s.split("\n",1)[1])
print("N=$K closed loop identical:", strip(a)==strip(b))
print(" ref   linsys_us:", re.findall(r"linsys_us mean ([0-9.]+)",a))
print(" dropin linsys_us:", re.findall(r"linsys_us mean ([0-9.]+)",b))
PY
done
