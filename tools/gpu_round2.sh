#!/bin/bash
# Second GPU script: parity tests incl. drop-in headers, golden minting from the reference, closed-loop
# reference example built against the reference headers vs against include/gbd_dropin.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== golden"; timeout -k 5 600 python tools/make_golden.py 2>&1 | tail -20 | tee gpurun_out/make_golden.log
for K in 32 128; do
  for W in ref dropin; do
    echo "== example $W $K"
    ( cd oracle/_ref/run && mkdir -p tmp/results && timeout -k 5 600 ./track_iiwa_pcg_${W}_$K > ../../../gpurun_out/example_${W}_$K.log 2>&1; echo "rc=$?" )
    grep -A3 -E "Exit tol|Tracking err|Linsys times" gpurun_out/example_${W}_$K.log | grep -E "Exit tol|Average|Median|Max" | head -30
  done
done
