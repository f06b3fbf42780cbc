#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout -k 5 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
echo "== dropin timings"; python - <<'PY'
import numpy as np, subprocess, os
from mpcgpu_b200 import synth
for N, cap in ((32,173),(128,167),(256,118),(512,67)):
    d = synth.make_systems(14, N, seed=9)
    np.concatenate([d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]]).astype(np.float32).tofile("/tmp/in.bin")
    for blk in (128, 64):
        print(N, blk, subprocess.run([f"tests/_build/dropin_demo_{N}", "/tmp/in.bin", "/tmp/out.bin", str(cap), "1e-4", str(blk), "200"], capture_output=True, text=True).stdout.strip(), flush=True)
PY
