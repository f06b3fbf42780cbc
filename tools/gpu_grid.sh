#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 2>&1 | tail -5
AB_MODES=4,7 AB_SHAPES=64x256,14x32,14x64 AB_QUICK=1 AB_NOREF=1 timeout -k 5 900 python tools/ab_bench.py 2>&1 | grep "impl': 'ours'" | cut -c1-200
python - <<'PY'
import numpy as np, subprocess
from mpcgpu_b200 import synth
for N, cap in ((32,173),(64,167),(128,167),(256,118),(512,67)):
    d = synth.make_systems(14, N, seed=9)
    np.concatenate([d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]]).astype(np.float32).tofile("/tmp/in.bin")
    for blk in (128, 64):
        print("dropin", N, blk, subprocess.run([f"tests/_build/dropin_demo_{N}", "/tmp/in.bin", "/tmp/out.bin", str(cap), "1e-4", str(blk), "200"], capture_output=True, text=True).stdout.strip(), flush=True)
PY
