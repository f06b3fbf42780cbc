#!/bin/bash
# quick A/B: variant parity test + per-variant timings (AB_SHAPES / AB_MODES select)
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=300 -k "every_variant or batched" 2>&1 | tail -5
AB_QUICK=1 AB_NOREF=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/ab.log 2>&1; tail -3 gpurun_out/ab.log
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/ab_bench.json")):
        if r["impl"] == "ours": print(r["n"], r["N"], "C", r["cluster"], "mode", r["mode"], "us/iter %.3f" % r["us_per_iter"], "kernel_us %.1f" % r["kernel_us"])
        elif r["impl"] == "ours_batched": print("batched", r["N"], "C", r["cluster"], "mode", r["mode"], "ms %.3f" % r["ms"])
        else: print(r)
except Exception as e: print("no ab json", e)
PY
