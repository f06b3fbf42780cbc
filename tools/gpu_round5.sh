#!/bin/bash
# round-1 session 4: v4 (packet) kernel parity + A/B incl. config 5 (n=64, N=256) + bench
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== ab"; AB_QUICK=1 timeout -k 5 900 python tools/ab_bench.py > gpurun_out/ab.log 2>&1; tail -5 gpurun_out/ab.log
python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/ab_bench.json")):
        if r["impl"] == "ours": print(r["n"], r["N"], "C", r["cluster"], "mode", r["mode"], "us/iter %.3f" % r["us_per_iter"], "kernel_us %.1f" % r["kernel_us"])
        elif r["impl"] == "ours_batched": print("batched", r["N"], "C", r["cluster"], "mode", r["mode"], "ms %.3f" % r["ms"])
        else: print(r)
except Exception as e: print("no ab json", e)
PY
echo "== bench"; timeout -k 5 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
