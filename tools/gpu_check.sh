#!/bin/bash
# First-contact GPU script: smoke, parity tests, A/B timing, bench.  Every stage under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== ab"; timeout -k 5 600 python tools/ab_bench.py 2>&1 | tail -80 | tee gpurun_out/ab_bench.log
echo "== bench"; timeout -k 5 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== e2e modes"; for mode in 0 1 2; do GBD_PCG_ZEROCOPY=$mode timeout -k 5 300 python bench.py --no-cpu --no-batched --steps 500 --prewarm 0.2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zerocopy', $mode, 'e2e us/solve', round(d['e2e']['us_per_solve'],1), 'kernel us', round(d['roofline']['kernel_us'],1))"; done 2>&1 | tee gpurun_out/e2e_modes.log
