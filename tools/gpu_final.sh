#!/bin/bash
# round-end check on one B200: smoke, the whole GPU suite, the bench line and the reference arm
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout -k 5 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 800 gpurun_out/bench_ref.json
