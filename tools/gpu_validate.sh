#!/bin/bash
# full validation + bench + ncu evidence (produces the profiles/r01c_* files)
mkdir -p gpurun_out
echo "== smoke"; timeout -k 5 180 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout -k 5 900 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout -k 5 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"; timeout -k 5 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 300 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 100 --warmup 5 --no-cpu --no-configs --prewarm 0.02 --batched-steps 2 --ring 64 \
   > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu full: single N=128 (default), single N=32 (v4), batched 256 (v5)"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster -s 20 -c 1 \
   -f -o gpurun_out/prof_single env BATCH=1 python tools/one_solve.py > gpurun_out/prof_single.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster -s 20 -c 1 \
   -f -o gpurun_out/prof_single32 env BATCH=1 KNOTS=32 python tools/one_solve.py > gpurun_out/prof_single32.log 2>&1
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:pcg_cluster -s 25 -c 1 \
   -f -o gpurun_out/prof_batched env BATCH=256 python tools/one_solve.py > gpurun_out/prof_batched.log 2>&1
echo "== timeline"; timeout -k 5 300 python tools/timeline.py > gpurun_out/timeline.log 2>&1; tail -3 gpurun_out/timeline.log
echo "== micro"; timeout 120 tools/micro/dsmem_exchange > gpurun_out/micro_dsmem.log 2>&1; timeout 120 tools/micro/l2_exchange > gpurun_out/micro_l2.log 2>&1
ls -la gpurun_out | tail -20
