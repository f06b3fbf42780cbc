#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -k 5 300 python -m pytest tests -m gpu -q -x --timeout=200 2>&1 | tail -3
for B in 1024 128; do
GBD_PCG_STATIC_BATCH=1 SPREAD=1 BATCH=$B timeout -k 5 90 python tools/ab_batched_draw.py 2>&1 | tail -1 | tee -a gpurun_out/ab_batched_draw_spread.log
SPREAD=1 BATCH=$B timeout -k 5 90 python tools/ab_batched_draw.py 2>&1 | tail -1 | tee -a gpurun_out/ab_batched_draw_spread.log
done
