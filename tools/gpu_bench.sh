#!/bin/bash
mkdir -p gpurun_out
echo "== bench ours"; timeout -k 5 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
def show(k): print(k, json.dumps(d.get(k))[:1500])
for k in ['value','data','numerics','pcg_iters_per_sec','mean_iters_per_solve','max_iter_exit_frac','linsys_us','e2e','roofline','bitexact','reference_gbdpcg','tolerance_sweep','batched','other_configs','cpu_baseline','clocks','config','dropin_pcg_template','sqp_neighbours','direct_solver']: show(k)
PY
echo "== bench reference arm"; timeout -k 5 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench_ref.json
