#!/bin/bash
# round 2: parity of the assembly / dz / step kernels, then per-kernel durations of one batched SQP step under ncu
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_gpu_schur.py tests/test_gpu_direct.py -m gpu -q -x --timeout=600 2>&1 | tail -5
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_step_launches.csv python tools/step_profile.py > gpurun_out/r2n_step.log 2>&1
tail -2 gpurun_out/r2n_step.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2n_step_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    d.setdefault(r[ki][:60], []).append(v)
for k, v in d.items(): print(f"{k:60s} n={len(v):3d} last={v[-1]:10.1f} us")
PY
AB_SCHUR_QUICK=1 timeout -k 5 300 python tools/ab_schur.py 2>&1 | tail -6
