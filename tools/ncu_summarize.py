#!/usr/bin/env python
"""Turn the round-2 ncu captures (gpurun_out/prof_r02_*.ncu-rep, tools/gpu_profile_r2.sh) into the tracked evidence:
profiles/r02_ncu_<name>_raw.csv (the raw page), profiles/r02_ncu_<name>_source.csv (per-instruction stall samples, trimmed to the
sampled lines) and profiles/r02_ncu_summary.json; also refreshes the `kernels` list of profiles/ncu_summary.json that bench.py
reads for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = {"gpu__time_duration.sum": "gpu_time", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block", "launch__cluster_dim_x": "cluster",
        "lts__t_sector_hit_rate.pct": "lts_hit_rate_pct", "l1tex__t_sector_hit_rate.pct": "l1tex_hit_rate_pct",
        "sm__cycles_active.avg": "sm_cycles_active_avg", "sm__cycles_active.max": "sm_cycles_active_max",
        "smsp__cycles_active.avg": "smsp_cycles_active_avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return x


summary = {"round": 2, "how": "ncu --set full --clock-control none --import-source on (tools/gpu_profile_r2.sh), exported by tools/ncu_summarize.py",
           "kernels": []}
for name, extra in (("single128", {"workload": "one solve n=14 N=128 (synthetic ring of tools/one_solve.py), default numerics"}),
                    ("single32", {"workload": "one solve n=14 N=32, default numerics"}),
                    ("batched256", {"workload": "256 systems n=14 N=128 in one launch, default numerics", "systems": 256}),
                    ("cfg5", {"workload": "one solve n=64 N=256 (BASELINE config 5), default numerics: the tolerance-parity grid kernel"})):
    rep = os.path.join(OUT, f"prof_r02_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = ncu(["-i", rep, "--page", "raw", "--csv"])
    open(os.path.join(PROF, f"r02_ncu_{name}_raw.csv"), "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    ent = {"name": name}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    for h, u, v in zip(hdr, units, vals):
        if h == "Kernel Name":
            ent["kernel"] = v
        if h in WANT:
            x = to_float(v)
            if isinstance(x, float) and u in scale and ("bytes" in h or "time" in h):
                x *= scale[u]
            ent[WANT[h] + ("_us" if h == "gpu__time_duration.sum" else "")] = x
    ent.update(extra)
    if "dram_bytes_read" in ent:
        ent["dram_bytes_per_launch"] = ent["dram_bytes_read"] + ent.get("dram_bytes_write", 0)
    src = ncu(["-i", rep, "--page", "source", "--csv"])
    srows = list(csv.reader(io.StringIO(src)))
    h2 = srows[1]
    ix = {h: i for i, h in enumerate(h2)}
    stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: 0 for s in stalls}
    kept = [srows[0], h2]
    for r in srows[2:]:
        if len(r) != len(h2):
            continue
        if int(r[ix["# Samples"]] or 0) > 0:
            kept.append(r)
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
    with open(os.path.join(PROF, f"r02_ncu_{name}_source.csv"), "w", newline="") as f:
        csv.writer(f).writerows(kept)
    ent["stall_samples"] = {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v}
    summary["kernels"].append(ent)
# launch list of the bench command (gpu__time_duration.sum per launch; cold-cache and serialised: shares, not absolutes)
ll = os.path.join(OUT, "r02_launches_bench.csv")
if os.path.exists(ll):
    import collections
    import shutil
    shutil.copy(ll, os.path.join(PROF, "r02_launches_bench.csv"))
    rows = list(csv.reader(open(ll)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(r[ix["Metric Unit"]], 1.0)
        k = r[ix["Kernel Name"]].split("(")[0]
        tot[k][0] += 1
        tot[k][1] += v
    ours = {k: v for k, v in tot.items() if "gbd::" in k}
    T = sum(v[1] for v in ours.values())
    summary["launch_list"] = {"file": "profiles/r02_launches_bench.csv",
                              "cmd": "python bench.py --steps 100 --warmup 5 --no-cpu --prewarm 0.02 --batched-steps 2 --ring 64 --no-refgpu --no-configs",
                              "note": "the capture window (-s 20 -c 400) ends inside the bit-exact leg; the timed region launches only the first "
                                      "kernel below; the reference's assembly kernels in the list belong to the input-minting subprocess",
                              "kernels": [{"kernel": k, "launches": v[0], "mean_us_under_ncu": v[1] / v[0], "share_of_our_gpu_time": v[1] / T}
                                          for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1])]}
with open(os.path.join(PROF, "r02_ncu_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
# the list bench.py reads (kernel family name + cluster size -> DRAM bytes per launch)
main = json.load(open(os.path.join(PROF, "ncu_summary.json")))
fam = []
for e in summary["kernels"]:
    k = e.get("kernel", "")
    family = "gbd::pcg_cluster_kernel_fastb" if "fastb" in k else ("gbd::pcg_cluster_kernel_fast" if "kernel_fast" in k else k.split("<")[0])
    if e["name"].startswith("single"):
        fam.append({"kernel": family, "cluster": int(e.get("cluster", 0)), "N": 128 if e["name"] == "single128" else 32,
                    "dram_bytes_per_launch": e.get("dram_bytes_per_launch"), "source": f"profiles/r02_ncu_{e['name']}_raw.csv"})
main["kernels"] = fam
with open(os.path.join(PROF, "ncu_summary.json"), "w") as f:
    json.dump(main, f, indent=1)
print(json.dumps(summary, indent=1)[:3000])
