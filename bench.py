#!/usr/bin/env python
"""bench.py -- GBD-PCG hot-path benchmark (contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): BASELINE.json configs[1] -- the Kuka IIWA tracking problem, state_size n=14, knot_points N=128, fp32, PCG cap
167 (include/common/settings.cuh:128-130), exit tol 1e-4 (examples/track_iiwa_pcg.cu:62-68).  INPUTS are the reference's own:
a ring of RING=256 distinct Schur systems assembled by the reference's generate_kkt_submatrices + form_schur_system on sliding
128-knot windows of examples/trajfiles/0_0_traj.csv (oracle/iiwa.py -> oracle/_ref/ref_capture_128, built from the reference
headers; `data: "iiwa"`).  Where those binaries are absent the ring is synthetic (mpcgpu_b200/synth.py) and `data` says so.
A "step" is ONE linear-system solve (one pass of the hot path): step s solves system s mod RING (S+Pinv of the ring = 147 MiB
> 126 MB L2, so every step reads its tiles cold) into a fresh zero lambda.  value = solves/s summed over ranks (for N>1 every
rank runs its own trajectory stream: the single-trajectory path does not shard, "replicas only"; the batched path does and is
reported in "batched" with one NCCL all-gather of converged flags per step).

  value            device-resident throughput of the library default (tolerance-parity "fast" kernels), CUDA events on the
                   launching stream, max over ranks;  "bitexact" = the same steps with GBD_PCG_NUMERICS_BITEXACT
  reference_gbdpcg the UNMODIFIED reference kernel pcg<float,14,128> (oracle/_ref/libref_gbdpcg.so) on the same ring, launched
                   as include/pcg/sqp.cuh:230 does: cudaEvent kernel time + the reference's own stopwatch window (:224-241)
  tolerance_sweep  the five pcg_exit_tol values of examples/track_iiwa_pcg.cu:62-68, ours and the reference kernel
  e2e              the same solves through the host-buffer C-ABI entry (gbd_pcg_plan_solve_host_f32, the solvePCG(h_S,...)
                   replacement): pinned host inputs, H2D + solve + D2H inside the timed region
  roofline         dominant kernel (name read back from the library); achieved = sum(iters) * B_iter(n,N) / event time
                   (B_iter = 4[2(3N-2)n^2 + 6Nn] bytes per PCG iteration per system).  Tiles live in registers
                   across iterations, so this SpMV-equivalent figure is NOT DRAM traffic; "compulsory_gbs" is the true
                   per-solve traffic rate and "traffic" the ncu DRAM bytes per launch of THIS kernel (profiles/ncu_summary.json)
  cpu_baseline     the reference's own QDLDL (oracle/_ref) on a bounded sample of the same ring, on this box's host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STATE, N_KNOT, MAX_ITER, EXIT_TOL = 14, 128, 167, 1e-4
RING = 256
RING_STRIDE = 2
BATCH_TOTAL = 1024
SWEEP_TOLS = (1e-5, 5e-5, 1e-4, 5e-4, 1e-3)          # examples/track_iiwa_pcg.cu:62-68
CAPS = {32: 173, 64: 167, 128: 167, 256: 118, 512: 67}  # include/common/settings.cuh:123-138


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic(kernel: str, cluster: int):
    """DRAM bytes per launch of the named kernel from the committed ncu capture (keyed by kernel and cluster size), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            for e in json.load(f).get("kernels", []):
                if e.get("kernel") == kernel and e.get("cluster") == cluster:
                    return e.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def load_ring(N: int, count: int, stride: int = RING_STRIDE, seed: int = 1000):
    """The ring of systems both arms solve: reference-minted IIWA systems when oracle/_ref has the capture binaries (needs the GPU),
    synthetic ones otherwise.  Returns (systems dict, data label)."""
    try:
        from oracle import iiwa
        if iiwa.available(N):
            return iiwa.ring(N, count, stride), "iiwa"
    except Exception as e:                                    # a broken capture must not take the bench line down
        print(f"bench: IIWA capture failed ({e!r}); falling back to synthetic systems", file=sys.stderr)
    from mpcgpu_b200 import synth
    d = synth.make_systems(N_STATE, N, batch=count, seed=seed)
    d["source"] = "synthetic LQR-like systems (mpcgpu_b200/synth.py), seed %d" % seed
    return d, "synthetic"


def workload_config(data: str, ring: int, source: str):
    """The `config` object: identical for both arms so the driver can see they ran the same thing."""
    return {"workload": "Kuka IIWA track: n=14, N=128, fp32, tol 1e-4, cap 167 (BASELINE.json configs[1]); one linear-system solve "
                        "per step, lambda0 = 0", "ring": ring, "inputs": source,
            "l2": f"ring of {ring} distinct systems = {ring * 2 * 3 * N_STATE * N_STATE * N_KNOT * 4 >> 20} MiB > L2, so each step reads cold tiles",
            "multi_gpu": "replicas only (one trajectory stream per GPU); batched path in 'batched'", "data": data}


class ClockSampler:
    """Samples SM clock + throttle reasons while the timed region runs (pynvml, else nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.stamps, self.window = [], [None, None]
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        if nv:
            for k in dir(nv):
                if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int):
                    names[getattr(nv, k)] = k.replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                if nv:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    self.stamps.append(time.perf_counter())
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, nm in names.items():
                        if bit and (mask & bit) and nm not in ("None", "GpuIdle", "All"):
                            self.reasons.add(nm)
                else:
                    import subprocess
                    out = subprocess.check_output(
                        ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], text=True, timeout=5).strip().split(",")
                    self.samples.append(int(out[0]))
                    self.stamps.append(time.perf_counter())
                    self.max_mhz = int(out[1])
                    for nm, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), out[2:]):
                        if v.strip().lower() == "active":
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        t0, t1 = self.window
        inside = [c for c, t in zip(self.samples, self.stamps) if t0 is not None and t0 <= t <= t1]
        src = "timed region"
        if len(inside) < 3:      # a sub-millisecond timed region: fall back to the samples taken under the same load
            inside, src = list(self.samples[len(self.samples) // 2:]), "pre-warm + timed region (same load)"
        s = sorted(inside)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "sampled_over": src}


# ----------------------------------------------------------------------------- CPU arms

def cpu_baseline(systems, seconds: float, nthreads: int):
    """Reference CPU path on a bounded sample.  Returns dict(value=solves/s, kind, cores, sample, us_per_solve)."""
    n, N = systems["n"], systems["N"]
    B = systems["S"].shape[0]
    try:
        from oracle import qdldl
        have_ref = qdldl.available()
    except Exception:
        have_ref = False
    if have_ref:
        vals = qdldl.values(systems["S"], n, N)             # CSC packing is outside the reference's timed window
        sec1, _ = qdldl.time_batched(vals, systems["gamma"], n, N, reps=1, nthreads=nthreads)
        reps = max(1, int(seconds / max(sec1, 1e-6)))
        sec, _ = qdldl.time_batched(vals, systems["gamma"], n, N, reps=reps, nthreads=nthreads)
        solves = reps * B
        return dict(value=solves / sec, unit="solves/s", cores=nthreads, kind="reference",
                    us_per_solve=1e6 * sec * nthreads / solves,
                    sample=f"{solves} QDLDL_factor+QDLDL_solve pairs (qdldl v0.1.7 float/int, reference source "
                           f"compiled -O3) over {B} distinct n={n} N={N} systems, {nthreads} thread(s), {sec:.1f} s")
    from oracle import pcg as opcg                           # "port": our C restatement of the reference PCG
    t0 = time.perf_counter()
    solves = 0
    while time.perf_counter() - t0 < seconds:
        i = solves % B
        opcg.pcg(systems["S"][i], systems["Pinv"][i], systems["gamma"][i], systems["lambda0"][i], n, N, MAX_ITER, EXIT_TOL)
        solves += 1
    sec = time.perf_counter() - t0
    return dict(value=solves / sec, unit="solves/s", cores=1, kind="port", us_per_solve=1e6 * sec / solves,
                sample=f"{solves} oracle PCG solves (C restatement of pcg.cuh, 1 thread), {sec:.1f} s")


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (QDLDL) on the same config, inputs and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mpcgpu_b200 import synth
    full, data = load_ring(N_KNOT, args.ring)
    ring_n = full["S"].shape[0]
    systems = {k: (full[k][:64] if isinstance(full[k], np.ndarray) else full[k]) for k in full}
    # one trajectory's solves are sequentially dependent and QDLDL is sequential: 1 thread is all this
    # workload can use.  Each "step" is a bounded sample of solves; K steps are timed.
    per_step_s = min(2.0, max(0.02, 60.0 / max(1, args.steps + args.warmup)))
    # --gpus N: the GPU arm runs N independent trajectory streams (replicas), so this arm runs N streams too, one thread each
    nth = max(1, min(int(args.gpus), os.cpu_count() or 1))
    vals = None
    from oracle import qdldl
    have_ref = qdldl.available()
    if have_ref:
        vals = qdldl.values(systems["S"], N_STATE, N_KNOT)
        sec1, _ = qdldl.time_batched(vals, systems["gamma"], N_STATE, N_KNOT, reps=1, nthreads=nth)
        reps = max(1, int(per_step_s / sec1))

        def step():
            s, _ = qdldl.time_batched(vals, systems["gamma"], N_STATE, N_KNOT, reps=reps, nthreads=nth)
            return s, reps * 64
        kind = "reference"
    else:
        from oracle import pcg as opcg

        def step():
            t0 = time.perf_counter()
            for i in range(8):
                opcg.pcg(systems["S"][i], systems["Pinv"][i], systems["gamma"][i], systems["lambda0"][i], N_STATE,
                         N_KNOT, MAX_ITER, EXIT_TOL)
            return time.perf_counter() - t0, 8
        kind = "port"
        nth = 1
    for _ in range(args.warmup):
        step()
    tot_s, tot_n = 0.0, 0
    for _ in range(args.steps):
        s, k = step()
        tot_s += s
        tot_n += k
    value = tot_n / tot_s
    ncores = os.cpu_count()
    extra = {}
    if have_ref:                                             # batched config 4 on all host cores, for context
        secb, _ = qdldl.time_batched(vals, systems["gamma"], N_STATE, N_KNOT, reps=max(1, 16 * ncores // 64 + 1),
                                     nthreads=ncores)
        nb = max(1, 16 * ncores // 64 + 1) * 64
        extra["batched"] = {"traj_per_sec": nb / secb, "cores": ncores,
                            "sample": f"{nb} independent QDLDL solves over {ncores} threads"}
        # BASELINE.json configs[0]: the reference's own CPU-runnable case (knot_points = 32, include/qdldl/sqp.cuh:22-49)
        s32, d32 = load_ring(32, 64, 8, seed=2025)
        v32 = qdldl.values(s32["S"], N_STATE, 32)
        t1, _ = qdldl.time_batched(v32, s32["gamma"], N_STATE, 32, reps=1, nthreads=1)
        r32 = max(1, int(1.0 / max(t1, 1e-6)))
        t32, _ = qdldl.time_batched(v32, s32["gamma"], N_STATE, 32, reps=r32, nthreads=1)
        extra["config0_n32"] = {"us_per_solve": 1e6 * t32 / (r32 * s32["S"].shape[0]), "solves_per_sec": r32 * s32["S"].shape[0] / t32,
                                "cores": 1, "data": d32,
                                "sample": f"{r32 * s32['S'].shape[0]} QDLDL factor+solve pairs, n=14 N=32, 1 thread"}
    cfg = workload_config(data, ring_n, full["source"])
    cfg["solver"] = "QDLDL factor+solve (include/qdldl/sqp.cuh:22-49)" if have_ref else "oracle PCG port"
    cfg["host_cores_available"] = ncores
    line = {
        "impl": "reference", "metric": "linsys_solves_per_sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": data,
        "linsys_us": 1e6 / value, "config": cfg,
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": nth, "kind": kind,
                         "sample": f"{tot_n} solves in {args.steps} steps over the first 64 systems of the ring, {nth} thread(s): one per "
                                   f"trajectory stream, as many streams as the GPU arm has replicas (one trajectory is sequential; QDLDL is "
                                   f"single-threaded by design); factor + solve only, without the reference's D2H/H2D of include/qdldl/sqp.cuh:268-273"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import mpcgpu_b200 as m
    from mpcgpu_b200 import _capi, build, synth

    build.build_lib()                                         # no-op when the in-tree .so is current
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mpcgpu_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _capi.lib()
    n, N, K, W = N_STATE, N_KNOT, args.steps, args.warmup
    esz = 4
    stream = torch.cuda.current_stream().cuda_stream
    peak, peak_src = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed_us(fn, reps, warm=10):
        for q in range(warm):
            fn(q)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for q in range(reps):
            fn(warm + q)
        a1.record()
        torch.cuda.synchronize()
        return 1e3 * a0.elapsed_time(a1) / reps

    # ---------------- inputs: a ring of distinct systems, resident in HBM, larger than L2 (every rank: the same ring)
    host, data = load_ring(N, args.ring)
    ring = host["S"].shape[0]
    dS, dP, dg = (torch.from_numpy(host[k]).to(dev) for k in ("S", "Pinv", "gamma"))
    steps_total = K + W
    lam = torch.zeros(steps_total, n * N, device=dev)
    iters = torch.zeros(steps_total, dtype=torch.int32, device=dev)
    flags = torch.zeros(steps_total, dtype=torch.uint8, device=dev)

    def launch(step, tol=EXIT_TOL):
        i = step % ring
        rc = L.gbd_pcg_solve_f32(n, N, dS[i].data_ptr(), dP[i].data_ptr(), dg[i].data_ptr(), lam[step % steps_total].data_ptr(),
                                 0, 0, 0, 0, iters[step % steps_total:].data_ptr(), flags[step % steps_total:].data_ptr(), MAX_ITER, tol, stream)
        if rc:
            raise _capi.GbdPcgError(rc, "gbd_pcg_solve_f32")

    numerics_default = "fast" if L.gbd_pcg_get_numerics() == _capi.NUMERICS_FAST else "bitexact"
    resolved = _capi.resolved_variant(n, N)
    sampler = ClockSampler(local)
    sampler.start()
    t_pre = time.perf_counter()                               # untimed pre-warm: bring clocks up under this exact load
    while time.perf_counter() - t_pre < args.prewarm:
        for s in range(W, W + min(K, 64)):
            launch(s)
        torch.cuda.synchronize()
    lam.zero_()
    for s in range(W):
        launch(s)
    barrier()
    sampler.window[0] = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_t0 = L.gbd_pcg_launch_count()
    e0.record()
    for s in range(W, W + K):
        launch(s)
    e1.record()
    barrier()
    sampler.window[1] = time.perf_counter()
    launches_timed = L.gbd_pcg_launch_count() - launches_t0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    it_np = iters[W:].cpu().numpy().astype(np.int64)
    fl_np = flags[W:].cpu().numpy()
    tot_iters = sum_over_ranks(float(it_np.sum()))
    value = world * K / (ms * 1e-3)
    b_iter = synth.bytes_per_iteration(n, N, esz)
    kernel_s = ms * 1e-3 / K                                   # average launch duration (back-to-back on one stream)
    achieved = (it_np.sum() / K) * b_iter / kernel_s / 1e9     # this rank's kernel, GB/s
    compulsory = (esz * (2 * (3 * N - 2) * n * n + 3 * N * n) + esz * N * n) / kernel_s / 1e9

    extras = {}
    # ---------------- the same steps with the bit-exact kernels (identical results to the reference kernel, bit for bit)
    if rank == 0:
        prev = _capi.set_numerics(_capi.NUMERICS_BITEXACT)
        try:
            rb = _capi.resolved_variant(n, N)
            kb = min(K, 400)
            lam.zero_()
            us_b = timed_us(lambda q: launch(q), kb, warm=min(W, 10))
            itb = iters[:min(steps_total, kb)].float().mean().item()
            extras["bitexact"] = {"what": "GBD_PCG_NUMERICS_BITEXACT on the same ring: same floating-point operation order as the reference "
                                          "kernel, bit-identical lambda / iterations / flag (tests/test_gpu_parity.py)",
                                  "kernel": f"{rb['kernel']} (cluster {rb['cluster']}, mode {rb['mode']})", "kernel_us": us_b,
                                  "mean_iters": itb, "us_per_iter": us_b / max(itb, 1.0), "solves_per_sec": 1e6 / us_b}
        finally:
            _capi.set_numerics(prev)

    # ---------------- the reference's own stopwatch window (sqp.cuh:224-241), device-resident inputs
    win = []
    lam_w = torch.zeros(64, n * N, device=dev)
    r_s, p_s = torch.zeros(n * N, device=dev), torch.zeros(n * N, device=dev)
    for s in range(64):
        _, _, us = m.linsys_window(n, N, dS[s % ring], dP[s % ring], dg[s % ring], lam_w[s], r_s, p_s, iters[:1],
                                   flags[:1], MAX_ITER, EXIT_TOL)
        win.append(us)
    win = np.array(win[8:])

    # ---------------- the UNMODIFIED reference GBD-PCG kernel on the same ring, same process (rank 0)
    ref_ws = None
    if rank == 0 and not args.no_refgpu:
        try:
            from oracle import refgpu
            if refgpu.available():
                ref_ws = refgpu.RefWorkspace(n, N)
                rlam = torch.zeros(n * N, device=dev)
                nsys = min(ring, 32)

                def ref_launch(q, tol=EXIT_TOL):
                    i = q % nsys
                    rlam.zero_()
                    refgpu.launch(n, N, dS[i], dP[i], dg[i], rlam, ref_ws, MAX_ITER, tol, 128, stream)

                t_zero = timed_us(lambda q: rlam.zero_(), 200)
                us_ref = timed_us(ref_launch, 96, warm=8) - t_zero
                rit = []
                for i in range(nsys):                           # its iteration counts on these systems
                    ref_launch(i)
                    torch.cuda.synchronize()
                    rit.append(int(ref_ws.iters.item()))
                wref = []
                for q in range(40):
                    i = q % nsys
                    rlam.zero_()
                    wref.append(refgpu.linsys_window(n, N, dS[i], dP[i], dg[i], rlam, ref_ws, MAX_ITER, EXIT_TOL)[2])
                extras["reference_gbdpcg"] = {
                    "what": "unmodified reference pcg<float,14,128> (GBD-PCG/include/pcg.cuh:54-218 compiled for sm_100a into oracle/_ref/"
                            "libref_gbdpcg.so), cooperative launch grid 128 x 128 threads as include/pcg/sqp.cuh:230, same ring (first "
                            f"{nsys} systems), same process",
                    "kernel_us": us_ref, "mean_iters": float(np.mean(rit)), "us_per_iter": us_ref / max(1.0, float(np.mean(rit))),
                    "pcg_iters_per_sec": float(np.mean(rit)) / (us_ref * 1e-6),
                    "linsys_us": {"median": float(np.median(wref[5:])), "mean": float(np.mean(wref[5:])),
                                  "what": "the reference's own stopwatch window include/pcg/sqp.cuh:224-241"},
                    "ours_over_reference_kernel_time": us_ref / (1e6 * kernel_s)}
                # ---- the five exit tolerances of examples/track_iiwa_pcg.cu:62-68: ours (default numerics) and the reference kernel
                sweep = []
                nsw = min(nsys, 16)

                def ours_sweep(q, tol):                             # the same systems the reference kernel gets, lambda0 = 0
                    i = q % nsw
                    lam[i].zero_()
                    rc = L.gbd_pcg_solve_f32(n, N, dS[i].data_ptr(), dP[i].data_ptr(), dg[i].data_ptr(), lam[i].data_ptr(), 0, 0, 0, 0,
                                             iters[i:].data_ptr(), flags[i:].data_ptr(), MAX_ITER, tol, stream)
                    if rc:
                        raise _capi.GbdPcgError(rc, "gbd_pcg_solve_f32")

                for tol in SWEEP_TOLS:
                    us_o = timed_us(lambda q: ours_sweep(q, tol), 48, warm=nsw) - t_zero
                    ito = iters[:nsw].float().mean().item()
                    capo = float((flags[:nsw] != 0).float().mean().item())
                    us_r = timed_us(lambda q: ref_launch(q, tol), 48, warm=4) - t_zero
                    itr = []
                    for i in range(nsw):
                        ref_launch(i, tol)
                        torch.cuda.synchronize()
                        itr.append(int(ref_ws.iters.item()))
                    sweep.append({"pcg_exit_tol": tol, "systems": nsw, "ours_kernel_us": us_o, "ours_mean_iters": ito,
                                  "ours_max_iter_exit_frac": capo, "reference_kernel_us": us_r, "reference_mean_iters": float(np.mean(itr))})
                extras["tolerance_sweep"] = sweep
            else:
                extras["reference_gbdpcg"] = {"unavailable": "oracle/_ref/libref_gbdpcg.so not present (built only where /root/reference exists)"}
        except Exception as e:                                 # never fails the bench line
            extras["reference_gbdpcg"] = {"error": repr(e)[:300]}

    # ---------------- e2e: host buffers through the C ABI (pinned inputs, H2D + solve + D2H timed)
    ering = min(ring, 64)
    hS, hP, hg = (torch.from_numpy(host[k][:ering]).pin_memory() for k in ("S", "Pinv", "gamma"))
    hl = torch.zeros(ering, n * N).pin_memory()
    plan = m.HostPlan(n, N, 1)
    Ke = min(K, 2000)

    hl_np = hl.numpy()
    pS, pP, pg, pl = ([int(t[i].data_ptr()) for i in range(ering)] for t in (hS, hP, hg, hl))

    def e2e_step(s):
        i = s % ering
        hl_np[i].fill(0.0)                                      # fresh initial guess in the caller's host buffer
        return plan.solve_raw(pS[i], pP[i], pg[i], pl[i], MAX_ITER, EXIT_TOL)   # H2D + solve + D2H + sync inside

    for s in range(max(3, min(W, 20))):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(Ke):
        e2e_step(s)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    plan.close()
    h2d = esz * (2 * 3 * n * n * N + 2 * n * N)
    d2h = esz * n * N + 4 + 1

    # ---------------- batched config 4: 1024 trajectories sharded over ranks + all-gather of converged flags
    batched = None
    if not args.no_batched:
        Bl = BATCH_TOTAL // world
        bdata = "synthetic"
        hb = None
        try:
            from oracle import iiwa
            if iiwa.available(N):
                full_b = iiwa.perturbed(N, BATCH_TOTAL)           # every rank mints the same batch and keeps its contiguous shard
                hb = {k: full_b[k][rank * Bl:(rank + 1) * Bl] for k in ("S", "Pinv", "gamma")}
                bdata, bsource = "iiwa", full_b["source"]
                del full_b
        except Exception as e:
            print(f"bench: IIWA batch capture failed ({e!r}); synthetic batch", file=sys.stderr)
        if hb is None:
            hb = synth.make_systems(n, N, batch=Bl, seed=5000 + rank)
            bsource = "synthetic LQR-like systems (mpcgpu_b200/synth.py)"
        bS, bP, bg = (torch.from_numpy(np.ascontiguousarray(hb[k])).to(dev) for k in ("S", "Pinv", "gamma"))
        Kb, Wb = args.batched_steps, 3
        # rotating copies of the shard so that consecutive steps never find their inputs in L2 (a 128-system shard is 77 MB)
        shard_bytes = Bl * 2 * 3 * n * n * N * esz
        copies = 1 + max(1, -(-(256 << 20) // shard_bytes))
        bSc = [bS] + [bS.clone() for _ in range(copies - 1)]
        bPc = [bP] + [bP.clone() for _ in range(copies - 1)]
        blam = torch.zeros(Kb + Wb, Bl, n * N, device=dev)
        bit = torch.zeros(Kb + Wb, Bl, dtype=torch.int32, device=dev)
        bfl = torch.zeros(Kb + Wb, Bl, dtype=torch.uint8, device=dev)
        from mpcgpu_b200.sharding import ShardedBatch
        shard = ShardedBatch(n, N, BATCH_TOTAL, world, rank, dev)
        assert shard.local == Bl
        rbat = _capi.resolved_variant(n, N, batched=True)

        def run_steps(first, count):
            """`count` outer steps, pipelined as an SQP loop would: step s = one launch on the shard + the all-gather of its flags,
            which overlaps step s+1's solve (it is waited for one step later).  Device time of the whole sequence, max over ranks."""
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            b0.record()
            pending = None
            for s in range(first, first + count):
                nxt = shard.solve_step_async(bSc[s % copies], bPc[s % copies], bg, blam[s], MAX_ITER, EXIT_TOL, iters=bit[s], flags=bfl[s])
                if pending is not None:
                    pending.wait()                              # the flags of step s-1: what the loop needs before it can plan step s+1
                pending = nxt
            gflags = pending.wait()
            b1.record()
            torch.cuda.synchronize()
            return max_over_ranks(b0.elapsed_time(b1)), gflags

        run_steps(0, Wb)
        tot_ms, _ = run_steps(Wb, Kb)
        bi = sum_over_ranks(float(bit[Wb:].sum().item()))
        conv = sum_over_ranks(float((bfl[Wb:] == 0).sum().item())) / (world * Bl * Kb)
        bsec = tot_ms * 1e-3 / Kb
        fma_per_iter = 2 * (3 * N - 2) * n * n                   # fused multiply-adds of the two band products per system-iteration
        batched = {"workload": f"BASELINE.json configs[3]: {BATCH_TOTAL} trajectories n=14 N=128 sharded {Bl}/GPU, tol {EXIT_TOL:g}, "
                               f"cap {MAX_ITER}", "data": bdata, "inputs": bsource, "traj_per_sec": world * Bl / bsec,
                   "pcg_iters_per_sec": bi / Kb / bsec, "ms_per_step": 1e3 * bsec, "steps": Kb,
                   "mean_iters": bi / Kb / (world * Bl), "converged_frac": conv,
                   "kernel": f"{rbat['kernel']} (cluster {rbat['cluster']}, mode {rbat['mode']}, numerics {'fast' if rbat['fast'] else 'bitexact'})",
                   "collective": ("nccl all_gather of converged flags per step, overlapped with the next step's solve" if world > 1
                                  else "none (1 rank)"),
                   "l2": f"{copies} rotating copies of the shard's matrices ({copies * shard_bytes >> 20} MiB) > L2: every step reads cold tiles",
                   "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                                "achieved": Bl * esz * (2 * (3 * N - 2) * n * n + 4 * N * n) / bsec / 1e9,
                                "frac": Bl * esz * (2 * (3 * N - 2) * n * n + 4 * N * n) / bsec / 1e9 / peak,
                                "what": "COMPULSORY HBM traffic per rank (every system's tiles and vectors read once, lambda written once) / "
                                        "step time: tiles stay in registers across iterations, so this batch is not HBM-bound",
                                "fp32_fma_tflops": 2 * (bi / Kb / world) * fma_per_iter / bsec / 1e12,
                                "fp32_fma_frac_of_peak": 2 * (bi / Kb / world) * fma_per_iter / bsec / 1e12 / 74.4,
                                "fp32_peak_note": "148 SMs x 128 FMA/clk x 1.965 GHz = 74.4 TFLOP/s nominal"}}
        # ---- the same steps with the bit-exact batch kernel (identical results to the reference kernel, system by system)
        prev_num = _capi.set_numerics(_capi.NUMERICS_BITEXACT)
        try:
            rbx = _capi.resolved_variant(n, N, batched=True)
            blam.zero_()                                        # every solve starts from lambda0 = 0, as in the default-numerics steps
            run_steps(0, 1)
            kx = min(Kb, 3)
            x_ms, _ = run_steps(Wb, kx)
            batched["bitexact"] = {"kernel": f"{rbx['kernel']} (cluster {rbx['cluster']}, mode {rbx['mode']})",
                                   "traj_per_sec": world * Bl / (x_ms * 1e-3 / kx), "ms_per_step": x_ms / kx, "steps": kx,
                                   "mean_iters": sum_over_ranks(float(bit[Wb:Wb + kx].sum().item())) / kx / (world * Bl)}
        finally:
            _capi.set_numerics(prev_num)
        # ---- the same batch as whole SQP linear-system steps (row f3): KKT blocks in, dz out; assembly -> solve -> dz per
        # shard in one enqueue, then the flag all-gather
        try:
            from mpcgpu_b200.sharding import ShardedStep
            mctl = n // 2
            kG, kC, kg, kc = (torch.from_numpy(x).to(dev) for x in synth.make_kkt_batch(n, mctl, N, Bl, seed=7000 + rank))
            Ks = 3
            kGs = [kG.clone().reshape(-1) for _ in range(Ks + 1)]
            kC, kg, kc = kC.reshape(-1), kg.reshape(-1), kc.reshape(-1)
            slam = torch.zeros(Ks + 1, Bl * n * N, device=dev)
            sdz = torch.zeros(Bl * ((n + mctl) * (N - 1) + n), device=dev)
            sstep = ShardedStep(n, mctl, N, BATCH_TOTAL, world, rank)
            sstep.step(kGs[Ks], kC, kg, kc, 1e-3, slam[Ks], sdz, MAX_ITER, EXIT_TOL)
            barrier()
            s_ms = 0.0
            for q in range(Ks):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                s0.record()
                gflags = sstep.step(kGs[q], kC, kg, kc, 1e-3, slam[q], sdz, MAX_ITER, EXIT_TOL)
                s1.record()
                torch.cuda.synchronize()
                s_ms += max_over_ranks(s0.elapsed_time(s1))
            sit, _ = sstep.plan.results()
            batched["sqp_step"] = {"what": "gbd_step_run_f32 per shard (form_schur_system -> pcg -> compute_dz, no host round trip) + flag "
                                           "all-gather; synthetic KKT blocks (mpcgpu_b200/synth.py make_kkt_batch)",
                                   "traj_per_sec": world * Bl / (s_ms * 1e-3 / Ks), "ms_per_step": s_ms / Ks, "steps": Ks,
                                   "mean_iters_rank0": float(sit.mean()), "converged_frac_all": float((gflags == 0).float().mean().item())}
        except Exception as e:                                 # never fails the bench line
            batched["sqp_step"] = {"error": repr(e)[:200]}
    # ---------------- the header drop-in pcg<> under the reference's own launch geometry (cooperative, grid = N, 128 threads)
    if rank == 0:
        demo = os.path.join(ROOT, "tests", "_build", "dropin_demo_128")
        if os.path.exists(demo):
            import subprocess
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
                np.concatenate([host["S"][0], host["Pinv"][0], host["gamma"][0], np.zeros(n * N, np.float32)]).tofile(fin)
                try:
                    out = subprocess.run([demo, fin, fout, str(MAX_ITER), repr(EXIT_TOL), "128", "400"], capture_output=True,
                                         text=True, timeout=120).stdout.split()
                    extras["dropin_pcg_template"] = {
                        "what": "include/gbd_dropin pcg<float,14,128> launched as include/pcg/sqp.cuh:230 does "
                                "(cudaLaunchCooperativeKernel, grid 128, block 128), first system of the ring", "iters": int(out[1]),
                        "kernel_us": float(out[5]), "us_per_iter": float(out[5]) / max(1, int(out[1]))}
                except Exception as e:                         # the extras never fail the bench line
                    extras["dropin_pcg_template"] = {"error": repr(e)[:200]}
    # ---------------- the other BASELINE.json configs (rank 0, short): kernel time per solve and SpMV-equivalent GB/s
    if rank == 0 and not args.no_configs:
        others = []
        for (cn, cN, ccap, ctol, what) in ((14, 32, CAPS[32], 1e-4, "configs[0] size on the GPU (n=14, N=32)"),
                                           (14, 512, CAPS[512], 1e-4, "configs[2] long horizon (n=14, N=512)"),
                                           (64, 256, 200, 1e-6, "configs[4] synthetic block=64, N=256, tol 1e-6, cap 200")):
            try:
                if cn == 14:
                    hc, cdata = load_ring(cN, 16, 8, seed=4242)
                else:
                    hc, cdata = synth.make_systems(cn, cN, batch=2, seed=4242), "synthetic"
                nsys = hc["S"].shape[0]
                cS, cP, cg = (torch.from_numpy(hc[k]).to(dev) for k in ("S", "Pinv", "gamma"))
                reps = 64 if cn == 14 else 16
                cl = torch.zeros(reps + 4, cn * cN, device=dev)
                cit = torch.zeros(reps + 4, dtype=torch.int32, device=dev)
                cfl = torch.zeros(reps + 4, dtype=torch.uint8, device=dev)
                rv = _capi.resolved_variant(cn, cN)

                def csolve(q):
                    i = q % nsys
                    rc = L.gbd_pcg_solve_f32(cn, cN, cS[i].data_ptr(), cP[i].data_ptr(), cg[i].data_ptr(), cl[q].data_ptr(), 0, 0,
                                             0, 0, cit[q:].data_ptr(), cfl[q:].data_ptr(), ccap, ctol, stream)
                    if rc:
                        raise _capi.GbdPcgError(rc, "gbd_pcg_solve_f32")

                for q in range(4):
                    csolve(q)
                torch.cuda.synchronize()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for q in range(4, reps + 4):
                    csolve(q)
                c1.record()
                torch.cuda.synchronize()
                us = 1e3 * c0.elapsed_time(c1) / reps
                mit = float(cit[4:].float().mean().item())
                bi_ = synth.bytes_per_iteration(cn, cN, esz)
                entry = {"config": what, "data": cdata, "kernel": f"{rv['kernel']} (cluster {rv['cluster']}, mode {rv['mode']})",
                         "kernel_us": us, "mean_iters": mit, "us_per_iter": us / max(mit, 1.0),
                         "pcg_iters_per_sec": mit / (us * 1e-6), "converged_frac": float((cfl[4:] == 0).float().mean().item()),
                         "spmv_equiv_gbs": mit * bi_ / (us * 1e-6) / 1e9, "frac_of_hbm_peak": mit * bi_ / (us * 1e-6) / 1e9 / peak,
                         "bytes_per_iter": bi_, "tiles_resident_on_chip": True}
                if ref_ws is not None or (not args.no_refgpu):
                    try:
                        from oracle import refgpu
                        if refgpu.available() and (cn, cN) in refgpu.INSTANTIATED:
                            ws2 = refgpu.RefWorkspace(cn, cN)
                            rl2 = torch.zeros(cn * cN, device=dev)

                            def rsolve(q):
                                rl2.zero_()
                                refgpu.launch(cn, cN, cS[q % nsys], cP[q % nsys], cg[q % nsys], rl2, ws2, ccap, ctol, 128, stream)

                            entry["reference_kernel_us"] = timed_us(rsolve, 24 if cn == 14 else 8, warm=3)
                            entry["reference_last_iters"] = int(ws2.iters.item())
                    except Exception as e:
                        entry["reference_kernel_error"] = repr(e)[:160]
                others.append(entry)
            except Exception as e:                             # the extras never fail the bench line
                others.append({"config": what, "error": repr(e)[:200]})
        extras["other_configs"] = others
    # ---------------- rows f1 / f2: Schur + preconditioner assembly and dz recovery around the solve
    if rank == 0 and not args.no_configs:
        try:
            mctl = n // 2
            kGb, kCb, kgb, kcb = synth.make_kkt_batch(n, mctl, N, 1, seed=99)
            kG0, kC, kg, kc = (torch.from_numpy(np.ascontiguousarray(x).reshape(-1)).to(dev) for x in (kGb, kCb, kgb, kcb))
            kG = kG0.clone()
            kS, kP = torch.zeros(3 * n * n * N, device=dev), torch.zeros(3 * n * n * N, device=dev)
            kgam, klam = torch.zeros(n * N, device=dev), torch.randn(n * N, device=dev)
            kdz = torch.zeros((n + mctl) * (N - 1) + n, device=dev)
            pk = [int(x.data_ptr()) for x in (kG, kC, kg, kc, kS, kP, kgam, klam, kdz)]

            def f_schur(q):
                kG.copy_(kG0)
                rc = L.gbd_form_schur_system_f32(n, mctl, N, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], 1e-3, stream)
                assert rc == 0, rc

            t_cp = timed_us(lambda q: kG.copy_(kG0), 200)
            t_fs = timed_us(f_schur, 200) - t_cp
            t_dzz = timed_us(lambda q: L.gbd_compute_dz_f32(n, mctl, N, pk[0], pk[1], pk[2], pk[7], pk[8], stream), 200)
            extras["sqp_neighbours"] = {
                "what": "rows f1/f2: gbd_form_schur_system_f32 (replaces form_schur_system, include/pcg/linsys_setup.cuh:621-657) and "
                        "gbd_compute_dz_f32 (replaces compute_dz, include/common/dz.cuh:125-136), n=14 m=7 N=128, device-resident, "
                        "back-to-back launches; reference kernels on the same box: profiles/r01c_ab_schur.json",
                "form_schur_us": t_fs, "compute_dz_us": t_dzz, "launches_per_call": {"form_schur": 2, "compute_dz": 1}}
        except Exception as e:                                 # the extras never fail the bench line
            extras["sqp_neighbours"] = {"error": repr(e)[:200]}
    # ---------------- row f4: the direct solver (block cyclic reduction) on the same ring of systems
    if rank == 0 and not args.no_configs:
        try:
            dlam = torch.zeros(n * N, device=dev)

            def dsolve(q):
                i = q % ring
                rc = L.gbd_bcr_solve_f32(n, N, dS[i].data_ptr(), dg[i].data_ptr(), dlam.data_ptr(), stream)
                assert rc == 0, rc

            us_d = timed_us(dsolve, 400, warm=20)
            extras["direct_solver"] = {
                "what": "gbd_bcr_solve_f32: block cyclic reduction in one 16-CTA cluster, no preconditioner, no iteration cap "
                        "(GPU alternative to the reference's CPU QDLDL path; accuracy and A/B in profiles/r01c_ab_direct.json)",
                "kernel_us": us_d, "solves_per_sec": 1e6 / us_d}
        except Exception as e:
            extras["direct_solver"] = {"error": repr(e)[:200]}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline({k: (host[k][:64] if isinstance(host[k], np.ndarray) else host[k]) for k in host},
                           seconds=args.cpu_seconds, nthreads=1)

    if rank == 0:
        line = {
            "metric": "linsys_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": data, "numerics": numerics_default,
            "pcg_iters_per_sec": tot_iters / (ms * 1e-3), "mean_iters_per_solve": tot_iters / (world * K),
            "max_iter_exit_frac": float((fl_np != 0).mean()),
            "linsys_us": {"mean": float(win.mean()), "median": float(np.median(win)), "p95": float(np.percentile(win, 95)),
                          "what": "reference stopwatch window include/pcg/sqp.cuh:224-241 (sync, launch, 2 D2H, sync), "
                                  "device-resident inputs"},
            "config": workload_config(data, ring, host["source"]),
            "e2e": {"value": world * Ke / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "us_per_solve": 1e6 * e2e_s / Ke, "steps": Ke,
                    "api": "gbd_pcg_plan_solve_host_f32 (replaces solvePCG(h_S,...), interface.cuh:24-89)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _ncu_traffic(resolved["kernel"], resolved["cluster"]), "peak_source": peak_src,
                         "kernel": f"{resolved['kernel']}<14,128,{resolved['cluster']}> (mode {resolved['mode']}, {resolved['threads']} threads x "
                                   f"{resolved['cluster']} CTAs, read back from gbd_pcg_resolved_variant)",
                         "kernel_us": 1e6 * kernel_s, "us_per_iter": 1e6 * kernel_s / max(1.0, it_np.mean()), "bytes_per_iter": b_iter,
                         "compulsory_gbs": compulsory,
                         "note": "achieved is SpMV-equivalent bytes (tiles stay on-chip across iterations), not DRAM traffic; "
                                 "this config is latency-bound by construction (0.6 MB working set, one cluster-wide reduction per iteration)"},
            "clocks": clocks,
            "gpu_launches": int(launches_timed),
        }
        line.update(extras)
        if batched:
            line["batched"] = batched
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=RING)
    ap.add_argument("--batched-steps", type=int, default=5)
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short legs for BASELINE.json configs 0, 2 and 4")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the reference GBD-PCG kernel legs (oracle/_ref)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--prewarm", type=float, default=0.5, help="seconds of untimed solves before the W warm-up steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
