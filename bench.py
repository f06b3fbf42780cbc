#!/usr/bin/env python
"""bench.py -- GBD-PCG hot-path benchmark (contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): BASELINE.json configs[1] -- one IIWA-sized trajectory, state_size n=14,
knot_points N=128, fp32, PCG cap 167 (include/common/settings.cuh:128-130), exit tol 1e-4
(examples/track_iiwa_pcg.cu:62-68), synthetic Schur systems (mpcgpu_b200/synth.py).  A "step" is ONE
linear-system solve (one pass of the hot path): step s solves system s mod RING from a ring of RING=256
distinct device-resident systems (S+Pinv = 154 MB > 126 MB L2, so every step reads its tiles cold)
into a fresh zero lambda.  value = solves/s summed over ranks (for N>1 every rank runs its own
trajectory stream: the single-trajectory path does not shard, "replicas only"; the batched path does
and is reported in the "batched" object with one NCCL all-gather of converged flags per step).

  value        device-resident throughput, CUDA events on the launching stream, max over ranks
  e2e          same solves through the host-buffer C-ABI entry (gbd_pcg_plan_solve_host_f32, the
               solvePCG(h_S,...) replacement): pinned host inputs, H2D + solve + D2H inside the
               timed region, wall clock with device syncs on both sides
  roofline     dominant kernel = pcg_cluster_kernel; achieved = sum(iters) * B_iter(n,N) / event time
               (B_iter = 4[2(3N-2)n^2 + 6Nn] bytes per PCG iteration per system).  Tiles
               live in registers/smem across iterations, so this SpMV-equivalent figure is NOT DRAM
               traffic; "compulsory_gbs" is the true per-solve traffic rate and "traffic" the ncu DRAM
               bytes per launch (profiles/).
  cpu_baseline the reference's own QDLDL (oracle/_ref) -- or the oracle PCG port if that is absent --
               on a bounded sample of the same systems, on this box's host cores.
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
BATCH_TOTAL = 1024


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            return json.load(f).get("single_solve", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


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
    """--impl reference: the reference's own CPU implementation (QDLDL) on the same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mpcgpu_b200 import synth
    systems = synth.make_systems(N_STATE, N_KNOT, batch=64, seed=2024)
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
        s32 = synth.make_systems(N_STATE, 32, batch=64, seed=2025)
        v32 = qdldl.values(s32["S"], N_STATE, 32)
        t1, _ = qdldl.time_batched(v32, s32["gamma"], N_STATE, 32, reps=1, nthreads=1)
        r32 = max(1, int(1.0 / max(t1, 1e-6)))
        t32, _ = qdldl.time_batched(v32, s32["gamma"], N_STATE, 32, reps=r32, nthreads=1)
        extra["config0_n32"] = {"us_per_solve": 1e6 * t32 / (r32 * 64), "solves_per_sec": r32 * 64 / t32, "cores": 1,
                                "sample": f"{r32 * 64} QDLDL factor+solve pairs, n=14 N=32, 1 thread"}
    line = {
        "impl": "reference", "metric": "linsys_solves_per_sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "linsys_us": 1e6 / value,
        "config": {"workload": "IIWA-size single trajectory: n=14, N=128, fp32 (BASELINE.json configs[1])",
                   "solver": "QDLDL factor+solve (include/qdldl/sqp.cuh:22-49)" if have_ref else "oracle PCG port",
                   "host_cores_available": ncores},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": nth, "kind": kind,
                         "sample": f"{tot_n} solves in {args.steps} steps, {nth} thread(s): one per trajectory stream, as many "
                                   f"streams as the GPU arm has replicas (one trajectory is sequential; QDLDL is "
                                   f"single-threaded by design)"},
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
    ring = args.ring

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

    # ---------------- inputs: a ring of distinct systems, resident in HBM, larger than L2
    host = synth.make_systems(n, N, batch=ring, seed=1000 + rank)
    dS, dP, dg = (torch.from_numpy(host[k]).to(dev) for k in ("S", "Pinv", "gamma"))
    steps_total = K + W
    lam = torch.zeros(steps_total, n * N, device=dev)
    iters = torch.zeros(steps_total, dtype=torch.int32, device=dev)
    flags = torch.zeros(steps_total, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    esz = 4

    def launch(step):
        i = step % ring
        rc = L.gbd_pcg_solve_f32(n, N, dS[i].data_ptr(), dP[i].data_ptr(), dg[i].data_ptr(), lam[step].data_ptr(),
                                 0, 0, 0, 0, iters[step:].data_ptr(), flags[step:].data_ptr(), MAX_ITER, EXIT_TOL, stream)
        if rc:
            raise _capi.GbdPcgError(rc, "gbd_pcg_solve_f32")

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
    tot_iters = sum_over_ranks(float(it_np.sum()))
    value = world * K / (ms * 1e-3)
    b_iter = synth.bytes_per_iteration(n, N, esz)
    peak, peak_src = _peaks()
    kernel_s = ms * 1e-3 / K                                   # average launch duration (back-to-back on one stream)
    achieved = (it_np.sum() / K) * b_iter / kernel_s / 1e9     # this rank's kernel, GB/s
    compulsory = (esz * (2 * (3 * N - 2) * n * n + 3 * N * n) + esz * N * n) / kernel_s / 1e9

    # ---------------- the reference's own stopwatch window (sqp.cuh:224-241), device-resident inputs
    win = []
    lam_w = torch.zeros(64, n * N, device=dev)
    r_s, p_s = torch.zeros(n * N, device=dev), torch.zeros(n * N, device=dev)
    for s in range(64):
        _, _, us = m.linsys_window(n, N, dS[s % ring], dP[s % ring], dg[s % ring], lam_w[s], r_s, p_s, iters[:1],
                                   flags[:1], MAX_ITER, EXIT_TOL)
        win.append(us)
    win = np.array(win[8:])

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

    # ---------------- batched config 4: 1024 systems sharded over ranks + all-gather of converged flags
    batched = None
    if not args.no_batched:
        Bl = BATCH_TOTAL // world
        hb = synth.make_systems(n, N, batch=Bl, seed=5000 + rank)
        bS, bP, bg = (torch.from_numpy(hb[k]).to(dev) for k in ("S", "Pinv", "gamma"))
        Kb, Wb = args.batched_steps, 3
        blam = torch.zeros(Kb + Wb, Bl, n * N, device=dev)
        bit = torch.zeros(Kb + Wb, Bl, dtype=torch.int32, device=dev)
        bfl = torch.zeros(Kb + Wb, Bl, dtype=torch.uint8, device=dev)
        from mpcgpu_b200.sharding import ShardedBatch
        shard = ShardedBatch(n, N, BATCH_TOTAL, world, rank, dev)
        assert shard.local == Bl
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if Bl * 2 * 3 * n * n * N * esz < (160 << 20) else None

        def bstep(s):
            shard.iters, shard.flags = bit[s], bfl[s]
            return shard.solve_step(bS, bP, bg, blam[s], MAX_ITER, EXIT_TOL)   # 1 launch + 1 all-gather of flags

        for s in range(Wb):
            bstep(s)
        barrier()
        tot_ms = 0.0
        for s in range(Wb, Wb + Kb):
            if flush is not None:
                flush.fill_(s & 0xFF)                          # shard fits in L2: flush between timed iterations
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            b0.record()
            bstep(s)
            b1.record()
            torch.cuda.synchronize()
            tot_ms += max_over_ranks(b0.elapsed_time(b1))
        bi = sum_over_ranks(float(bit[Wb:].sum().item()))
        conv = sum_over_ranks(float((bfl[Wb:] == 0).sum().item())) / (world * Bl * Kb)
        bsec = tot_ms * 1e-3 / Kb
        batched = {"workload": f"BASELINE.json configs[3]: {BATCH_TOTAL} systems n=14 N=128 sharded {Bl}/GPU, "
                               f"tol {EXIT_TOL:g}, cap {MAX_ITER}", "traj_per_sec": world * Bl / bsec,
                   "pcg_iters_per_sec": bi / Kb / bsec, "ms_per_step": 1e3 * bsec, "steps": Kb,
                   "mean_iters": bi / Kb / (world * Bl), "converged_frac": conv,
                   "collective": "nccl all_gather of converged flags per step" if world > 1 else "none (1 rank)",
                   "l2": "flushed between timed steps" if flush is not None else "inputs larger than L2",
                   "roofline": {"bound": "hbm", "achieved": bi / Kb * b_iter / bsec / 1e9 / world, "peak": peak,
                                "unit": "GB/s", "frac": bi / Kb * b_iter / bsec / 1e9 / world / peak,
                                "compulsory_gbs": Bl * esz * (2 * (3 * N - 2) * n * n + 4 * N * n) / bsec / 1e9}}
        launches_b = Kb
        # ---- the same batch as whole SQP linear-system steps (row f3): KKT blocks in, dz out; assembly -> solve -> dz per
        # shard in one enqueue (4 launches), then the flag all-gather
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
                if flush is not None:
                    flush.fill_(q & 0xFF)
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                s0.record()
                gflags = sstep.step(kGs[q], kC, kg, kc, 1e-3, slam[q], sdz, MAX_ITER, EXIT_TOL)
                s1.record()
                torch.cuda.synchronize()
                s_ms += max_over_ranks(s0.elapsed_time(s1))
            sit, _ = sstep.plan.results()
            batched["sqp_step"] = {"what": "gbd_step_run_f32 per shard (form_schur_system -> pcg -> compute_dz, 4 launches, no host "
                                           "round trip) + flag all-gather; synthetic KKT blocks (mpcgpu_b200/synth.py make_kkt_batch)",
                                   "traj_per_sec": world * Bl / (s_ms * 1e-3 / Ks), "ms_per_step": s_ms / Ks, "steps": Ks,
                                   "mean_iters_rank0": float(sit.mean()), "converged_frac_all": float((gflags == 0).float().mean().item())}
        except Exception as e:                                 # never fails the bench line
            batched["sqp_step"] = {"error": repr(e)[:200]}
    # ---------------- extras (rank 0): the reference-minted IIWA system and the header drop-in pcg<> under the
    # reference's own launch geometry (cooperative, grid = N, 128 threads), both on tests/golden/iiwa_128_0.npz
    extras = {}
    gpath = os.path.join(ROOT, "tests", "golden", "iiwa_128_0.npz")
    if rank == 0 and os.path.exists(gpath):
        g = np.load(gpath)
        gS, gP, gg = (torch.from_numpy(g[k]).to(dev) for k in ("S", "Pinv", "gamma"))
        gl = torch.zeros(n * N, device=dev)

        def gsolve():
            gl.zero_()
            rc = L.gbd_pcg_solve_f32(n, N, gS.data_ptr(), gP.data_ptr(), gg.data_ptr(), gl.data_ptr(), 0, 0, 0, 0,
                                     iters[:1].data_ptr(), flags[:1].data_ptr(), MAX_ITER, EXIT_TOL, stream)
            assert rc == 0

        for _ in range(20):
            gsolve()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(200):
            gsolve()
        g1.record()
        torch.cuda.synchronize()
        git = int(iters[0].item())
        extras["iiwa_golden"] = {"system": "tests/golden/iiwa_128_0.npz (reference KKT+Schur assembly of examples/trajfiles/0_0, "
                                           "first SQP iteration)", "iters": git, "reference_kernel_iters": int(g["run0_iters"]),
                                 "lambda_bit_identical_to_reference_kernel": bool(np.array_equal(gl.cpu().numpy(), g["run0_lam"])),
                                 "kernel_us": 1e3 * g0.elapsed_time(g1) / 200 - 2.0, "note": "includes a ~2 us lambda memset per solve (subtracted)"}
        demo = os.path.join(ROOT, "tests", "_build", "dropin_demo_128")
        if os.path.exists(demo):
            import subprocess
            import tempfile
            with tempfile.TemporaryDirectory() as td:
                fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
                np.concatenate([g["S"], g["Pinv"], g["gamma"], np.zeros(n * N, np.float32)]).tofile(fin)
                try:
                    out = subprocess.run([demo, fin, fout, str(MAX_ITER), repr(EXIT_TOL), "128", "400"], capture_output=True,
                                         text=True, timeout=120).stdout.split()
                    extras["dropin_pcg_template"] = {
                        "what": "include/gbd_dropin pcg<float,14,128> launched as include/pcg/sqp.cuh:230 does "
                                "(cudaLaunchCooperativeKernel, grid 128, block 128)", "iters": int(out[1]),
                        "kernel_us": float(out[5]), "us_per_iter": float(out[5]) / max(1, int(out[1]))}
                except Exception as e:                         # the extras never fail the bench line
                    extras["dropin_pcg_template"] = {"error": repr(e)[:200]}
    # ---------------- the other BASELINE.json configs (rank 0, short): kernel time per solve and SpMV-equivalent GB/s
    if rank == 0 and not args.no_configs:
        others = []
        for (cn, cN, ccap, ctol, what) in ((14, 32, 173, 1e-4, "configs[0] size on the GPU (n=14, N=32)"),
                                           (14, 512, 67, 1e-4, "configs[2] long horizon (n=14, N=512)"),
                                           (64, 256, 200, 1e-6, "configs[4] synthetic block=64, N=256, tol 1e-6, cap 200")):
            try:
                nsys = 8 if cn == 14 else 2
                hc = synth.make_systems(cn, cN, batch=nsys, seed=4242)
                cS, cP, cg = (torch.from_numpy(hc[k]).to(dev) for k in ("S", "Pinv", "gamma"))
                reps = 64 if cn == 14 else 16
                cl = torch.zeros(reps + 4, cn * cN, device=dev)
                cit = torch.zeros(reps + 4, dtype=torch.int32, device=dev)
                cfl = torch.zeros(reps + 4, dtype=torch.uint8, device=dev)

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
                bi = synth.bytes_per_iteration(cn, cN, esz)
                others.append({"config": what, "kernel_us": us, "mean_iters": mit, "us_per_iter": us / max(mit, 1.0),
                               "pcg_iters_per_sec": mit / (us * 1e-6), "converged_frac": float((cfl[4:] == 0).float().mean().item()),
                               "spmv_equiv_gbs": mit * bi / (us * 1e-6) / 1e9, "frac_of_hbm_peak": mit * bi / (us * 1e-6) / 1e9 / peak,
                               "bytes_per_iter": bi, "tiles_resident_on_chip": True})
            except Exception as e:                             # the extras never fail the bench line
                others.append({"config": what, "error": repr(e)[:200]})
        extras["other_configs"] = others
    # ---------------- rows f1 / f2: Schur + preconditioner assembly and dz recovery around the solve
    if rank == 0 and not args.no_configs:
        try:
            mctl = n // 2
            rngk = np.random.default_rng(99)
            Gs, Cs, gs = [], [], []
            for kk in range(N):
                Mq = rngk.standard_normal((n, n))
                Gs.append((Mq @ Mq.T / n + np.eye(n)).T.ravel())
                gs.append(rngk.standard_normal(n))
                if kk < N - 1:
                    Mr = rngk.standard_normal((mctl, mctl))
                    Gs.append((Mr @ Mr.T / mctl + np.eye(mctl)).T.ravel())
                    Cs.append((np.eye(n) + rngk.standard_normal((n, n)) / 16).T.ravel())
                    Cs.append((rngk.standard_normal((n, mctl)) / 16).T.ravel())
                    gs.append(rngk.standard_normal(mctl))
            kG0, kC, kg = (torch.from_numpy(np.concatenate(x).astype(np.float32)).to(dev) for x in (Gs, Cs, gs))
            kc = torch.from_numpy((0.1 * rngk.standard_normal(n * N)).astype(np.float32)).to(dev)
            kG = kG0.clone()
            kS, kP = torch.zeros(3 * n * n * N, device=dev), torch.zeros(3 * n * n * N, device=dev)
            kgam, klam = torch.zeros(n * N, device=dev), torch.randn(n * N, device=dev)
            kdz = torch.zeros((n + mctl) * (N - 1) + n, device=dev)

            def t_us(fn, reps=200):
                for _ in range(10):
                    fn()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    fn()
                a1.record()
                torch.cuda.synchronize()
                return 1e3 * a0.elapsed_time(a1) / reps

            pk = [int(x.data_ptr()) for x in (kG, kC, kg, kc, kS, kP, kgam, klam, kdz)]

            def f_schur():
                kG.copy_(kG0)
                rc = L.gbd_form_schur_system_f32(n, mctl, N, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], 1e-3, stream)
                assert rc == 0, rc

            t_cp = t_us(lambda: kG.copy_(kG0))
            t_fs = t_us(f_schur) - t_cp
            t_dzz = t_us(lambda: L.gbd_compute_dz_f32(n, mctl, N, pk[0], pk[1], pk[2], pk[7], pk[8], stream))
            extras["sqp_neighbours"] = {
                "what": "rows f1/f2: gbd_form_schur_system_f32 (replaces form_schur_system, include/pcg/linsys_setup.cuh:621-657) and "
                        "gbd_compute_dz_f32 (replaces compute_dz, include/common/dz.cuh:125-136), n=14 m=7 N=128, device-resident, "
                        "back-to-back launches; reference kernels on the same box: profiles/r01c_ab_schur.json",
                "form_schur_us": t_fs, "compute_dz_us": t_dzz, "launches_per_call": {"form_schur": 2, "compute_dz": 1}}
        except Exception as e:                                 # the extras never fail the bench line
            extras["sqp_neighbours"] = {"error": repr(e)[:200]}
    # ---------------- row f4: the direct solver (block cyclic reduction) on the same ring of systems and on a 1024 batch
    if rank == 0 and not args.no_configs:
        try:
            dlam = torch.zeros(n * N, device=dev)

            def dsolve(q):
                i = q % ring
                rc = L.gbd_bcr_solve_f32(n, N, dS[i].data_ptr(), dg[i].data_ptr(), dlam.data_ptr(), stream)
                assert rc == 0, rc

            for q in range(20):
                dsolve(q)
            torch.cuda.synchronize()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            for q in range(400):
                dsolve(q)
            d1.record()
            torch.cuda.synchronize()
            extras["direct_solver"] = {
                "what": "gbd_bcr_solve_f32: block cyclic reduction in one 16-CTA cluster, no preconditioner, no iteration cap "
                        "(GPU alternative to the reference's CPU QDLDL path; accuracy and A/B in profiles/r01c_ab_direct.json)",
                "kernel_us": 1e3 * d0.elapsed_time(d1) / 400, "solves_per_sec": 400 / (d0.elapsed_time(d1) * 1e-3)}
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
            "dtype": "f32", "data": "synthetic",
            "pcg_iters_per_sec": tot_iters / (ms * 1e-3), "mean_iters_per_solve": tot_iters / (world * K),
            "linsys_us": {"mean": float(win.mean()), "median": float(np.median(win)), "p95": float(np.percentile(win, 95)),
                          "what": "reference stopwatch window include/pcg/sqp.cuh:224-241 (sync, launch, 2 D2H, sync), "
                                  "device-resident inputs"},
            "config": {"workload": "IIWA-size single trajectory: n=14, N=128, fp32, tol 1e-4, cap 167 "
                                   "(BASELINE.json configs[1]); one solve per step",
                       "ring": ring, "l2": f"ring of {ring} distinct systems = {ring * 2 * 3 * n * n * N * esz >> 20} MiB "
                                           "> L2, so each step reads cold tiles",
                       "multi_gpu": "replicas only (one trajectory stream per GPU); batched path in 'batched'"},
            "e2e": {"value": world * Ke / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "us_per_solve": 1e6 * e2e_s / Ke, "steps": Ke,
                    "api": "gbd_pcg_plan_solve_host_f32 (replaces solvePCG(h_S,...), interface.cuh:24-89)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _ncu_traffic(), "peak_source": peak_src, "kernel": "gbd::pcg_cluster_kernel_v2<float,14,128,16,1>",
                         "kernel_us": 1e6 * kernel_s, "bytes_per_iter": b_iter, "compulsory_gbs": compulsory,
                         "note": "achieved is SpMV-equivalent bytes (tiles stay on-chip across iterations), not DRAM traffic; "
                                 "this config is latency-bound by construction (0.6 MB working set)"},
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
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--prewarm", type=float, default=0.5, help="seconds of untimed solves before the W warm-up steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
