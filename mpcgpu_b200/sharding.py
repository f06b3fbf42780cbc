"""Multi-GPU batching: independent systems are sharded by contiguous blocks, one process per GPU.

The single-trajectory solve does not shard (2 global reductions + 2 halo exchanges per iteration over
a <= 25 MB problem: a cross-GPU hop per step would dominate) -- "replicas only".  A BATCH of systems
does: system i lives on rank ``i // ceil(batch / world)``; a solve needs no communication, and the one
collective per outer SQP step is an all-gather of the per-system converged flags (<= 1 byte each),
which is what the caller's SQP loop needs to decide which trajectories take another step.
Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of systems owned by `rank` (last ranks may own fewer or none)."""
    if batch < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def owner_of(system: int, batch: int, world: int) -> int:
    per = -(-batch // world)
    return system // per


def gather_converged(local_not_converged, batch: int, world: int, rank: int, group=None):
    """All-gather the per-system `max_iter_exit` bytes of every shard into one [batch] uint8 tensor
    (same device as the input).  One collective, `ceil(batch/world)` bytes per rank."""
    import torch
    import torch.distributed as dist

    per = -(-batch // world)
    lo, hi = shard_range(batch, world, rank)
    if local_not_converged.numel() != hi - lo or local_not_converged.dtype != torch.uint8:
        raise ValueError("expected this rank's uint8 flags")
    send = local_not_converged
    if hi - lo != per:                                   # ragged tail: pad to the common block size
        send = torch.zeros(per, dtype=torch.uint8, device=local_not_converged.device)
        send[: hi - lo] = local_not_converged
    if world == 1:
        return send[:batch].clone()
    out = torch.empty(world * per, dtype=torch.uint8, device=send.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out[:batch]


class PendingFlags:
    """The all-gather of one step's flags, in flight.  `wait()` makes the CURRENT stream (not the host) wait for it and returns the
    [batch] uint8 `max_iter_exit` vector of the whole batch; until then the collective overlaps whatever is enqueued after it --
    normally the next step's solve, which does not depend on it."""

    def __init__(self, work, out, batch):
        self._work, self._out, self._batch = work, out, batch

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._out[: self._batch]


def gather_converged_async(local_not_converged, batch: int, world: int, rank: int, group=None) -> PendingFlags:
    """gather_converged without the stream-level wait: the collective is enqueued behind the work already on the current stream
    (it needs this step's flags) and runs on the communicator's own stream."""
    import torch
    import torch.distributed as dist

    per = -(-batch // world)
    lo, hi = shard_range(batch, world, rank)
    if local_not_converged.numel() != hi - lo or local_not_converged.dtype != torch.uint8:
        raise ValueError("expected this rank's uint8 flags")
    send = local_not_converged
    if hi - lo != per:
        send = torch.zeros(per, dtype=torch.uint8, device=local_not_converged.device)
        send[: hi - lo] = local_not_converged
    if world == 1:
        return PendingFlags(None, send, batch)              # no copy, no collective: the flags are already where they are needed
    out = torch.empty(world * per, dtype=torch.uint8, device=send.device)
    work = dist.all_gather_into_tensor(out, send.contiguous(), group=group, async_op=True)
    return PendingFlags(work, out, batch)


class ShardedBatch:
    """This rank's shard of a batch of systems, resident on its GPU, plus the per-step collective."""

    def __init__(self, n: int, N: int, batch: int, world: int, rank: int, device):
        import torch
        self.n, self.N, self.batch, self.world, self.rank = n, N, batch, world, rank
        self.lo, self.hi = shard_range(batch, world, rank)
        self.local = self.hi - self.lo
        self.device = device
        self.iters = torch.zeros(max(1, self.local), dtype=torch.int32, device=device)
        self.flags = torch.zeros(max(1, self.local), dtype=torch.uint8, device=device)

    def solve_step(self, d_S, d_Pinv, d_gamma, d_lambda, max_iter: int, exit_tol: float):
        """Solve the local shard (one launch) and all-gather the converged flags (one collective).
        Returns the [batch] uint8 `max_iter_exit` vector of the whole batch."""
        from . import solver
        if self.local:
            solver.solve_batched(self.n, self.N, self.local, d_S, d_Pinv, d_gamma, d_lambda, self.iters, self.flags,
                                 max_iter, exit_tol)
        return gather_converged(self.flags[: self.local], self.batch, self.world, self.rank)

    def solve_step_async(self, d_S, d_Pinv, d_gamma, d_lambda, max_iter: int, exit_tol: float, iters=None, flags=None) -> PendingFlags:
        """solve_step for a pipelined outer loop: the launch and the collective are enqueued and the call returns; the collective
        overlaps whatever the caller enqueues next.  `iters` / `flags` (per-step result slots of `local` elements) default to the
        shard's own, which the next step overwrites."""
        from . import solver
        it = self.iters if iters is None else iters
        fl = self.flags if flags is None else flags
        if self.local:
            solver.solve_batched(self.n, self.N, self.local, d_S, d_Pinv, d_gamma, d_lambda, it, fl, max_iter, exit_tol)
        return gather_converged_async(fl[: self.local], self.batch, self.world, self.rank)


class ShardedStep:
    """This rank's shard of a batch of trajectories for the whole SQP linear-system step (KKT blocks in, dz out):
    one StepPlan run on the local shard (assembly -> solve -> dz, no host round trip) + the one all-gather of the
    per-trajectory max_iter_exit flags per outer step."""

    def __init__(self, n: int, m: int, N: int, batch: int, world: int, rank: int, plan_factory=None, direct_fallback: bool = False):
        """plan_factory(n, m, N, local_batch) -> object with run(...) and device_flags(); defaults to solver.StepPlan (the CPU
        tests of the host logic inject a stand-in, there is no CPU solver)."""
        self.n, self.m, self.N, self.batch, self.world, self.rank = n, m, N, batch, world, rank
        self.lo, self.hi = shard_range(batch, world, rank)
        self.local = self.hi - self.lo
        self.direct_fallback = direct_fallback
        if plan_factory is None:
            from . import solver
            plan_factory = solver.StepPlan
        self.plan = plan_factory(n, m, N, self.local) if self.local else None

    def step(self, d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter: int, exit_tol: float):
        """Returns the [batch] uint8 max_iter_exit vector of the whole batch (on this rank's device)."""
        import torch
        if self.plan is not None:
            if self.direct_fallback:
                self.plan.run(d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol, direct_fallback=True)
            else:
                self.plan.run(d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol)
            flags = self.plan.device_flags()
        else:
            flags = torch.zeros(0, dtype=torch.uint8, device=d_lambda.device)
        return gather_converged(flags, self.batch, self.world, self.rank)
