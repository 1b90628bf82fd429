"""Multi-GPU execution of the U-Net path on one 8xB200 node.

Two partitionings, both one process per GPU:

* `DataParallelTrainer` — batch-sharded training.  Every rank holds a full replica, runs
  fwd + loss + bwd on its shard (s2s_unet_backward_only with grad_scale = 1/world), the dense grad
  arena is all-reduced (SUM) over NCCL/NVLink, then the identical fused Adam runs on every replica.
  BatchNorm uses per-replica batch statistics (per-GPU batch 16 = the reference's single-device
  batch; documented deviation from global-batch statistics when the global batch is split).
* `sweep` — the reference's independent loops (lead week x bootstrap fold x model x trial,
  training.py:87,257,322) sharded one task at a time per GPU with no data-path collective; only the
  scalar results are gathered.

The reference has no parallelism of its own (its joblib path is dead code, training.py:290-302).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Iterable, Sequence

import numpy as np


# ------------------------------------------------------------------ pure host logic (CPU-testable)
def shard_batch(n: int, rank: int, world: int) -> slice:
    """Contiguous shard of a global batch of n samples for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_weight(n_local: int, n_global: int) -> float:
    """Loss-gradient scale so that SUM-all-reduced local mean-loss gradients equal the global mean-loss
    gradient: each rank's loss is a mean over its n_local samples."""
    return n_local / float(n_global)


def assign_tasks(costs: Sequence[float], n_workers: int) -> list[list[int]]:
    """Static longest-processing-time-first assignment of task indices to workers (trial cost varies
    ~13x between filters=2,n_blocks=3 and filters=3,n_blocks=5)."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * n_workers
    out: list[list[int]] = [[] for _ in range(n_workers)]
    for i in order:
        w = min(range(n_workers), key=lambda k: loads[k])
        out[w].append(i)
        loads[w] += costs[i]
    return out


def trial_cost(filters: int, n_blocks: int, ct_kernel: int, H: int = 64, W: int = 64) -> float:
    """Relative cost (train MFLOP/sample) of one fit, used to balance the sweep."""
    fl, c_prev = 0.0, 1
    for b in range(n_blocks):
        c, px = filters * 4 * 2 ** b, (H >> b) * (W >> b)
        fl += 18.0 * px * (c_prev * c + c * c)
        c_prev = c
    c, px = filters * 4 * 2 ** n_blocks, (H >> n_blocks) * (W >> n_blocks)
    fl += 18.0 * px * (c_prev * c + c * c)
    for b in range(n_blocks):
        c, px = filters * 4 * 2 ** b, (H >> b) * (W >> b)
        fl += 2.0 * ct_kernel ** 2 * 2 * c * c * px / 4 + 18.0 * px * (2 * c * c + c * c)
    return 3.0 * fl / 1e6


def allreduce_mean_grads(grads: "np.ndarray | object", n_local: int, n_global: int, dist) -> None:
    """In-place SUM all-reduce of already shard-weighted gradients (torch tensor, any backend)."""
    del n_local, n_global
    dist.all_reduce(grads)


# ------------------------------------------------------------------ data-parallel trainer (NCCL)
class DataParallelTrainer:
    """Wraps a compiled `Model` replica; `train_on_batch` takes this rank's shard of the global batch."""

    def __init__(self, model, process_group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (backend 'nccl') before DataParallelTrainer")
        self.model, self.dist, self.pg = model, dist, process_group
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        dev = torch.cuda.current_device()

        class _Ptr:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        self.grads = torch.as_tensor(_Ptr(model._grads_ptr, model.n_params_padded), device=f"cuda:{dev}")
        self.params = torch.as_tensor(_Ptr(model._params_ptr, model.n_params_padded), device=f"cuda:{dev}")
        self.stream = torch.cuda.ExternalStream(model.stream.ptr, device=dev)
        self._torch = torch

    def broadcast_weights(self, src: int = 0) -> None:
        with self._torch.cuda.stream(self.stream):
            self.dist.broadcast(self.params, src=src, group=self.pg)
        self.model.stream.synchronize()

    def train_on_batch(self, x_shard, y_shard, n_global: int | None = None):
        n_local = len(x_shard)
        n_global = n_global or n_local * self.world
        stats = self.model.backward_on_batch(x_shard, y_shard, grad_scale=shard_weight(n_local, n_global))
        with self._torch.cuda.stream(self.stream):
            self.dist.all_reduce(self.grads, group=self.pg)
        self.model.apply_adam()
        return stats


def exchange_ipc_handles(mine: bytes, dist, process_group=None) -> bytes:
    """All ranks' 64-byte cudaIpcMemHandle_t blobs concatenated in rank order (what s2s_dp_connect expects).
    Pure host logic over any torch.distributed backend (CPU-testable with gloo)."""
    if len(mine) != 64:
        raise ValueError(f"an IPC handle is 64 bytes, got {len(mine)}")
    world = dist.get_world_size(process_group)
    gathered: list = [None] * world
    dist.all_gather_object(gathered, bytes(mine), group=process_group)
    if any(g is None or len(g) != 64 for g in gathered):
        raise RuntimeError("a rank did not contribute a 64-byte IPC handle")
    return b"".join(gathered)


class PeerDataParallelTrainer:
    """Batch-sharded training with the exchange done by this library's own kernels over NVLink peer memory
    (csrc/dp.cuh): the gradient all-reduce is fused with the Adam update in one kernel, and with `sync_bn=True`
    every BatchNormalization layer normalises with the statistics of the GLOBAL batch, so G GPUs x (B/G) samples
    reproduce the reference's single-device batch of B (model.fit, training.py:102) up to fp32 summation order.
    torch.distributed (any backend) is used once, as the rendezvous that exchanges the 64-byte IPC handles."""

    def __init__(self, model, sync_bn: bool = True, process_group=None):
        import torch
        import torch.distributed as dist
        from ._lib import call
        self.model, self.dist, self.pg = model, dist, process_group
        if dist.is_initialized():
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        else:                                           # single replica: same kernels, no peer to wait for
            self.world, self.rank = 1, 0
        self._call = call
        self._dp = None
        # Set-up is collective: a rank that cannot export / map peer memory must not leave the others waiting, so every
        # rank runs every collective and the outcome is agreed on before anyone proceeds.
        err = None
        mine = (C.c_ubyte * 64)()
        try:
            if os.environ.get("S2S_FORCE_PEER_FAIL") == "1":
                raise RuntimeError("peer-memory exchange disabled by S2S_FORCE_PEER_FAIL")
            h = C.c_void_p()
            call("s2s_dp_create", self.rank, self.world, C.c_size_t(model.n_params_padded), C.byref(h))
            self._dp = h
            call("s2s_dp_ipc_handle", self._dp, mine)
        except Exception as e:          # noqa: BLE001 - reported uniformly below
            err = e
        if self.world > 1:
            blob = exchange_ipc_handles(bytes(mine), dist, process_group)
            if err is None:
                try:
                    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
                    call("s2s_dp_connect", self._dp, buf)
                except Exception as e:  # noqa: BLE001
                    err = e
            outcome: list = [None] * self.world
            dist.all_gather_object(outcome, None if err is None else repr(err), group=process_group)
            bad = {r: o for r, o in enumerate(outcome) if o is not None}
            if bad:
                if self._dp is not None:
                    call("s2s_dp_destroy", self._dp)
                    self._dp = None
                raise RuntimeError(f"peer-memory data parallelism is unavailable on rank(s) {sorted(bad)}: {next(iter(bad.values()))}")
            dist.barrier(group=process_group)          # every rank has mapped every buffer before the first flag is written
        elif err is not None:
            raise err
        call("s2s_unet_attach_dp", model._h, self._dp, 1 if sync_bn else 0)
        self._torch = torch

    def broadcast_weights(self, src: int = 0) -> None:
        """Replicas must start identical (the fused all-reduce keeps them bit-identical afterwards)."""
        if self.world == 1:
            return
        w = self.model.get_weights(as_dict=False) if self.rank == src else None
        box = [w]
        self.dist.broadcast_object_list(box, src=src, group=self.pg)
        if self.rank != src:
            self.model.set_weights(box[0])

    def train_on_batch(self, x_shard, y_shard, n_global: int | None = None):
        n_global = n_global or len(x_shard) * self.world
        out = self.model.dp_train_on_batch(x_shard, y_shard, n_global)
        self.check()        # a peer that timed out left this step unapplied: raise instead of training on
        return out

    def train_on_batches(self, x_shards, y_shards, n_global: int | None = None):
        """A stream of data-parallel steps from lists of pinned host shards (Model.train_on_batches): the copy of shard i + 1
        is staged while step i computes; returns the global (loss, accuracy) per step as an array [steps, 2]."""
        n_global = n_global or len(x_shards[0]) * self.world
        out = self.model.train_on_batches(x_shards, y_shards, n_global=n_global)
        self.check()
        return out

    def check(self) -> None:
        err = C.c_int(0)
        self._call("s2s_dp_error", self._dp, C.byref(err))
        if err.value:
            raise RuntimeError(f"data-parallel exchange timed out waiting for a peer (sync group {err.value - 1})")

    def close(self) -> None:
        if self._dp is not None:
            self.model.stream.synchronize()
            self._call("s2s_unet_attach_dp", self.model._h, None, 0)
            if self.world > 1:
                self.dist.barrier(group=self.pg)       # nobody unmaps while a peer may still read
            self._call("s2s_dp_destroy", self._dp)
            self._dp = None


# ------------------------------------------------------------------ sweep sharding (no collective)
def _sweep_worker(gpu: int, task_ids: list[int], tasks: list, fn: Callable, out_q) -> None:
    os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu)
    for t in task_ids:
        try:
            out_q.put((t, fn(tasks[t]), None))
        except Exception as e:   # a failed trial must not take the sweep down
            out_q.put((t, None, repr(e)))
    out_q.put((None, gpu, None))


class SweepError(RuntimeError):
    """Raised by `sweep` when tasks failed or a worker died; `.results` holds what completed (None elsewhere),
    `.errors` the (task, message) pairs and `.dead` the {gpu: exit code} of workers that did not finish."""

    def __init__(self, msg: str, results: list, errors: list, dead: dict):
        super().__init__(msg)
        self.results, self.errors, self.dead = results, errors, dead


def sweep(tasks: list, fn: Callable, n_gpus: int, costs: Iterable[float] | None = None, poll_s: float = 1.0) -> list:
    """Run fn(task) for every task, one process per GPU, static cost-balanced assignment.
    fn must be a picklable top-level function that builds its own Model (it sees one visible GPU).
    Returns results in task order; raises SweepError (carrying the partial results) if any task failed or a
    worker process died (segfault in the library, sticky CUDA fault, OOM killer, failing import): a dead worker
    never posts its sentinel, so the queue is polled with a timeout and the workers' liveness is checked."""
    import multiprocessing as mp
    import queue as _queue
    costs = list(costs) if costs is not None else [1.0] * len(tasks)
    plan = assign_tasks(costs, n_gpus)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(g, plan[g], tasks, fn, q)) for g in range(n_gpus)]
    for p in procs:
        p.start()
    results, errors = [None] * len(tasks), []
    finished: set[int] = set()          # workers that posted their sentinel
    got: set[int] = set()               # task ids answered (result or error)
    dead: dict[int, int] = {}

    def drain(block: bool) -> bool:
        try:
            t, r, err = q.get(timeout=poll_s) if block else q.get_nowait()
        except _queue.Empty:
            return False
        if t is None:
            finished.add(r)
        else:
            got.add(t)
            if err is not None:
                errors.append((t, err))
            else:
                results[t] = r
        return True

    while len(finished) + len(dead) < n_gpus:
        if drain(block=True):
            continue
        for g, p in enumerate(procs):
            if g in finished or g in dead or p.is_alive():
                continue
            while drain(block=False):   # results it posted before dying are still in the pipe
                pass
            if g not in finished:
                dead[g] = p.exitcode if p.exitcode is not None else -1
                for t in plan[g]:
                    if t not in got:
                        errors.append((t, f"worker for GPU {g} died (exit code {dead[g]}) before running this task"))
    for p in procs:
        p.join(timeout=10)
    if errors or dead:
        msg = f"{len(errors)} sweep task(s) failed"
        if dead:
            msg += f"; worker(s) died on GPU(s) {sorted(dead)} with exit code(s) {[dead[g] for g in sorted(dead)]}"
        raise SweepError(msg + f": {errors[:3]}", results, errors, dead)
    return results
