"""Data parallelism for the hot path: one process per GPU, torch.distributed for plumbing.

The reference is single-process/single-device (no distributed call site at
all); everything here is new work the north_star asks for:

  * C1  gradient all-reduce(sum)/world over NVLink (NCCL): gradients are accumulated into a
        few large flat fp32 buckets (registered as the parameters' .grad storage, so there is
        no copy in or out) and each bucket is all-reduced asynchronously as soon as the
        backward pass has produced all of its gradients -- size is chosen for launch latency
        and overlap, not link count (NVSwitch gives every peer full bandwidth);
  * C2  the tiny gate-sum all-reduce lives in balanced_mmtm._MMTMFunction;
  * rank-identical controller decisions: the learning-speed statistic is computed from the
        all-reduced gradients and replicated weights, so every rank takes the same branch
        without an extra collective; `seed_everything` keeps `random` in lock-step for
        Bias_Mitigation_Random.

MMTM itself is per-sample -> no collective inside the block.  BatchNorm
statistics stay per-rank like torch DDP's default (state this when comparing 8 x 256 with
1 x 2048); pass sync_bn=True to `setup_model` to convert to SyncBatchNorm.
"""
from __future__ import annotations

import os
import random
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """Initialise the default process group from torchrun's environment (no-op if single)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return world > 1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, rank=int(os.environ["RANK"]), world_size=world)
    return True


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def seed_everything(seed: int = 777):
    """Same seeds the reference sets in dataset.py:26-33, identical on every rank."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def shard_batch(n_global: int, rank_: Optional[int] = None, world: Optional[int] = None):
    """Contiguous slice [lo, hi) of a global batch owned by this rank (remainder to low ranks)."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    base, rem = divmod(n_global, world)
    lo = rank_ * base + min(rank_, rem)
    return lo, lo + base + (1 if rank_ < rem else 0)


class ShardedBatches:
    """Wrap a loader of GLOBAL batches `(idx, data, label)`; yield this rank's shard of each.
    Keeps the reference's tuple layout (dataset.py:116-128).

    Every rank gets the SAME number of samples: a global batch whose size is not a multiple of the world size
    loses its last `len % world` samples (remainder="drop", default) or repeats its first samples to fill up
    (remainder="pad").  Equal shards are what makes the mean of the per-rank mean-loss gradients
    (GradientAllReduce: sum / world) equal the global-batch mean gradient, and no rank ever sees an empty shard
    (an empty shard would skip the gate-sum all-reduce the other ranks are waiting in).  Batches smaller than the
    world size are skipped on every rank under "drop"."""

    def __init__(self, loader, rank_=None, world=None, remainder="drop"):
        if remainder not in ("drop", "pad"):
            raise ValueError("remainder must be 'drop' or 'pad'")
        self.loader, self.rank, self.world, self.remainder = loader, rank_, world, remainder

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        world = world_size() if self.world is None else self.world
        for idx, data, label in self.loader:
            n = len(label)
            if n % world:
                if self.remainder == "drop":
                    n -= n % world
                    if n == 0:
                        continue
                    idx, data, label = idx[:n], data[:n], label[:n]
                else:
                    extra = world - n % world
                    rep = [i % n for i in range(extra)]
                    idx = torch.cat([torch.as_tensor(idx), torch.as_tensor(idx)[rep]])
                    data = torch.cat([torch.as_tensor(data), torch.as_tensor(data)[rep]])
                    label = torch.cat([torch.as_tensor(label), torch.as_tensor(label)[rep]])
                    n += extra
            lo, hi = shard_batch(n, self.rank, world)
            yield idx[lo:hi], data[lo:hi], label[lo:hi]


class GradientAllReduce:
    """Bucketed, overlapped gradient averaging (C1).

    Parameters are walked in REVERSE registration order (roughly the order backward produces
    gradients) and packed into flat buckets of ~`bucket_mb`.  Each parameter's `.grad` is a
    view into its bucket, kept alive across steps (`optimizer.zero_grad(set_to_none=False)`
    semantics are enforced by `zero_grad`).  A post-accumulate-grad hook counts arrivals; the
    bucket's all-reduce is issued on the communication stream when it is complete.
    `finish()` waits for all buckets and scales by 1/world.

    Parameters that received NO gradient in a step (the substituted side's excitation FC inside a curation window,
    reference src/balanced_mmtm.py:135-152) get `.grad = None` for the optimizer step, exactly like the single-GPU
    path under `optimizer.zero_grad()` (set_to_none=True): SGD with momentum / weight decay must skip them rather
    than step them with a zero gradient.  `zero_grad()` re-attaches their bucket views.  (Every rank runs the same
    mode, so the set is rank-identical and the all-reduce of the zero slice is harmless.)
    """

    def __init__(self, model: torch.nn.Module, bucket_mb: float = 32.0, group=None, average: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        params = [p for p in model.parameters() if p.requires_grad]
        self.params = params
        cap = int(bucket_mb * (1 << 20) / 4)
        self.buckets: List[torch.Tensor] = []
        self._bucket_of = {}
        self._pending: List[int] = []
        cur, cur_elems = [], 0
        groups = []
        for p in reversed(params):
            if cur and cur_elems + p.numel() > cap:
                groups.append(cur)
                cur, cur_elems = [], 0
            cur.append(p)
            cur_elems += p.numel()
        if cur:
            groups.append(cur)
        for b, grp in enumerate(groups):
            total = sum(p.numel() for p in grp)
            flat = torch.zeros(total, dtype=grp[0].dtype, device=grp[0].device)
            off = 0
            for p in grp:
                p.grad = flat[off:off + p.numel()].view_as(p)
                self._bucket_of[p] = b
                off += p.numel()
            self.buckets.append(flat)
        self._sizes = [len(g) for g in groups]
        self._views = {p: p.grad for p in params}
        self._fired = set()
        self._arrived = [0] * len(groups)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self._comm_stream = torch.cuda.Stream() if params and params[0].is_cuda else None

    def zero_grad(self):
        """Zero the flat buckets in place (keeps .grad views; one memset per bucket)."""
        for p in self.params:
            if p.grad is not self._views[p]:
                p.grad = self._views[p]
        for flat in self.buckets:
            flat.zero_()
        self._arrived = [0] * len(self.buckets)
        self._fired = set()

    def _on_grad(self, p):
        self._fired.add(p)
        b = self._bucket_of[p]
        self._arrived[b] += 1
        if self._arrived[b] == self._sizes[b]:
            self._launch(b)

    def _launch(self, b):
        if self.world <= 1:
            return
        flat = self.buckets[b]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                w = dist.all_reduce(flat, group=self.group, async_op=True)
        else:
            w = dist.all_reduce(flat, group=self.group, async_op=True)
        self._works.append((w, flat))

    def finish(self):
        """Block the compute stream until every bucket is reduced; average."""
        # parameters that received no gradient this step (e.g. the substituted side's
        # excitation FC during a curation window) never fire their hook: flush those buckets
        for b, n in enumerate(self._arrived):
            if 0 < n < self._sizes[b] or (n == 0 and self.world > 1):
                self._launch(b)
        for w, flat in self._works:
            w.wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        if self.world > 1 and self.average:
            for flat in self.buckets:
                flat.div_(self.world)
        self._works = []
        self._arrived = [0] * len(self.buckets)
        if len(self._fired) != len(self.params):
            for p in self.params:
                if p not in self._fired:
                    p.grad = None  # no gradient this step: the optimizer skips it (single-GPU semantics)

    def remove(self):
        for h in self._hooks:
            h.remove()


class DPOptimizer:
    """Thin wrapper so `optimizer.zero_grad()` in the step engine keeps the flat gradient
    views alive (torch >= 2.0 defaults to set_to_none=True, which would drop them)."""

    def __init__(self, optimizer, reducer: GradientAllReduce):
        self.optimizer, self.reducer = optimizer, reducer

    def zero_grad(self, set_to_none: bool = False):
        self.reducer.zero_grad()

    def step(self, *a, **k):
        return self.optimizer.step(*a, **k)

    def __getattr__(self, name):
        return getattr(self.optimizer, name)


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def setup_model(model: torch.nn.Module, optimizer_factory, bucket_mb: float = 32.0, sync_bn: bool = False):
    """Replicate weights from rank 0, install the bucketed reducer, wrap the optimizer."""
    if sync_bn and world_size() > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    broadcast_parameters(model)
    reducer = GradientAllReduce(model, bucket_mb=bucket_mb)
    optimizer = DPOptimizer(optimizer_factory(model.parameters()), reducer)
    return model, optimizer, reducer
