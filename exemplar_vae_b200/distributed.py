"""Multi-GPU plumbing (one process per GPU, torch.distributed / NCCL over NVLink).

The reference is single-device.  Here the exemplar bank is RANGE-SHARDED over the ranks of one
box: rank r draws and encodes only N/G exemplars per step (which splits the dominant exemplar-
encoder cost G ways), the batch is data-parallel, and the log-sum-exp over exemplars is merged
from per-shard (max, sum, count) partials with ONE small all-gather.  Parameter
gradients live in one flat buffer and are averaged with a single all-reduce.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


class ShardedBank(tuple):
    """(centers_shard, log_variance, index_shard) + the global exemplar count."""

    def __new__(cls, items, c_total: int):
        self = super().__new__(cls, items)
        self.c_total = int(c_total)
        return self


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the rows rank ``rank`` owns when ``n`` rows are range-sharded over ``world``."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class FlatGrads:
    """All parameter gradients as views of ONE contiguous buffer: a single all-reduce averages them,
    and the buffer keeps gradient pointers stable for the fused optimizer / CUDA graphs."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.buf = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.buf[off:off + n].view_as(p)
            off += n

    def zero_(self):
        if self.buf.is_cuda:
            from . import ops
            ops.zero_(self.buf)          # cudaMemsetAsync through the C ABI (no ATen fill kernel on the step)
        else:
            self.buf.zero_()

    def all_reduce_mean(self, group=None):
        world = dist.get_world_size(group)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        self.buf.mul_(1.0 / world)


class GradBuckets:
    """Data-parallel gradient averaging in a few contiguous buckets of the flat buffer, each all-reduced
    (NCCL AVG, asynchronously on the communicator's stream) as soon as the backward has produced every
    gradient in it, so that the exchange of the decoder / head / layer-2 gradients hides behind the rest of
    the backward.  Buckets are contiguous parameter ranges in REVERSE registration order (the backward
    produces the last layers first); the bucket that holds the first-registered parameters completes last and
    is issued by ``finish()``, on the caller's stream after ``backward()`` has joined every side stream."""

    def __init__(self, flat: FlatGrads, group=None, n_buckets: int = 3):
        self.flat, self.group = flat, group
        self._avg = dist.get_backend(group) == "nccl"      # gloo (CPU tests) has no AVG: sum, then scale in finish()
        total = flat.buf.numel()
        target = max(1, total // n_buckets)
        self.ranges, self.bucket_of = [], {}
        off, lo, acc = 0, 0, 0
        for i, p in enumerate(flat.params):
            n = p.numel()
            self.bucket_of[p.grad.data_ptr()] = len(self.ranges)
            off += n
            acc += n
            if acc >= target and len(self.ranges) < n_buckets - 1 and i + 1 < len(flat.params):
                self.ranges.append((lo, off))
                lo, acc = off, 0
        self.ranges.append((lo, off))
        self.totals = [0] * len(self.ranges)
        for b in self.bucket_of.values():
            self.totals[b] += 1
        self.pending = list(self.totals)
        self.works = []
        self.fired = [False] * len(self.ranges)
        # gradients that reach .grad through autograd's own accumulation are only COUNTED here: AccumulateGrad may run
        # on a stream outside a CUDA-graph capture, so a collective is never issued from this hook (the next fused
        # dense backward, or finish(), issues every bucket that has become complete)
        for p in flat.params:
            p.register_post_accumulate_grad_hook(lambda t: self.ready(t.grad, may_fire=False))

    def reset(self):
        self.pending = list(self.totals)
        self.fired = [False] * len(self.ranges)
        self.works = []

    def _fire(self, b):
        lo, hi = self.ranges[b]
        self.fired[b] = True
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self.works.append(dist.all_reduce(self.flat.buf[lo:hi], op=op, group=self.group, async_op=True))

    def ready(self, grad, may_fire=True):
        """``grad`` (a view of the flat buffer) is final for this step."""
        b = self.bucket_of.get(grad.data_ptr())
        if b is None:
            return
        self.pending[b] -= 1
        if may_fire:
            for k in range(len(self.ranges) - 1, 0, -1):         # bucket 0 is always issued by finish()
                if self.pending[k] <= 0 and not self.fired[k]:
                    self._fire(k)

    def finish(self):
        for b in range(len(self.ranges) - 1, -1, -1):
            if not self.fired[b]:
                self._fire(b)
        for w in self.works:
            w.wait()
        if not self._avg:
            self.flat.buf.mul_(1.0 / dist.get_world_size(self.group))
        self.reset()


def shard_bank(model, optimizer=None, group=None, shard=True):
    """Switch ``model`` to a range-sharded exemplar bank + data-parallel gradients over ``group``.
    ``shard=False`` keeps the bank replicated (pure data parallelism: every rank draws all N exemplars /
    runs the kNN selection against its own cache) and only averages the gradients."""
    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if shard:
        model.bank_group, model.bank_world, model.bank_rank = group, world, rank
    # decorrelate the per-rank draws (exemplar indices, eps, binarisation)
    model.rng.seed = (model.rng.seed * 1000003 + 7919 * (rank + 1)) & 0xFFFFFFFFFFFFFFFF
    # identical initial weights on every rank
    for p in model.parameters():
        dist.broadcast(p.data, src=0, group=group)
    model.flat_grads = FlatGrads(model.parameters())
    import os
    if os.environ.get("EXVAE_GRAD_SYNC", "buckets") == "single":      # one blocking all-reduce after the backward
        model.grad_sync = lambda: model.flat_grads.all_reduce_mean(group)
        return model
    buckets = GradBuckets(model.flat_grads, group)
    model.grad_buckets = buckets
    from . import ops
    ops.set_grad_ready_hook(buckets.ready)       # dense layers report their in-kernel accumulated gradients
    model.grad_sync = buckets.finish
    return model
