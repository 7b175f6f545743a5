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
        self.buf.zero_()

    def all_reduce_mean(self, group=None):
        world = dist.get_world_size(group)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        self.buf.mul_(1.0 / world)


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
    model.grad_sync = lambda: model.flat_grads.all_reduce_mean(group)
    return model
