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
    """All parameter gradients as views of ONE contiguous buffer: a few all-reduces average them, and the buffer
    keeps gradient pointers stable for the fused optimizer / CUDA graphs.  The parameters are split into
    ``n_buckets`` contiguous ranges (by size, in registration order); every range starts at a multiple of ``align``
    floats and is padded to one (the multimem all-reduce needs 16-byte aligned slices per rank), the gaps stay zero.
    ``alloc(n)`` may supply the storage (a region of the symmetric-memory arena)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 1, align: int = 1, alloc=None,
                 groups=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        sizes = [p.numel() for p in self.params]
        total = sum(sizes)
        # bucket boundaries: walk the parameters from the LAST one (the backward produces those gradients first) and
        # cut whenever a bucket has reached total / n_buckets -- but only between parameter GROUPS (`groups[i]` = id
        # of the layer parameter i belongs to), so that no bucket waits for half a layer of the next one
        groups = list(groups) if groups is not None else list(range(len(self.params)))
        target = max(1, total // max(n_buckets, 1))
        cuts, acc = [], 0
        for i in range(len(self.params) - 1, 0, -1):
            acc += sizes[i]
            if acc >= target and groups[i] != groups[i - 1] and len(cuts) < n_buckets - 1:
                cuts.append(i)              # a new bucket starts at parameter i
                acc = 0
        cuts = sorted(cuts)
        up = lambda v: (v + align - 1) // align * align
        offsets, self.ranges, self.bucket_of = [], [], {}
        off, lo = 0, 0
        for i, p in enumerate(self.params):
            if i in cuts:
                off = up(off)
                self.ranges.append((lo, off))
                lo = off
            offsets.append(off)
            self.bucket_of[i] = len(self.ranges)
            off += sizes[i]
        off = up(off)
        self.ranges.append((lo, off))
        ref = self.params[0]
        self.buf = alloc(off) if alloc is not None else torch.zeros(off, dtype=ref.dtype, device=ref.device)
        self.buf.zero_()
        for p, o in zip(self.params, offsets):
            p.grad = self.buf[o:o + p.numel()].view_as(p)

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


class McComm:
    """The symmetric-memory arena of one process group and the exvae multimem exchange kernels (csrc/mc_coll.cu):
    all-reduce / all-gather / reduce-scatter as single kernels over NVLink / NVSwitch multicast instead of NCCL rings.
    Regions are bump-allocated in the same order on every rank, so their offsets are symmetric."""

    GRAD_CH0, GRAD_BLOCKS = 0, 16         # signal-pad channels of the gradient all-reduces (comm stream)
    XCHG_CH0, XCHG_BLOCKS = 32, 4         # ... of the K1 exchanges (prior branch stream)

    def __init__(self, group, device, arena_mb: int = 96):
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.arena = symm.empty(arena_mb * (1 << 18), dtype=torch.float32, device=device)
        self.arena.zero_()
        self.hdl = symm.rendezvous(self.arena, group)
        if not self.hdl.has_multicast_support:
            raise RuntimeError("no NVSwitch multicast support")
        if self.hdl.signal_pad_size // 4 // self.world < self.XCHG_CH0 + self.XCHG_BLOCKS:
            raise RuntimeError("signal pad too small")
        self.mc_base = int(self.hdl.multicast_ptr)
        self.pads = int(self.hdl.signal_pad_ptrs_dev)
        self.off = 0
        self.regions = {}
        self._stream = None
        torch.cuda.synchronize()
        dist.barrier(group)

    def region(self, name: str, n: int):
        """(tensor view, offset in floats) of a named region of n floats (allocated once per (name, n))."""
        key = (name, int(n))
        ent = self.regions.get(key)
        if ent is None:
            lo = self.off
            self.off = (lo + int(n) + 63) // 64 * 64
            if self.off > self.arena.numel():
                raise RuntimeError("symmetric arena exhausted")
            ent = self.regions[key] = (self.arena[lo:lo + int(n)], lo)
        return ent

    def comm_stream(self):
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        return self._stream

    def _s(self):
        return torch.cuda.current_stream().cuda_stream

    def all_reduce_(self, off: int, n: int, scale: float):
        from ._lib import lib
        L = lib()
        L.check(L.exvae_mc_allreduce(self.mc_base + 4 * off, self.pads, n, self.rank, self.world, self.GRAD_CH0,
                                     self.GRAD_BLOCKS, scale, self._s()), "mc_allreduce")

    def all_gather(self, src: torch.Tensor, name: str) -> torch.Tensor:
        """src (contiguous, a multiple of 16 bytes) -> [world, *src.shape] in the arena, identical on every rank."""
        from ._lib import lib
        L = lib()
        n = src.numel() * src.element_size() // 4
        view, off = self.region(name, n * self.world)
        L.check(L.exvae_mc_allgather(src.data_ptr(), self.mc_base + 4 * off, self.pads, n, self.rank, self.world,
                                     self.XCHG_CH0, self.XCHG_BLOCKS, self._s()), "mc_allgather")
        return view.view(src.dtype).view(self.world, *src.shape)

    def reduce_scatter(self, region_off: int, out: torch.Tensor):
        """out = sum over ranks of slice ``rank`` of the [world][out.numel()] region at ``region_off``."""
        from ._lib import lib
        L = lib()
        L.check(L.exvae_mc_reduce_scatter(self.mc_base + 4 * region_off, out.data_ptr(), self.pads, out.numel(),
                                          self.rank, self.world, self.XCHG_CH0, self.XCHG_BLOCKS, self._s()),
                "mc_reduce_scatter")
        return out


_MC = {}


def mc_comm(group):
    """The McComm of ``group`` (None when the multimem path is unavailable or disabled with EXVAE_MC=0)."""
    return _MC.get(id(group))


def _make_mc(group, device):
    import os
    if os.environ.get("EXVAE_MC", "1") == "0" or dist.get_backend(group) != "nccl":
        return None
    if id(group) in _MC:
        return _MC[id(group)]
    ok = torch.ones(1, device=device)
    comm = None
    try:
        comm = McComm(group, device)
    except Exception as ex:                      # no multicast / symmetric memory on this box: NCCL collectives
        ok.zero_()
        if dist.get_rank(group) == 0:
            print(f"[exvae] multimem exchanges unavailable ({ex}); using NCCL", flush=True)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)        # all ranks or none
    if ok.item() == 0:
        comm = None
    _MC[id(group)] = comm
    return comm


class GradBuckets:
    """Data-parallel gradient averaging in a few contiguous buckets of the flat buffer, each all-reduced
    (NCCL AVG, asynchronously on the communicator's stream) as soon as the backward has produced every
    gradient in it, so that the exchange of the decoder / head / layer-2 gradients hides behind the rest of
    the backward.  Buckets are contiguous parameter ranges in REVERSE registration order (the backward
    produces the last layers first); the bucket that holds the first-registered parameters completes last and
    is issued by ``finish()``, on the caller's stream after ``backward()`` has joined every side stream."""

    def __init__(self, flat: FlatGrads, group=None, comm=None, arena_off: int = 0):
        self.flat, self.group = flat, group
        self.comm, self.arena_off = comm, arena_off      # multimem path: flat.buf is the arena region at arena_off
        self._avg = dist.get_backend(group) == "nccl"      # gloo (CPU tests) has no AVG: sum, then scale in finish()
        self.ranges = list(flat.ranges)
        self.bucket_of = {p.grad.data_ptr(): flat.bucket_of[i] for i, p in enumerate(flat.params)}
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
        if self.comm is not None:
            # one multimem kernel on the communication stream: rank r reduces slice r in the switch and multicasts it
            cur, cs = torch.cuda.current_stream(), self.comm.comm_stream()
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                self.comm.all_reduce_(self.arena_off + lo, hi - lo, 1.0 / self.comm.world)
                self.works.append(cs.record_event())
            return
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
            if self.comm is not None:
                torch.cuda.current_stream().wait_event(w)
            else:
                w.wait()
        if not self._avg:
            self.flat.buf.mul_(1.0 / dist.get_world_size(self.group))
        self.reset()


def _layer_groups(model):
    """Layer id of every trainable parameter (first two components of its name: 'q_z_layers.0', 'p_x_mean.linear')."""
    ids, out = {}, []
    for name, p in model.named_parameters():
        if p.requires_grad:
            out.append(ids.setdefault(".".join(name.split(".")[:2]), len(ids)))
    return out


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
    import os
    dev = next(model.parameters()).device
    comm = _make_mc(group, dev) if dev.type == "cuda" else None
    model.mc_comm = comm
    if os.environ.get("EXVAE_GRAD_SYNC", "buckets") == "single":      # one blocking all-reduce after the backward
        model.flat_grads = FlatGrads(model.parameters())
        model.grad_sync = lambda: model.flat_grads.all_reduce_mean(group)
        return model
    arena_off = 0
    if comm is not None:
        total = sum(p.numel() for p in model.parameters() if p.requires_grad)
        cap = total + 3 * 4 * world * 16
        view, arena_off = comm.region(f"grads{id(model)}", cap)
        model.flat_grads = FlatGrads(model.parameters(), n_buckets=3, align=4 * world, alloc=lambda n: view[:n],
                                     groups=_layer_groups(model))
    else:
        model.flat_grads = FlatGrads(model.parameters(), n_buckets=3, groups=_layer_groups(model))
    buckets = GradBuckets(model.flat_grads, group, comm, arena_off)
    model.grad_buckets = buckets
    from . import ops
    ops.set_grad_ready_hook(buckets.ready)       # dense layers report their in-kernel accumulated gradients
    model.grad_sync = buckets.finish
    return model
