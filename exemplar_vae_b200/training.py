"""Training loop — mirror of utils/training.py:5-51, plus a CUDA-graph captured step."""
from __future__ import annotations

import torch

from . import ops


def set_beta(args, epoch):
    """utils/training.py:5-12"""
    if args.warmup == 0:
        beta = 1.
    else:
        beta = 1. * epoch / args.warmup
        if beta > 1.:
            beta = 1.
    return beta


def train_one_epoch(epoch, args, train_loader, model, optimizer):
    """utils/training.py:15-51 — same arguments and return value (epoch means of loss, -RE, KL).
    The three per-step ``.item()`` host syncs of the reference are replaced by one device-side
    accumulator read once per epoch."""
    model.train()
    beta = set_beta(args, epoch)
    dev = next(model.parameters()).device
    if args.approximate_prior is True:
        with torch.no_grad():
            cache = model.cache_z(train_loader.dataset)
    else:
        cache = None
    acc = torch.zeros(3, dtype=torch.float32, device=dev)
    n = 0
    for batch_idx, (data, indices, target) in enumerate(train_loader):
        data, indices = data.to(dev, non_blocking=True), indices.to(dev, non_blocking=True)
        x = model.rng.bernoulli(data) if args.dynamic_binarization else data
        optimizer.zero_grad()
        loss, RE, KL = model.calculate_loss((x, indices), beta, average=True, cache=cache,
                                            dataset=train_loader.dataset)
        loss.backward()
        if model.grad_sync is not None:
            model.grad_sync()
        optimizer.step()
        with torch.no_grad():
            acc += torch.stack((loss.detach(), -RE.detach(), KL.detach()))
        n += 1
    out = (acc / max(n, 1)).tolist()
    return out[0], out[1], out[2]


class GraphedTrainStep:
    """One training step (dynamic binarisation, loss, backward, AdamNormGrad) captured ONCE in a
    CUDA graph and replayed: the launch-bound inner loop of utils/training.py:27-46 without any
    per-step Python or host sync.  Inputs are copied into static device buffers (``data`` [B,P]
    fp32 probabilities, ``indices`` [B,1] int64); ``out`` holds (loss, RE, KL) of the last step."""

    def __init__(self, model, optimizer, args, dataset, batch_size, beta=1.0, warmup_steps=3, use_graph=True):
        self.model, self.opt, self.args, self.dataset, self.beta = model, optimizer, args, dataset, beta
        dev = next(model.parameters()).device
        P = model.resident(dataset).shape[1]
        self.data = torch.zeros(batch_size, P, dtype=torch.float32, device=dev)
        self.indices = torch.zeros(batch_size, 1, dtype=torch.int64, device=dev)
        self.out = torch.zeros(3, dtype=torch.float32, device=dev)
        self.graph = None
        self.launches_per_step = 0
        if getattr(model, "flat_grads", None) is None:
            from .distributed import FlatGrads
            model.flat_grads = FlatGrads(model.parameters())     # stable grad pointers, one memset per step
        ops.set_fused_grad_accumulation(True)                    # dW/db are added into the flat buffer in-kernel
        model.overlap_prior = getattr(model, "bank_group", None) is None   # prior || decoder as parallel graph branches
        model.train()
        # eager warm-up on a side stream (allocator + optimizer state + grad buffers settle)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(warmup_steps, 1)):
                n0 = ops.launch_count()
                self._body()
                self.launches_per_step = ops.launch_count() - n0
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    def _body(self):
        x = self.model.rng.bernoulli(self.data) if self.args.dynamic_binarization else self.data
        self.model.flat_grads.zero_()
        loss, RE, KL = self.model.calculate_loss((x, self.indices), self.beta, average=True, cache=None,
                                                 dataset=self.dataset)
        loss.backward()
        if self.model.grad_sync is not None:
            self.model.grad_sync()                 # data-parallel: one all-reduce of the flat gradient buffer
        self.opt.step()
        with torch.no_grad():
            base = loss._base            # calculate_loss(average=True) returns three views of one [3] tensor
            if base is not None and base.numel() == 3 and RE._base is base and KL._base is base:
                self.out.copy_(base.detach())
            else:
                self.out.copy_(torch.stack((loss.detach(), RE.detach(), KL.detach())))

    def step(self, data=None, indices=None):
        if data is not None:
            self.data.copy_(data, non_blocking=True)
            self.indices.copy_(indices.reshape(-1, 1), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.out
