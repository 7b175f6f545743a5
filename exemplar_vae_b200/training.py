"""Training loop — mirror of utils/training.py:5-51, plus a CUDA-graph captured step."""
from __future__ import annotations

import os

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
    fp32 probabilities, ``indices`` [B,1] int64); ``out`` holds (loss, RE, KL) of the last step.

    Constructing the object does NOT train: the eager warm-up steps that settle the allocator, the
    optimizer state and the gradient buffers run on the (all-zero) static buffers, and parameters,
    Adam moments, the step counters and the RNG counter are restored afterwards.  ``beta`` lives in a
    device scalar (``set_beta``), so one captured graph follows the warm-up schedule of
    utils/training.py:5-12.  ``rng_override`` = {"eps": [static tensors], "exemplar_indices": static
    tensor} injects the random draws (tests): overwrite those tensors in place between steps."""

    def __init__(self, model, optimizer, args, dataset, batch_size, beta=1.0, warmup_steps=3, use_graph=True,
                 rng_override=None, cache=None):
        self.model, self.opt, self.args, self.dataset = model, optimizer, args, dataset
        dev = next(model.parameters()).device
        P = model.resident(dataset).shape[1]
        self.data = torch.zeros(batch_size, P, dtype=torch.float32, device=dev)
        self.indices = torch.zeros(batch_size, 1, dtype=torch.int64, device=dev)
        self.out = torch.zeros(3, dtype=torch.float32, device=dev)
        self.beta_dev = torch.full((1,), float(beta), dtype=torch.float32, device=dev)
        self._g3 = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        # The loss of this step is mean_b(-RE_b + beta*KL_b) with KL_b = log q_b - log p(z_b), and the backward starts
        # from d loss = 1: d loss / d log p(z_b) = -beta/B for every row (all G*B rows when the bank is range-sharded:
        # every rank's loss has the same form).  Announcing it lets the K1 backward run right behind the K1 forward
        # on the prior branch, overlapping the decoder, instead of between the loss and the encoder backward.
        self.batch_size = batch_size
        rows = batch_size * (getattr(model, "bank_world", 1) if getattr(model, "bank_group", None) is not None else 1)
        self.g_prior = (torch.full((rows,), -float(beta) / batch_size, dtype=torch.float32, device=dev)
                        if os.environ.get("EXVAE_EAGER_PRIOR_BWD", "1") != "0" else None)
        self.rng_override = rng_override
        self.cache = cache
        # Exemplar prefetch: the next step's N exemplar indices are drawn, and their rows gathered into a persistent
        # [B+N, P] operand, behind the backward of the current step (next to the optimizer), so the step does not start
        # with a 27 us HBM-bound gather in front of the first GEMM.  Every step still draws and encodes a fresh exemplar
        # set.  With injected indices (tests) the tensor must hold the NEXT step's indices when a step runs;
        # ``prime_exemplars()`` (re)loads the first set.
        self.prefetch = (os.environ.get("EXVAE_PREFETCH_EXEMPLARS", "1") != "0" and dataset is not None
                         and getattr(args, "prior", None) == "exemplar_prior" and args.approximate_prior is False
                         and getattr(model, "fuse_exemplar_encoder", False))
        self._pf = None
        if self.prefetch:
            n = model.exemplar_count()
            self._pf = {"rows": torch.zeros(batch_size + n, P, dtype=torch.float32, device=dev),
                        "idx": torch.zeros(n, dtype=torch.int64, device=dev), "B": batch_size}
        self.graph = None
        self.launches_per_step = 0
        if getattr(model, "flat_grads", None) is None:
            from .distributed import FlatGrads
            model.flat_grads = FlatGrads(model.parameters())     # stable grad pointers, one memset per step
        # prior || decoder as parallel graph branches (EXVAE_OVERLAP_SHARDED=0: only with a replicated bank)
        model.overlap_prior = (getattr(model, "bank_group", None) is None
                               or os.environ.get("EXVAE_OVERLAP_SHARDED", "1") != "0")
        model.train()
        saved = self._snapshot()
        self.prime_exemplars()
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
        self._restore(saved)
        self.prime_exemplars()            # from the restored RNG state: the first real step's exemplar set
        if use_graph:
            self.graph = torch.cuda.CUDAGraph()
            # (Negative result, profiles/r2_step_schedule.md: capturing on a high-priority stream and instantiating the
            #  graph with cudaGraphInstantiateFlagUseNodePriority changed nothing measurable; what removed the holes
            #  next to the whole-GPU K1 kernels was to give no launch pending CTAs and to keep the small kernels' CTAs
            #  small enough to fit next to a resident K1 CTA.)
            with torch.cuda.graph(self.graph):
                self._body()
        torch.cuda.synchronize()

    # -------------------------------------------------------------- state kept across the warm-up
    def _snapshot(self):
        dev = self.data.device
        params = [p.detach().clone() for p in self.model.parameters()]
        opt_state = {}
        for group in self.opt.param_groups:
            for p in group['params']:
                st = self.opt.state.get(p)
                if st:
                    opt_state[p] = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
        counters = {gi: ent['step'].clone() for gi, ent in getattr(self.opt, "_tables", {}).items()}
        cache = None if self.cache is None else tuple(c.detach().clone() for c in self.cache)
        return params, opt_state, counters, self.model.rng.counter(dev).clone(), cache

    @torch.no_grad()
    def _restore(self, saved):
        params, opt_state, counters, rng_counter, cache = saved
        for p, v in zip(self.model.parameters(), params):
            p.copy_(v)
        for group in self.opt.param_groups:
            for p in group['params']:
                st = self.opt.state.get(p)
                if not st:
                    continue
                old = opt_state.get(p)
                for k, v in st.items():           # in place: the fused step's pointer table stays valid
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if old is not None else v.zero_()
                    else:
                        st[k] = old[k] if old is not None else 0
        for gi, ent in getattr(self.opt, "_tables", {}).items():
            if gi in counters:
                ent['step'].copy_(counters[gi])
            else:
                ent['step'].fill_(self.opt.initial_step(gi))     # from the restored per-parameter steps
        self.model.rng.counter(self.data.device).copy_(rng_counter)
        if cache is not None:
            for c, v in zip(self.cache, cache):
                c.copy_(v)
        self.model.flat_grads.zero_()
        self.out.zero_()

    def prime_exemplars(self):
        """Draw + gather the exemplar set of the NEXT step now (no-op without prefetch).  Called by the constructor; call
        it again after changing injected indices (``rng_override``) or the RNG state by hand."""
        if self.prefetch:
            ro, model = self.rng_override, self.model
            if ro is not None:
                model.rng_override = {"eps": list(ro.get("eps", [])), "exemplar_indices": ro.get("exemplar_indices")}
            model.prefetch_exemplars(self._pf, self.dataset)
            torch.cuda.current_stream().synchronize()

    def set_beta(self, beta: float):
        """New KL weight for the following steps (no re-capture: the kernels read the device scalar)."""
        self.beta_dev.fill_(float(beta))
        if self.g_prior is not None:
            self.g_prior.fill_(-float(beta) / self.batch_size)

    def _body(self):
        model = self.model
        x = model.rng.bernoulli(self.data) if self.args.dynamic_binarization else self.data
        model.flat_grads.zero_()
        ro = self.rng_override
        if ro is not None:
            model.rng_override = {"eps": list(ro.get("eps", [])), "exemplar_indices": ro.get("exemplar_indices")}
        prev = ops.set_fused_grad_accumulation(True)             # dW/db are added into the flat buffer in-kernel
        prev_defer = ops.set_deferred_dw_finish(os.environ.get("EXVAE_DEFER_DW_FINISH", "1") != "0")
        model.prior_grad_known = self.g_prior
        model._exemplar_prefetch = self._pf
        aux = None
        try:
            loss, RE, KL = model.calculate_loss((x, self.indices), self.beta_dev, average=True, cache=self.cache,
                                                dataset=self.dataset)
            base = loss._base            # calculate_loss(average=True) returns three views of one [3] tensor
            if base is not None and base.numel() == 3:
                # d(loss) = 1 through the [3] result itself: no ones-fill, no zero-fill + copy of a select backward
                torch.autograd.backward(base, grad_tensors=self._g3)
            else:
                loss.backward()
            if self.prefetch:
                # the backward has consumed the fused operand (its last reader, the layer-1 dW GEMM, is on this stream;
                # the deferred dW finishes read only their own workspaces): fill it for the next step on a side
                # branch, next to the gradient exchange / optimizer below (joined at the end of the step)
                cur = torch.cuda.current_stream()
                aux = model._aux_stream()
                aux.wait_stream(cur)
                with torch.cuda.stream(aux):
                    model.prefetch_exemplars(self._pf, self.dataset)
        finally:
            ops.flush_dense_bwd()                  # every parameter gradient is complete from here on
            ops.set_deferred_dw_finish(prev_defer)
            ops.set_fused_grad_accumulation(prev)
            model.prior_grad_known = None
            model._exemplar_prefetch = None
        if model.grad_sync is not None:
            model.grad_sync()                      # data-parallel: all-reduce of the flat gradient buffer
        self.opt.step()
        with torch.no_grad():
            base = loss._base            # calculate_loss(average=True) returns three views of one [3] tensor
            if base is not None and base.numel() == 3 and RE._base is base and KL._base is base:
                self.out.copy_(base.detach())
            else:
                self.out.copy_(torch.stack((loss.detach(), RE.detach(), KL.detach())))
        if aux is not None:
            torch.cuda.current_stream().wait_stream(aux)

    def step(self, data=None, indices=None):
        if data is not None:
            self.data.copy_(data, non_blocking=True)
            self.indices.copy_(indices.reshape(-1, 1), non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.out
