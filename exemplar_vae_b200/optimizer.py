"""AdamNormGrad — mirror of utils/optimizer.py:7-80 as ONE fused multi-tensor CUDA step
(per-tensor ||g||_2, normalisation, Adam moments, bias correction, update)."""
from __future__ import annotations

import torch
from torch.optim import Optimizer

from . import ops
from ._lib import lib


class AdamNormGrad(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}

    def _table(self, gi, group):
        """Device table {param, grad, exp_avg, exp_avg_sq, numel} for one param group; rebuilt only
        when a pointer changes (never inside a captured graph: grads are kept in place)."""
        rows = []
        for p in group['params']:
            st = self.state[p]
            if len(st) == 0:
                st['step'] = 0
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if st['exp_avg'].shape != p.shape or st['exp_avg_sq'].shape != p.shape:
                raise RuntimeError(
                    f"AdamNormGrad: optimizer state of shape {tuple(st['exp_avg'].shape)} does not match its parameter "
                    f"of shape {tuple(p.shape)} (state_dict loaded with a different parameter order?)")
            g = p.grad
            rows.append((p.data_ptr(), g.data_ptr() if g is not None else 0, st['exp_avg'].data_ptr(),
                         st['exp_avg_sq'].data_ptr(), p.numel()))
        key = tuple(rows)
        ent = self._tables.get(gi)
        if ent is None or ent['key'] != key:
            dev = group['params'][0].device
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("AdamNormGrad: parameter/grad pointers changed during CUDA-graph capture; "
                                   "call zero_grad(set_to_none=False) and run one eager step first")
            ent = {
                'key': key,
                'table': torch.tensor(rows, dtype=torch.int64).to(dev),
                'norms': torch.empty(16 * len(rows) + 4, dtype=torch.float32, device=dev),   # 16 partials per tensor + step size
                'step': (ent['step'] if ent is not None else
                         torch.full((1,), self.initial_step(gi), dtype=torch.int64, device=dev)),
                'max_numel': max(r[4] for r in rows),
            }
            self._tables[gi] = ent
        return ent

    def initial_step(self, gi: int) -> int:
        """Step count a fresh device counter of group ``gi`` starts from: the largest per-parameter ``step`` of the
        (possibly just loaded) state.  The reference keeps one count per parameter (utils/optimizer.py:60); they are
        all equal whenever every parameter receives a gradient each step, which holds for the in-scope models."""
        steps = [int(self.state[p].get('step', 0)) for p in self.param_groups[gi]['params'] if len(self.state[p])]
        return max(steps) if steps else 0

    def zero_grad(self, set_to_none: bool = False):
        """Keeps gradient buffers in place (stable pointers for the fused step and for CUDA graphs)."""
        super().zero_grad(set_to_none=set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        for gi, group in enumerate(self.param_groups):
            if not group['params']:
                continue
            for p in group['params']:
                if not p.is_cuda:
                    raise RuntimeError("AdamNormGrad (exvae_b200) needs CUDA parameters: there is no CPU path")
            ent = self._table(gi, group)
            b1, b2 = group['betas']
            L.check(L.exvae_adam_normgrad_step(ent['table'].data_ptr(), len(group['params']), ent['max_numel'],
                                               group['lr'], b1, b2, group['eps'], group['weight_decay'],
                                               ent['step'].data_ptr(), ent['norms'].data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "adam_normgrad_step")
            ops._count(2)
        return loss

    def state_dict(self):
        # mirror the device step counter into the per-parameter 'step' the reference checkpoints carry
        for gi, group in enumerate(self.param_groups):
            ent = self._tables.get(gi)
            if ent is not None:
                s = int(ent['step'].item())
                for p in group['params']:
                    if p in self.state and p.grad is not None:
                        self.state[p]['step'] = s
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}
