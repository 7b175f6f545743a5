"""Accuracy + timing of the exemplar-prior backward (run on a B200): tensor-core path against an fp64 torch
autograd reference, per shape.  EXVAE_PRIOR_BWD=simt selects the FMA-pipe kernel for comparison."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402


def ref64(z, mu, lv, z_idx, mu_idx, gout):
    z, mu, lv = (t.double().cuda().requires_grad_(True) for t in (z, mu, lv))
    sig = torch.exp(0.5 * lv)
    zs, ms = z / sig, mu / sig
    d2 = (zs * zs).sum(1, keepdim=True) + (ms * ms).sum(1)[None] - 2 * zs @ ms.t()
    logn = -0.5 * (lv + np.log(2 * np.pi)).sum() - 0.5 * d2
    if z_idx is not None:
        mask = z_idx.cuda()[:, None] == mu_idx.cuda()[None]
        logn = logn.masked_fill(mask, -float("inf"))
        cnt = mu.shape[0] - mask.sum(1)
    else:
        cnt = torch.full((z.shape[0],), mu.shape[0], device="cuda")
    lp = torch.logsumexp(logn, 1) - torch.log(cnt.double())
    lp.backward(gout.double().cuda())
    return lp.detach(), z.grad, mu.grad, lv.grad


def rel(a, b):
    return ((a.double() - b).abs().max() / (b.abs().max() + 1e-300)).item()


print("prior bwd path:", os.environ.get("EXVAE_PRIOR_BWD", "tensor-core"))
for (B, C, D, masked) in ((512, 25000, 40, True), (100, 1000, 40, True), (37, 333, 24, True), (130, 5000, 63, False),
                          (512, 25000, 40, False), (300, 1438, 40, True)):
    g = torch.Generator().manual_seed(B + C)
    mu = torch.randn(C, D, generator=g)
    lv = torch.full((D,), -2.4189) + 0.1 * torch.randn(D, generator=g)
    src = torch.randint(0, C, (B,), generator=g)
    z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
    mu_idx = torch.randint(0, 50000, (C,), generator=g)
    z_idx = mu_idx[src].clone()
    gout = torch.randn(B, generator=g)
    lp64, dz64, dmu64, dlv64 = ref64(z, mu, lv, z_idx if masked else None, mu_idx, gout)
    zc, mc, lc = (t.cuda().requires_grad_(True) for t in (z, mu, lv))
    zi, mi = (z_idx.cuda(), mu_idx.cuda()) if masked else (None, None)
    lp = ops.prior_lse(zc, mc, lc, zi, mi)
    lp.backward(gout.cuda())
    torch.cuda.synchronize()
    msg = (f"B={B} C={C} D={D} mask={masked}: lp {rel(lp.detach(), lp64):.1e} dz {rel(zc.grad, dz64):.1e} "
           f"dmu {rel(mc.grad, dmu64):.1e} dlogvar {rel(lc.grad, dlv64):.1e}")

    def fb():
        zc.grad = mc.grad = lc.grad = None
        out = ops.prior_lse(zc, mc, lc, zi, mi)
        out.backward(gout.cuda())

    def fw():
        with torch.no_grad():
            ops.prior_lse(zc, mc, lc, zi, mi)

    ts = []
    for fn in (fw, fb):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(msg + f" | fwd {ts[0]:.0f} us, fwd+bwd {ts[1]:.0f} us (eager)", flush=True)
print("prior_bwd_check done")
