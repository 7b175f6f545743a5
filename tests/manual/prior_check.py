"""Accuracy + timing of the exemplar-prior forward/backward at the cfg2 shape (run on a B200)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exemplar_vae_b200 import ops  # noqa: E402
from oracle import exvae_oracle as O  # noqa: E402

for (B, C, D) in ((512, 25000, 40), (100, 1000, 40), (37, 333, 24), (130, 5000, 63)):
    rng = np.random.default_rng(B)
    mu = rng.normal(size=(C, D)).astype(np.float32)
    lv = np.full((D,), -2.4189, dtype=np.float32)
    src = rng.integers(0, C, size=B)
    z = (mu[src] + np.exp(0.5 * lv) * rng.normal(size=(B, D))).astype(np.float32)
    mu_idx = rng.integers(0, 50000, size=C).astype(np.int64)
    z_idx = mu_idx[src].copy()
    f64 = O.log_p_z_exemplar_lse_f64(z[:64], z_idx[:64], mu, lv, mu_idx)
    ref = O.log_p_z_exemplar_lse_np(z, z_idx, mu, np.tile(lv, (C, 1)), mu_idx, test=False)
    zc, mc, lc = torch.tensor(z).cuda(), torch.tensor(mu).cuda(), torch.tensor(lv).cuda()
    zi, mi = torch.tensor(z_idx).cuda(), torch.tensor(mu_idx).cuda()
    got = ops.prior_lse(zc, mc, lc, zi, mi).cpu().numpy()
    ok = np.isfinite(f64)
    e64 = np.max(np.abs(got[:64][ok] - f64[ok]) / np.abs(f64[ok]))
    eref = np.max(np.abs(got - ref)[np.isfinite(ref)] / np.abs(ref[np.isfinite(ref)]))
    got_t = ops.prior_lse(zc, mc, lc).cpu().numpy()
    ref_t = O.log_p_z_exemplar_lse_np(z, None, mu, np.tile(lv, (C, 1)), mu_idx, test=True)
    et = np.max(np.abs(got_t - ref_t) / np.abs(ref_t))
    for _ in range(3):
        ops.prior_lse(zc, mc, lc, zi, mi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.prior_lse(zc, mc, lc, zi, mi)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} C={C} D={D}: rel_err vs fp64 {e64:.2e}, vs reference-formula {eref:.2e}, test-mode {et:.2e}, "
          f"fwd {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/call (stage+mask+main+merge+finalize, eager launch)", flush=True)
print("prior_check done")
