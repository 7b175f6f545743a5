"""Per-CTA phase trace of the tensor-core exemplar-prior backward (run on a B200)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402
from exemplar_vae_b200._lib import lib  # noqa: E402

B, C, D = 512, 25000, 40
g = torch.Generator().manual_seed(0)
mu = torch.randn(C, D, generator=g).cuda().requires_grad_(True)
lv = torch.full((D,), -2.4189).cuda().requires_grad_(True)
src = torch.randint(0, C, (B,), generator=g)
z = (mu.detach().cpu()[src] + 0.3 * torch.randn(B, D, generator=g)).cuda().requires_grad_(True)
mu_idx = torch.randint(0, 50000, (C,), generator=g).cuda()
z_idx = mu_idx[src.cuda()].clone()
L = lib()
for _ in range(3):
    ops.prior_lse(z, mu, lv, z_idx, mu_idx).sum().backward()
torch.cuda.synchronize()
buf = torch.zeros(8 * 400 * 4, dtype=torch.int64, device="cuda")
L.exvae_gemm_set_trace(buf.data_ptr())
ops.prior_lse(z, mu, lv, z_idx, mu_idx).sum().backward()
torch.cuda.synchronize()
L.exvae_gemm_set_trace(None)
for k, name in enumerate(("pass 1 (lanes = latents)", "pass 2 (lanes = exemplars)")):
    t = buf[k * 3200:(k + 1) * 3200].cpu().numpy().reshape(400, 8)
    t = t[t[:, 0] > 0]
    d = lambda a, b: (t[:, b] - t[:, a]) / 1e3
    print(f"{name}: {len(t)} CTAs, span {(t[:, 5].max() - t[:, 0].min()) / 1e3:.1f} us; per CTA median / max (us)")
    for lab, a, b in (("start -> X tile landed", 0, 1), ("X landed -> first Y tile landed", 1, 2),
                      ("pair loop (first S MMA -> last G MMA issued)", 2, 3), ("last issue -> G complete", 3, 4),
                      ("drain (G complete -> CTA end)", 4, 5), ("whole CTA", 0, 5)):
        v = d(a, b)
        print(f"   {lab:46s} {np.median(v):8.2f} {v.max():8.2f}")
    t0 = t[:, 0].min()
    for sm in (0, 1, 77):
        q = t[t[:, 6] == sm]
        q = q[np.argsort(q[:, 0])]
        print(f"   SM {sm}: " + "  ".join(f"[{(r[0] - t0) / 1e3:.1f} .. {(r[5] - t0) / 1e3:.1f}]" for r in q))
