import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops
B, C, D = 512, 25000, 40
rng = np.random.default_rng(B)
mu = torch.tensor(rng.normal(size=(C, D)).astype(np.float32)).cuda()
lv = torch.full((D,), -2.4189).cuda()
src = rng.integers(0, C, size=B)
z = (mu[src] + 0.3 * torch.randn(B, D, device="cuda"))
mu_idx = torch.tensor(rng.integers(0, 50000, size=C)).cuda()
variants = {"nohit": torch.full((B,), 10 ** 9, device="cuda"), "poisson": mu_idx[src].clone(),
            "unique": None}
perm = torch.randperm(C, device="cuda")
variants["unique"] = (perm[:B] + 100000)
mu_idx_u = perm + 100000
for name, zi in variants.items():
    mi = mu_idx_u if name == "unique" else mu_idx
    cnt = (zi[:, None] == mi[None, :]).sum(1)
    print(name, "cnt max", int(cnt.max()), "mean", float(cnt.float().mean()), flush=True)
    for _ in range(3):
        ops.prior_lse(z, mu, lv, zi, mi)
torch.cuda.synchronize()
