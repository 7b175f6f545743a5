"""Launch each hot kernel once at the cfg2 shapes (B=512, N=25000, D=40, hidden 300) so that
`ncu --set full` can capture them in a short run:

  ncu --set full --clock-control none --import-source on -k regex:'prior_lse|sgemm' -c 12 \
      -o gpurun_out/prof python tools/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402

B, N, D, H, P = 512, 25000, 40, 300, 784
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator(device="cuda").manual_seed(0)
mu = torch.randn(N, D, device="cuda", generator=g, requires_grad=True)
z = (mu[:B].detach() + 0.3 * torch.randn(B, D, device="cuda", generator=g)).requires_grad_(True)
lv = torch.full((D,), -2.4189, device="cuda", requires_grad=True)
mu_idx = torch.randint(0, 50000, (N,), device="cuda", generator=g)
z_idx = mu_idx[:B].clone()
x = torch.rand(N, P, device="cuda", generator=g)
W1h = torch.randn(H, P, device="cuda", generator=g, requires_grad=True)
W1g = torch.randn(H, P, device="cuda", generator=g, requires_grad=True)
b1 = torch.zeros(H, device="cuda", requires_grad=True)
W2h = torch.randn(H, H, device="cuda", generator=g, requires_grad=True)
W2g = torch.randn(H, H, device="cuda", generator=g, requires_grad=True)
Wm = torch.randn(D, H, device="cuda", generator=g, requires_grad=True)
bm = torch.zeros(D, device="cuda", requires_grad=True)
for _ in range(reps):
    lp = ops.prior_lse(z, mu, lv, z_idx, mu_idx)
    lp.sum().backward()
    h1 = ops.gated_dense(x, W1h, b1, W1g, b1)
    h2 = ops.gated_dense(h1, W2h, b1, W2g, b1)
    m = ops.linear(h2, Wm, bm)
    m.sum().backward()
    xs = x[:B]
    d1 = ops.gated_dense(xs, W1h, b1, W1g, b1)
    d1.sum().backward()
torch.cuda.synchronize()
print("done")
