"""Launch each hot kernel once at the BASELINE shapes so that `ncu --set full` can capture them in a short run:

  ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3|prior_|knn_fused' -c 40 \
      -o gpurun_out/prof python tools/prof_kernels.py

cfg2: B=512, N=25000, D=40, hidden 300 (K1 fwd/bwd, K3 layer GEMMs); cfg3: K2 at B=100 x 25000 x 40, one
convhvae decoder layer (64 -> 64, 3x3, 28x28, 100 images: implicit GEMM) forward + backward; cfg5: K1 at D=128.
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
x = torch.rand(N + B, P, device="cuda", generator=g)
W1h = torch.randn(H, P, device="cuda", generator=g, requires_grad=True)
W1g = torch.randn(H, P, device="cuda", generator=g, requires_grad=True)
b1 = torch.zeros(H, device="cuda", requires_grad=True)
W2h = torch.randn(H, H, device="cuda", generator=g, requires_grad=True)
W2g = torch.randn(H, H, device="cuda", generator=g, requires_grad=True)
Wm = torch.randn(D, H, device="cuda", generator=g, requires_grad=True)
bm = torch.zeros(D, device="cuda", requires_grad=True)
# cfg5 bank shard: B=512 x 12500 x 128
mu5 = torch.randn(12500, 128, device="cuda", generator=g, requires_grad=True)
z5 = (mu5[:B].detach() + 0.3 * torch.randn(B, 128, device="cuda", generator=g)).requires_grad_(True)
lv5 = torch.full((128,), -2.4189, device="cuda", requires_grad=True)
# cfg3: kNN and one decoder conv layer
zq = torch.randn(100, D, device="cuda", generator=g)
xc = torch.randn(100, 28, 28, 64, device="cuda", generator=g, requires_grad=True)
Wc = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24).requires_grad_(True)
Wcg = (torch.randn(64, 64, 3, 3, device="cuda", generator=g) / 24).requires_grad_(True)
bc = torch.zeros(64, device="cuda", requires_grad=True)
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
    lp5 = ops.prior_lse(z5, mu5, lv5, z_idx, mu_idx[:12500])
    lp5.sum().backward()
    ops.knn_topk(zq, mu.detach(), 10)
    yc = ops.conv2d_gated(xc, Wc, bc, Wcg, bc, 1, 1)
    yc.sum().backward()
torch.cuda.synchronize()
print("done")
