"""Stand-alone timing of the K1 entry points at cfg2 (B=512, N=25000, D=40) and the cfg5 shard (512 x 12500 x 128)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402

def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

g = torch.Generator(device="cuda").manual_seed(0)
for (B, N, D) in ((512, 25000, 40), (4096, 3125, 40), (512, 12500, 128), (5000, 50000, 40)):
    mu = torch.randn(N, D, device="cuda", generator=g)
    z = mu[torch.randint(0, N, (B,), device="cuda", generator=g)] + 0.3 * torch.randn(B, D, device="cuda", generator=g)
    lv = torch.full((D,), -2.4189, device="cuda")
    mi = torch.randint(0, 50000, (N,), device="cuda", generator=g); zi = torch.randint(0, 50000, (B,), device="cuda", generator=g)
    with torch.no_grad():
        f_mask = t(lambda: ops.prior_lse(z, mu, lv, zi, mi))
        f_nomask = t(lambda: ops.prior_lse(z, mu, lv, None, None))
    zr, mr, lr = z.clone().requires_grad_(True), mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    def fb():
        lp = ops.prior_lse(zr, mr, lr, zi, mi); lp.sum().backward(); zr.grad = mr.grad = lr.grad = None
    print(f"B={B} N={N} D={D}: fwd masked {f_mask:.1f} us, unmasked {f_nomask:.1f} us, fwd+bwd {t(fb):.1f} us (eager, incl. launch overhead)", flush=True)
