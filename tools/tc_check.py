"""Quick numerical check of the tcgen05 3xTF32 GEMM variants against fp64 (run on a B200)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402

print("backend:", ops.gemm_backend(), flush=True)


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b.cpu()).abs().max() / (b.abs().max() + 1e-30)).item()


def check(R, K, O, which):
    g = torch.Generator().manual_seed(R + K + O)
    x = torch.randn(R, K, generator=g)
    W = torch.randn(O, K, generator=g) / K ** 0.5
    Wg = torch.randn(O, K, generator=g) / K ** 0.5
    b = torch.randn(O, generator=g)
    dout = torch.randn(R, O, generator=g)
    xd, Wd, Wgd, bd = (t.double().requires_grad_(True) for t in (x, W, Wg, b))
    xc, Wc, Wgc, bc = (t.cuda().requires_grad_(True) for t in (x, W, Wg, b))
    if which == "linear":
        ref = xd @ Wd.t() + bd
        out = ops.linear(xc, Wc, bc)
    else:
        ref = (xd @ Wd.t() + bd) * torch.sigmoid(xd @ Wgd.t() + bd)
        out = ops.gated_dense(xc, Wc, bc, Wgc, bc)
    torch.cuda.synchronize()
    e_f = rel(out, ref.detach())
    print(f"{which:7s} R={R:6d} K={K:4d} O={O:4d} fwd_err={e_f:.2e}", end=" ", flush=True)
    ref.backward(dout.double())
    out.backward(dout.cuda())
    torch.cuda.synchronize()
    print(f"dx_err={rel(xc.grad, xd.grad):.2e} dW_err={rel(Wc.grad, Wd.grad):.2e} db_err={rel(bc.grad, bd.grad):.2e}",
          flush=True)


for which in ("linear", "gated"):
    for shp in ((128, 32, 128), (256, 64, 128), (300, 300, 40), (512, 784, 300), (1000, 40, 300), (25512, 784, 300),
                (12, 196, 24)):
        check(*shp, which)
print("tc_check done")
