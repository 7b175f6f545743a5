"""Per-CTA phase trace of the tensor-core GEMM (run on a B200): where does a tile's time go?

usage: python tools/gemm_trace.py [R K O]   (gated forward, then its backward)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402
from exemplar_vae_b200._lib import lib  # noqa: E402

R, K, O = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (25512, 784, 300)
g = torch.Generator().manual_seed(0)
x = torch.randn(R, K, generator=g).cuda().requires_grad_(True)
Wh = (torch.randn(O, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
Wg = (torch.randn(O, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
b = torch.zeros(O).cuda().requires_grad_(True)
dout = torch.randn(R, O, generator=g).cuda()
L = lib()


def summarize(name, buf, launch):
    t = buf[launch * 8 * 160: (launch + 1) * 8 * 160].cpu().numpy().astype(np.int64).reshape(160, 8)
    t = t[t[:, 0] > 0]
    if len(t) == 0:
        print(name, ": no trace")
        return
    span = (t[:, 6].max() - t[:, 0].min()) / 1e3
    life = (t[:, 6] - t[:, 0]) / 1e3
    print(f"{name}: {len(t)} persistent CTAs, kernel span {span:.1f} us; per CTA (median / max):")
    print(f"   tiles per CTA                      {np.median(t[:, 2]):8.0f} {t[:, 2].max():8.0f}")
    print(f"   CTA lifetime (us)                  {np.median(life):8.2f} {life.max():8.2f}")
    ghz = 1.965e3          # cycles per us (SM clock under load on the B200 boxes)
    for col, what in ((1, "producer waiting for a free stage"), (5, "converter warp waiting for TMA data"),
                      (3, "MMA thread waiting for converted operands"), (4, "MMA thread waiting for a drained accumulator")):
        v = t[:, col] / ghz
        print(f"   {what:44s} {np.median(v):8.2f} {v.max():8.2f}  us")


def timed(fn, n=20):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fn()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) * 1e3 / n


for _ in range(int(os.environ.get("WARM", "100"))):
    out = ops.gated_dense(x, Wh, b, Wg, b)
torch.cuda.synchronize()
buf = torch.zeros(8 * 160 * 8, dtype=torch.int64, device="cuda")
L.exvae_gemm_set_trace(buf.data_ptr())
out = ops.gated_dense(x, Wh, b, Wg, b)
out.backward(dout)
torch.cuda.synchronize()
L.exvae_gemm_set_trace(None)
summarize(f"gated fwd R={R} K={K} O={O}", buf, 0)
summarize("gated bwd dx", buf, 1)
summarize("gated bwd dW", buf, 2)
print("fwd entry point: %.1f us per call" % timed(lambda: ops.gated_dense(x, Wh, b, Wg, b)))


def fb():
    o = ops.gated_dense(x, Wh, b, Wg, b)
    o.backward(dout)


print("fwd+bwd: %.1f us per call" % timed(fb))
x.requires_grad_(False)
print("fwd+bwd without dx: %.1f us per call" % timed(fb))
