"""Correctness + timing of the big forward GEMMs (EXVAE_GEMM_PAIR=0/1): python tools/pair_check.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402

torch.manual_seed(0)
for (R, K, O) in ((25512, 784, 300), (25512, 300, 300), (3637, 784, 300), (25512, 300, 40), (2000, 96, 296)):
    x = torch.randn(R, K, device="cuda")
    Wh, Wg = torch.randn(O, K, device="cuda") / K ** 0.5, torch.randn(O, K, device="cuda") / K ** 0.5
    bh, bg = torch.randn(O, device="cuda"), torch.randn(O, device="cuda")
    out = ops.gated_dense(x, Wh, bh, Wg, bg)
    ref = ((x.double() @ Wh.double().t() + bh.double()) * torch.sigmoid(x.double() @ Wg.double().t() + bg.double()))
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    lin = ops.linear(x, Wh, bh, 1)
    refl = torch.sigmoid(x.double() @ Wh.double().t() + bh.double())
    errl = float((lin.double() - refl).abs().max())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gated_dense(x, Wh, bh, Wg, bg)
    e0.record()
    for _ in range(20):
        ops.gated_dense(x, Wh, bh, Wg, bg)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    print(f"pair={os.environ.get('EXVAE_GEMM_PAIR','0')} R={R} K={K} O={O}: gated err {err:.2e} linear err {errl:.2e}  gated fwd {us:.1f} us "
          f"({3 * 2.0 * R * K * 2 * O / us / 1e6:.0f} issued TFLOP/s)", flush=True)
