"""Per-CTA phase / wait trace of the fused exemplar-prior forward kernel (prior_fused.cu); run on a B200.

The trace stores are compiled in only with -DEXVAE_PF_TRACE (production builds carry no trace code: a %globaltimer
store guarded by a kernel-parameter null check faulted under the kNN-mode test, profiles/r2_step_schedule.md):
    NVCC_EXTRA=-DEXVAE_PF_TRACE python tools/build_lib.py --force && python tools/prior_fwd_trace.py

words per CTA: 0 start, 1 prologue done, 2 converter cycles waiting for a free stage, 3 MMA cycles waiting for a converted
stage, 4 MMA cycles waiting for a drained accumulator, 5 epilogue cycles waiting for an accumulator, 6 main loop done,
7 end (top bit: the CTA that merged its row block)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exemplar_vae_b200 import ops  # noqa: E402
from exemplar_vae_b200._lib import lib  # noqa: E402

L = lib()
g = torch.Generator().manual_seed(0)
for (B, C, D, grad) in ((512, 25000, 40, False), (512, 25000, 40, True), (4096, 3125, 40, False), (5000, 50000, 40, False)):
    mu = torch.randn(C, D, generator=g).cuda().requires_grad_(grad)
    lv = torch.full((D,), -2.4189).cuda()
    src = torch.randint(0, C, (B,), generator=g)
    z = (mu.detach().cpu()[src] + 0.3 * torch.randn(B, D, generator=g)).cuda()
    mu_idx = torch.randint(0, 50000, (C,), generator=g).cuda()
    z_idx = mu_idx[src.cuda()].clone()
    for _ in range(3):
        ops.prior_lse(z, mu, lv, z_idx, mu_idx)
    torch.cuda.synchronize()
    buf = torch.zeros(8 * 400, dtype=torch.int64, device="cuda")
    L.exvae_gemm_set_trace(buf.data_ptr())
    ops.prior_lse(z, mu, lv, z_idx, mu_idx)
    torch.cuda.synchronize()
    L.exvae_gemm_set_trace(None)
    t = buf.cpu().numpy().reshape(400, 8)
    t = t[t[:, 0] > 0]
    end = t[:, 7] & ((1 << 63) - 1)
    ntile = -(-C // 128); rbs = -(-B // 128); nsplit = max(1, min(148 // rbs, ntile))
    clk = 1.965e3                                     # cycles per us
    print(f"B={B} C={C} D={D} staging={'on' if grad else 'off'}: {len(t)} CTAs, ~{ntile / nsplit:.1f} tiles per CTA, "
          f"kernel span {(end.max() - t[:, 0].min()) / 1e3:.1f} us")
    for lab, v in (("prologue", (t[:, 1] - t[:, 0]) / 1e3), ("main loop", (t[:, 6] - t[:, 1]) / 1e3),
                   ("tail (ticket / merge)", (end - t[:, 6]) / 1e3), ("whole CTA", (end - t[:, 0]) / 1e3),
                   ("start skew", (t[:, 0] - t[:, 0].min()) / 1e3),
                   ("converter waits for a free stage", t[:, 2] / clk), ("MMA waits for a converted stage", t[:, 3] / clk),
                   ("MMA waits for a drained accumulator", t[:, 4] / clk), ("epilogue waits for an accumulator", t[:, 5] / clk)):
        print(f"   {lab:40s} median {np.median(v):8.2f}  max {v.max():8.2f} us")
