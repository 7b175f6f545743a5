"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard ranges, the associative
merge of per-shard log-sum-exp partials after an all-gather, and the flat gradient all-reduce.
The CUDA kernels themselves are covered by the -m gpu tests and tests/manual/mgpu_check.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import exvae_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exemplar_vae_b200.distributed import FlatGrads, shard_range
    rng = np.random.default_rng(0)                      # same data on both ranks
    B, C, D, T = 12, 101, 40, 60
    mu = rng.normal(size=(C, D)).astype(np.float32)
    lv = np.full((C, D), -1.1, dtype=np.float32)
    src = rng.integers(0, C, size=B)
    z = (mu[src] + 0.4 * rng.normal(size=(B, D))).astype(np.float32)
    mu_idx = rng.integers(0, T, size=C).astype(np.int64); z_idx = mu_idx[src].copy()
    lo, hi = shard_range(C, world, rank)
    # per-shard partials (what K1 emits): max, sum-of-exp, masked count of the UN-normalised logits
    prob = O.log_p_z_exemplar_np(z, z_idx, mu[lo:hi], lv[lo:hi], mu_idx[lo:hi], test=False)
    cnt = (z_idx.reshape(-1, 1) == mu_idx[lo:hi].reshape(1, -1)).sum(1).astype(np.float64)
    logits = prob.astype(np.float64) + np.log(float(hi - lo) - cnt)[:, None]
    m = logits.max(1)
    s = np.where(np.isneginf(m), 0.0, np.exp(logits - np.where(np.isneginf(m), 0.0, m)[:, None]).sum(1))
    mine = torch.tensor(np.stack([m, s, cnt], 1))
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    g = torch.stack(gathered).numpy()
    M, S = O.merge_lse_partials_np(g[:, :, 0], g[:, :, 1])
    merged = M + np.log(S) - np.log(C - g[:, :, 2].sum(0))
    full = O.log_p_z_exemplar_lse_np(z, z_idx, mu, lv, mu_idx, test=False)
    ok1 = np.allclose(merged, full, rtol=1e-5)
    # flat gradient all-reduce (mean)
    ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))]
    fg = FlatGrads(ps)
    ps[0].grad += (rank + 1); ps[1].grad += 10 * (rank + 1)
    fg.all_reduce_mean()
    ok2 = torch.allclose(ps[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(ps[1].grad, torch.full((5,), 15.0))
    ok3 = ps[0].grad.data_ptr() == fg.buf.data_ptr()
    # bucketed, overlapped gradient averaging: gradients reported out of registration order, one never reported
    from exemplar_vae_b200.distributed import GradBuckets
    qs = [torch.nn.Parameter(torch.zeros(n)) for n in (7, 40, 3, 50, 20)]
    fq = FlatGrads(qs, n_buckets=3, align=8)          # bucket starts / ends padded to 8 floats (multimem slices)
    gb = GradBuckets(fq, None)
    spans = gb.ranges
    ok4 = (spans[0][0] == 0 and spans[-1][1] == fq.buf.numel() and len(spans) == 3
           and all(a[1] == b[0] and a[1] % 8 == 0 for a, b in zip(spans, spans[1:])))
    # layer-aware cuts: parameters 0-1 form one layer, 2-4 another -> a 2-bucket split may only cut between them
    rs = [torch.nn.Parameter(torch.zeros(n)) for n in (7, 40, 3, 50, 20)]
    fr = FlatGrads(rs, n_buckets=2, groups=[0, 0, 1, 1, 1])
    ok4 = ok4 and fr.ranges == [(0, 47), (47, 120)] and [fr.bucket_of[i] for i in range(5)] == [0, 0, 1, 1, 1]
    for step in range(2):                                   # twice: the bucket state resets after finish()
        fq.zero_()
        for i, q in enumerate(qs):
            q.grad += (rank + 1) * (i + 1)
        for i in (4, 3, 1, 0):                              # parameter 2 "receives no gradient" this step
            gb.ready(qs[i].grad)
        gb.finish()
        ok4 = ok4 and all(torch.allclose(q.grad, torch.full_like(q, 1.5 * (i + 1))) for i, q in enumerate(qs))
    ret[rank] = bool(ok1 and ok2 and ok3 and ok4)
    dist.destroy_process_group()


def test_shard_range_partitions():
    from exemplar_vae_b200.distributed import shard_range
    for n, w in ((25000, 8), (11500, 8), (7, 3), (5, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_lse_partial_merge_and_flat_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
