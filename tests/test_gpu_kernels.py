"""GPU parity tests of the C-ABI kernels (through exemplar_vae_b200.ops) against the CPU oracle
and the reference-generated golden fixtures.  Tolerances: indices bit-exact; fp32 quantities
within 1e-4 relative of the reference (north_star), usually far tighter."""
import numpy as np
import pytest
import torch

from oracle import exvae_oracle as O

pytestmark = pytest.mark.gpu

TAGS = ("s", "r", "d24")


def dev(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def close(a, b, rtol, atol=0.0):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def ops():
    from exemplar_vae_b200 import ops
    return ops


# ------------------------------------------------------------------ a1/a2/a3 matrices
def test_pairwise_distance_matches_reference_bits(ops, golden):
    g = golden("prior")
    for t in TAGS:
        d = ops.pairwise_distance(dev(g[f"{t}:z"]), dev(g[f"{t}:mu"])).cpu().numpy()
        ref = g[f"{t}:pairwise_distance"]
        assert np.mean(d != ref) < 1e-3, "fp64-accumulate + single rounding should reproduce the reference bits"
        close(d, ref, rtol=3e-7)


def test_log_normal_diag_vectorized(ops, golden):
    g = golden("prior")
    for t in TAGS:
        D = g[f"{t}:mu"].shape[1]
        lv = dev(g[f"{t}:lv"]).expand(D).contiguous()
        ln, pd = ops.log_normal_diag_vectorized(dev(g[f"{t}:z"]), dev(g[f"{t}:mu"]), lv)
        close(pd, g[f"{t}:pair_dist_scaled"], rtol=2e-6, atol=1e-4)
        close(ln, g[f"{t}:log_normal"], rtol=2e-6, atol=1e-4)


def test_prior_logprob_matrix(ops, golden):
    g = golden("prior")
    for t in TAGS:
        D = g[f"{t}:mu"].shape[1]
        lv = dev(g[f"{t}:lv"]).expand(D).contiguous()
        prob = ops.prior_logprob_matrix(dev(g[f"{t}:z"]), dev(g[f"{t}:mu"]), lv, dev(g[f"{t}:z_idx"]),
                                        dev(g[f"{t}:mu_idx"])).cpu().numpy()
        ref = g[f"{t}:prob_train"]
        assert np.array_equal(np.isneginf(prob), np.isneginf(ref))
        fin = np.isfinite(ref)
        close(prob[fin], ref[fin], rtol=2e-6, atol=1e-4)


# ------------------------------------------------------------------ K1
def test_prior_lse_forward_golden(ops, golden):
    g = golden("prior")
    for t in TAGS:
        D = g[f"{t}:mu"].shape[1]
        lv = dev(g[f"{t}:lv"]).expand(D).contiguous()
        z, mu = dev(g[f"{t}:z"]), dev(g[f"{t}:mu"])
        lp = ops.prior_lse(z, mu, lv, dev(g[f"{t}:z_idx"]), dev(g[f"{t}:mu_idx"]))
        close(lp, g[f"{t}:lse_train"], rtol=5e-5)         # north_star bar: 1e-4 (3xTF32 tensor-core path ~1e-6..3e-5)
        lp_test = ops.prior_lse(z, mu, lv, None, None)
        close(lp_test, g[f"{t}:lse_test"], rtol=5e-5)


def test_prior_lse_backward_golden(ops, golden):
    g = golden("prior")
    for t in TAGS:
        D = g[f"{t}:mu"].shape[1]
        z = dev(g[f"{t}:z"]).requires_grad_(True)
        mu = dev(g[f"{t}:mu"]).requires_grad_(True)
        lvs = dev(g[f"{t}:lv"]).requires_grad_(True)
        lp = ops.prior_lse(z, mu, lvs.expand(D), dev(g[f"{t}:z_idx"]), dev(g[f"{t}:mu_idx"]))
        (lp * dev(g[f"{t}:w"])).sum().backward()
        # gradients are O(10); 1e-4 absolute is 1e-5 of their scale (fp32 round-off of the reference itself)
        close(z.grad, g[f"{t}:dz"], rtol=2e-4, atol=5e-4)
        close(mu.grad, g[f"{t}:dmu"], rtol=2e-4, atol=5e-4)
        close(lvs.grad, g[f"{t}:dlv"], rtol=3e-4, atol=5e-3)


@pytest.mark.parametrize("B,C,D", [(100, 1000, 40), (512, 25000, 40), (130, 5000, 128), (1, 1, 4), (257, 777, 7)])
def test_prior_lse_vs_oracle_sizes(ops, B, C, D):
    rng = np.random.default_rng(B + C + D)
    mu = rng.normal(size=(C, D)).astype(np.float32)
    lv = np.full((D,), -2.4189, dtype=np.float32)
    src = rng.integers(0, C, size=B)
    z = (mu[src] + np.exp(0.5 * lv) * rng.normal(size=(B, D))).astype(np.float32)
    T = 50000
    mu_idx = rng.integers(0, T, size=C).astype(np.int64)
    z_idx = mu_idx[src].copy()
    if C > 1:
        ref = O.log_p_z_exemplar_lse_np(z, z_idx, mu, np.tile(lv, (C, 1)), mu_idx, test=False)
        got = ops.prior_lse(dev(z), dev(mu), dev(lv), dev(z_idx), dev(mu_idx)).cpu().numpy()
        ok = np.isfinite(ref)
        close(got[ok], ref[ok], rtol=1e-4)
        # exact fp64 value: the CUDA path must sit inside the north_star tolerance of the truth too
        f64 = O.log_p_z_exemplar_lse_f64(z[:64], z_idx[:64], mu, lv, mu_idx)
        ok = np.isfinite(f64)
        close(got[:64][ok], f64[ok], rtol=1e-4)
    ref = O.log_p_z_exemplar_lse_np(z, None, mu, np.tile(lv, (C, 1)), mu_idx, test=True)
    got = ops.prior_lse(dev(z), dev(mu), dev(lv)).cpu().numpy()
    close(got, ref, rtol=1e-4)


def test_prior_lse_mask_list_overflow(ops):
    """Rows whose dataset index appears more often than the fast mask list holds (heavy duplication)."""
    rng = np.random.default_rng(7)
    B, C, D = 70, 3000, 40
    mu = rng.normal(size=(C, D)).astype(np.float32)
    lv = np.full((D,), -1.0, dtype=np.float32)
    z = (mu[:B] + 0.3 * rng.normal(size=(B, D))).astype(np.float32)
    mu_idx = rng.integers(0, 40, size=C).astype(np.int64)        # ~75 copies of every index
    z_idx = rng.integers(0, 60, size=B).astype(np.int64)          # some rows match nothing, most match ~75 columns
    ref = O.log_p_z_exemplar_lse_np(z, z_idx, mu, np.tile(lv, (C, 1)), mu_idx, test=False)
    got = ops.prior_lse(dev(z), dev(mu), dev(lv), dev(z_idx), dev(mu_idx)).cpu().numpy()
    close(got, ref, rtol=1e-4)


def test_prior_lse_shard_merge_property(ops):
    """Size-independent property at the full cfg2 size: LSE over a range-sharded bank, merged with
    the library's own finalize, equals the single-shard result."""
    from exemplar_vae_b200._lib import lib
    L = lib()
    B, C, D, G = 512, 25000, 40, 4
    gen = torch.Generator(device="cuda").manual_seed(5)
    mu = torch.randn(C, D, device="cuda", generator=gen)
    z = mu[torch.randint(0, C, (B,), device="cuda", generator=gen)] + 0.3 * torch.randn(B, D, device="cuda", generator=gen)
    lv = torch.full((D,), -2.4189, device="cuda")
    mu_idx = torch.randint(0, 50000, (C,), device="cuda", generator=gen)
    z_idx = mu_idx[torch.randint(0, C, (B,), device="cuda", generator=gen)]
    full = ops.prior_lse(z, mu, lv, z_idx, mu_idx)
    stats = torch.empty(G, B, 4, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    per = C // G
    for r in range(G):
        sl = slice(r * per, (r + 1) * per if r < G - 1 else C)
        m, mi = mu[sl].contiguous(), mu_idx[sl].contiguous()
        ws = torch.empty(L.exvae_prior_lse_workspace_bytes(B, m.shape[0], D), dtype=torch.uint8, device="cuda")
        L.check(L.exvae_prior_lse_fwd(z.data_ptr(), m.data_ptr(), lv.data_ptr(), z_idx.data_ptr(), mi.data_ptr(), B,
                                      m.shape[0], D, None, stats[r].data_ptr(), 0, None, None, ws.data_ptr(), ws.numel(), st))
    lp = torch.empty(B, device="cuda"); l2 = torch.empty(B, device="cuda")
    L.check(L.exvae_prior_lse_finalize(stats.data_ptr(), G, z.data_ptr(), lv.data_ptr(), B, D, C, None, lp.data_ptr(),
                                       l2.data_ptr(), st))
    close(lp, full, rtol=2e-5, atol=2e-4)     # same pairs, different summation order (in-kernel final vs merge kernel)


def test_prior_lse_backward_vs_oracle_cfg1(ops):
    B, C, D = 100, 1000, 40
    g = torch.Generator().manual_seed(3)
    mu = torch.randn(C, D, generator=g); lv = torch.tensor([-1.7])
    src = torch.randint(0, C, (B,), generator=g)
    z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
    mu_idx = torch.randint(0, 2000, (C,), generator=g); z_idx = mu_idx[src].clone()
    w = torch.randn(B, generator=g)
    zc, mc, lc = z.clone().requires_grad_(True), mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    lp = O.t_log_p_z_exemplar(zc, z_idx, mc, lc.expand(D), mu_idx, masked=True)
    (lp * w).sum().backward()
    zg, mg, lg = z.cuda().requires_grad_(True), mu.cuda().requires_grad_(True), lv.cuda().requires_grad_(True)
    lpg = ops.prior_lse(zg, mg, lg.expand(D), z_idx.cuda(), mu_idx.cuda())
    close(lpg, lp, rtol=5e-5)
    (lpg * w.cuda()).sum().backward()
    close(zg.grad, zc.grad, rtol=2e-4, atol=5e-4)
    close(mg.grad, mc.grad, rtol=2e-4, atol=5e-4)
    close(lg.grad, lc.grad, rtol=3e-4, atol=5e-3)


@pytest.mark.parametrize("B,C,D,masked", [(512, 25000, 40, True), (37, 333, 24, True), (130, 5000, 63, False),
                                          (300, 1438, 40, True), (70, 900, 128, True), (5, 3, 8, False),
                                          (4096, 3125, 40, True), (1000, 2900, 40, False),
                                          (512, 12500, 128, True), (300, 3000, 64, False), (100, 700, 512, True),
                                          (1024, 6250, 128, True), (2048, 14500, 40, True)])
def test_prior_lse_backward_vs_fp64_sizes(ops, B, C, D, masked):
    """K1 backward (tensor-core path for D <= 63, FMA-pipe path above) against fp64 torch autograd of the
    reference formula, incl. the cfg2 and cfg4-shard sizes, ragged tiles and the leave-one-out mask; the last two
    shapes are the 8-GPU shard geometry (all-gathered 8 x 512 rows against 25 000 / 8 exemplars), where pass 2 of
    the backward splits the row blocks over CTAs.  D >= 64 (cfg5's D=128 bank shard 512 x 12 500, the single_conv
    latent D=512) runs through the persistent tcgen05 GEMM kernel with the LSE / weight epilogues.  The last shape
    (16 row blocks x 114 bank tiles > 12 tiles per CTA) takes the STAGED forward (prior_lse.cu:fused_fwd_path) followed
    by the tensor-core backward on the operands it staged."""
    g = torch.Generator().manual_seed(B + C + D)
    mu = torch.randn(C, D, generator=g)
    lv = torch.full((D,), -2.4189) + 0.1 * torch.randn(D, generator=g)
    src = torch.randint(0, C, (B,), generator=g)
    z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
    mu_idx = torch.randint(0, 50000, (C,), generator=g)
    z_idx = mu_idx[src].clone()
    if masked and C <= B:       # keep at least one unmasked exemplar per row
        z_idx = z_idx + 100000
    gout = torch.randn(B, generator=g)
    zd, md, ld = (t.double().requires_grad_(True) for t in (z, mu, lv))
    lp = O.t_log_p_z_exemplar(zd, z_idx.view(-1, 1), md, ld, mu_idx, masked=masked)
    lp.backward(gout.double())
    zc, mc, lc = (t.cuda().requires_grad_(True) for t in (z, mu, lv))
    got = ops.prior_lse(zc, mc, lc, z_idx.cuda() if masked else None, mu_idx.cuda() if masked else None)
    close(got, lp.detach(), rtol=1e-4)
    got.backward(gout.cuda())
    for c, t in ((zc, zd), (mc, md), (lc, ld)):
        scale = t.grad.abs().max().item()
        close(c.grad, t.grad, rtol=1e-3, atol=5e-4 * scale)


def test_row_block_views_backward(ops):
    """ops.split_rows / ops.shared_rows: same values and gradients as plain slicing."""
    g = torch.Generator().manual_seed(0)
    t0 = torch.randn(50, 12, generator=g)
    w1, w2 = torch.randn(20, 12, generator=g).cuda(), torch.randn(30, 12, generator=g).cuda()
    a = t0.clone().cuda().requires_grad_(True)
    x, y = ops.split_rows(a * 1.0, 20)
    ((x * w1).sum() + (y * w2).sum()).backward()
    b = t0.clone().cuda().requires_grad_(True)
    ((b[:20] * w1).sum() + (b[20:] * w2).sum()).backward()
    assert torch.equal(a.grad, b.grad)
    a = t0.clone().cuda().requires_grad_(True)
    full, head = ops.shared_rows(a * 1.0, 20)
    wf = torch.randn(50, 12, generator=g).cuda()
    ((full * wf).sum() + (head * w1).sum()).backward()
    b = t0.clone().cuda().requires_grad_(True)
    ((b * wf).sum() + (b[:20] * w1).sum()).backward()
    close(a.grad, b.grad, rtol=1e-6)
    # only one consumer
    a = t0.clone().cuda().requires_grad_(True)
    x, y = ops.split_rows(a * 1.0, 20)
    (y * w2).sum().backward()
    assert torch.equal(a.grad[:20], torch.zeros(20, 12, device="cuda")) and torch.equal(a.grad[20:], w2)


def test_gemm_propagates_non_finite(ops):
    """The in-kernel hi/lo split must not hide NaN / inf operands (they poison the affected outputs only)."""
    x = torch.randn(300, 64, device="cuda")
    W = torch.randn(128, 64, device="cuda") / 8
    x[7, 3] = float("nan")
    x[100, 9] = float("inf")
    out = ops.linear(x, W, None)
    bad = ~torch.isfinite(out)
    assert bad[7].all() and bad[100].all()
    assert int(bad.any(dim=1).sum().item()) == 2


# ------------------------------------------------------------------ K2
def test_knn_topk_bit_exact_golden(ops, golden):
    g = golden("knn")
    k = int(g["k"])
    idx, dist = ops.knn_topk(dev(g["z"]), dev(g["bank"]), k)
    assert np.array_equal(idx.cpu().numpy(), g["topk_idx"])
    close(dist, g["topk_val"], rtol=3e-7)
    uniq, count = ops.unique_positions(idx, g["bank"].shape[0])
    n = int(count.item())
    assert np.array_equal(uniq[:n].cpu().numpy(), g["unique"])
    nn20, _ = ops.knn_topk(dev(g["z"]), dev(g["bank"]), 20, metric=1)
    assert np.array_equal(nn20.cpu().numpy(), g["nn20"])


def test_knn_topk_cfg3_size_vs_oracle(ops):
    rng = np.random.default_rng(0)
    B, N, D, k = 100, 25000, 40, 10
    bank = rng.normal(size=(N, D)).astype(np.float32)
    z = (bank[rng.integers(0, N, size=B)] + 0.2 * rng.normal(size=(B, D))).astype(np.float32)
    uniq_ref, idx_ref = O.nearest_exemplar_positions_np(z, bank, k)
    idx, _ = ops.knn_topk(dev(z), dev(bank), k)
    assert np.array_equal(idx.cpu().numpy(), idx_ref)
    uniq, count = ops.unique_positions(idx, N)
    assert np.array_equal(uniq[:int(count.item())].cpu().numpy(), uniq_ref)
    # duplicates (sampling with replacement): ties resolve to the lowest position, like the oracle
    bank2 = np.concatenate([bank[:500], bank[:500]], axis=0)
    _, idx_ref2 = O.nearest_exemplar_positions_np(z, bank2, k)
    idx2, _ = ops.knn_topk(dev(z), dev(bank2), k)
    assert np.array_equal(idx2.cpu().numpy(), idx_ref2)
    # shard + merge == global
    parts = [ops.knn_topk(dev(z), dev(bank[s:s + 6250]), k, pos_offset=s) for s in range(0, N, 6250)]
    mi, _ = ops.knn_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert np.array_equal(mi.cpu().numpy(), idx_ref)


@pytest.mark.parametrize("B,C,D,k,metric", [(7, 33, 5, 32, 0), (130, 1000, 128, 20, 1), (65, 10, 8, 16, 0),
                                            (300, 70000, 16, 5, 0), (64, 64, 40, 1, 0), (100, 100000, 128, 10, 0)])
def test_knn_fused_shapes_vs_oracle(ops, B, C, D, k, metric):
    """The fused distance + running top-k kernel at ragged / edge geometries: k = 32 (one list entry per lane), fewer
    candidates than k (tail = -1), several row blocks, many column splits, cfg5's N=100 000 x D=128 cache."""
    rng = np.random.default_rng(B + C + k)
    bank = rng.normal(size=(C, D)).astype(np.float32)
    z = (bank[rng.integers(0, C, size=B)] + 0.3 * rng.normal(size=(B, D))).astype(np.float32)
    idx, dist = ops.knn_topk(dev(z), dev(bank), k, metric=metric)
    idx = idx.cpu().numpy()
    kk = min(k, C)
    if metric == 0:
        _, ref = O.nearest_exemplar_positions_np(z, bank, kk)
    else:
        ref = O.find_nearest_neighbors_np(z, bank, kk)
    assert np.array_equal(idx[:, :kk], ref)
    assert (idx[:, kk:] == -1).all()
    # the unique of a fixed-capacity result: valid head, tail repeats the first entry
    uniq, count = ops.unique_positions(torch.as_tensor(idx[:, :kk]).cuda(), C)
    n = int(count.item())
    u = uniq.cpu().numpy()
    assert np.array_equal(u[:n], np.unique(ref.reshape(-1))) and (u[n:] == u[0]).all()


def test_prior_lse_valid_count(ops):
    """kNN mode keeps a fixed-capacity bank and a device-side count: rows beyond the count must not contribute to
    log p(z) (incl. the normaliser, models/BaseModel.py:107-108) nor receive / produce gradient."""
    g = torch.Generator().manual_seed(3)
    for B, cap, n, D in ((100, 1000, 983, 40), (64, 640, 300, 128)):
        mu = torch.randn(cap, D, generator=g)
        lv = torch.full((D,), -2.0)
        src = torch.randint(0, n, (B,), generator=g)
        z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
        mu_idx = torch.randperm(50000, generator=g)[:cap]
        z_idx = mu_idx[src].clone()
        gout = torch.randn(B, generator=g)
        cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
        outs = []
        for full in (False, True):
            zc, lc = z.cuda().requires_grad_(True), lv.cuda().requires_grad_(True)
            mc = (mu.cuda() if full else mu[:n].cuda()).requires_grad_(True)
            mi = mu_idx.cuda() if full else mu_idx[:n].cuda()
            lp = ops.prior_lse(zc, mc, lc, z_idx.cuda(), mi, c_valid=cnt if full else None)
            lp.backward(gout.cuda())
            outs.append((lp.detach(), zc.grad, mc.grad, lc.grad))
        (lp0, dz0, dm0, dl0), (lp1, dz1, dm1, dl1) = outs
        close(lp1, lp0, rtol=1e-6)
        close(dz1, dz0, rtol=1e-5, atol=1e-6 * float(dz0.abs().max()))
        close(dm1[:n], dm0, rtol=1e-5, atol=1e-6 * float(dm0.abs().max()))
        assert float(dm1[n:].abs().max()) == 0.0
        close(dl1, dl0, rtol=1e-4, atol=1e-5 * float(dl0.abs().max()))


def test_gather_scatter_rows(ops):
    src = torch.randn(1000, 784, device="cuda")
    idx = torch.randint(0, 1000, (333,), device="cuda")
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    src2 = torch.randn(50, 7, device="cuda")
    idx2 = torch.randperm(50, device="cuda")[:20]
    assert torch.equal(ops.gather_rows(src2, idx2), src2[idx2])
    dst = torch.zeros(100, 40, device="cuda"); rows = torch.randn(20, 40, device="cuda")
    ii = torch.randperm(100, device="cuda")[:20]
    ops.scatter_rows_(dst, ii, rows)
    ref = torch.zeros(100, 40, device="cuda"); ref[ii] = rows
    assert torch.equal(dst, ref)


# ------------------------------------------------------------------ VampPrior
@pytest.mark.parametrize("B,C,D", [(100, 500, 40), (7, 33, 24), (512, 1000, 40)])
def test_vamp_lse_fwd_bwd_vs_oracle(ops, B, C, D):
    g = torch.Generator().manual_seed(B + C)
    z = torch.randn(B, D, generator=g)
    pm = torch.randn(C, D, generator=g)
    plv = torch.clamp(torch.randn(C, D, generator=g), -6, 2)
    gout = torch.randn(B, generator=g)
    ts = [t.double().requires_grad_(True) for t in (z, pm, plv)]
    mat = O.t_vamp_logprob_matrix(*ts)
    ref = torch.logsumexp(mat, 1)
    ref.backward(gout.double())
    cs = [t.cuda().requires_grad_(True) for t in (z, pm, plv)]
    close(ops.vamp_logprob_matrix(*cs), mat.detach(), rtol=1e-5, atol=1e-4)
    lp = ops.vamp_lse(*cs)
    close(lp, ref.detach(), rtol=1e-5, atol=1e-4)
    lp.backward(gout.cuda())
    for c, t in zip(cs, ts):
        scale = t.grad.abs().max().item()
        close(c.grad, t.grad, rtol=1e-4, atol=1e-5 * scale)


# ------------------------------------------------------------------ K3
@pytest.mark.parametrize("R,K,Oo", [(12, 784, 48), (300, 300, 300), (1000, 40, 300), (77, 96, 24), (257, 300, 784),
                                    (5, 7, 3)])
def test_gated_dense_fwd_bwd(ops, R, K, Oo):
    g = torch.Generator().manual_seed(R + K + Oo)
    x = torch.randn(R, K, generator=g)
    Wh = torch.randn(Oo, K, generator=g) / K ** 0.5; Wg = torch.randn(Oo, K, generator=g) / K ** 0.5
    bh = torch.randn(Oo, generator=g); bg = torch.randn(Oo, generator=g)
    dout = torch.randn(R, Oo, generator=g)
    ts = [t.double().requires_grad_(True) for t in (x, Wh, bh, Wg, bg)]
    ref = (ts[0] @ ts[1].t() + ts[2]) * torch.sigmoid(ts[0] @ ts[3].t() + ts[4])
    ref.backward(dout.double())
    cs = [t.cuda().requires_grad_(True) for t in (x, Wh, bh, Wg, bg)]
    out = ops.gated_dense(*cs)
    close(out, ref, rtol=2e-5, atol=3e-5)      # 3xTF32 tensor-core path: ~5e-6 of the output scale
    out.backward(dout.cuda())
    for c, t in zip(cs, ts):
        close(c.grad, t.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("act", ["none", "sigmoid", "hardtanh", "relu"])
@pytest.mark.parametrize("R,K,Oo", [(12, 48, 784), (300, 300, 40), (130, 294, 40), (9, 5, 6)])
def test_linear_fwd_bwd(ops, act, R, K, Oo):
    from exemplar_vae_b200._lib import ACT_HARDTANH, ACT_NONE, ACT_RELU, ACT_SIGMOID
    g = torch.Generator().manual_seed(R * 3 + K + Oo)
    x = torch.randn(R, K, generator=g); W = torch.randn(Oo, K, generator=g) / K ** 0.5 * 3
    b = torch.randn(Oo, generator=g); dout = torch.randn(R, Oo, generator=g)
    ts = [t.double().requires_grad_(True) for t in (x, W, b)]
    pre = ts[0] @ ts[1].t() + ts[2]
    fn = {"none": lambda v: v, "sigmoid": torch.sigmoid, "hardtanh": lambda v: torch.clamp(v, -0.5, 0.7),
          "relu": torch.relu}[act]
    code = {"none": ACT_NONE, "sigmoid": ACT_SIGMOID, "hardtanh": ACT_HARDTANH, "relu": ACT_RELU}[act]
    ref = fn(pre); ref.backward(dout.double())
    cs = [t.cuda().requires_grad_(True) for t in (x, W, b)]
    out = ops.linear(cs[0], cs[1], cs[2], code, -0.5, 0.7)
    close(out, ref, rtol=2e-5, atol=3e-5)
    out.backward(dout.cuda())
    for c, t in zip(cs, ts):
        close(c.grad, t.grad, rtol=1e-4, atol=1e-4)
    # no bias / no input grad variant
    out2 = ops.linear(x.cuda(), cs[1], None, code, -0.5, 0.7)
    close(out2, fn(ts[0] @ ts[1].t()), rtol=2e-5, atol=3e-5)


# ------------------------------------------------------------------ element-wise
def test_elementwise_golden_and_grads(ops, golden):
    g = golden("prior")
    x, m, lv = (dev(g[k]).requires_grad_(True) for k in ("e:x", "e:m", "e:lv"))
    out = ops.log_normal_diag(x, m, lv)
    close(out, g["e:log_normal_diag"], rtol=2e-6)
    w = torch.randn(out.shape[0], generator=torch.Generator().manual_seed(5)).cuda()   # seeded: the tolerances are tight
    (out * w).sum().backward()
    xc, mc, lc = (torch.tensor(g[k]).requires_grad_(True) for k in ("e:x", "e:m", "e:lv"))
    (O.t_log_normal_diag(xc, mc, lc) * w.cpu()).sum().backward()
    for a, b in ((x, xc), (m, mc), (lv, lc)):
        close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
    close(ops.log_normal_standard(dev(g["e:x"])), g["e:log_normal_standard"], rtol=2e-6)
    pm = dev(g["e:pm"]).requires_grad_(True)
    lb = ops.log_bernoulli(dev(g["e:xb"]), pm)
    close(lb, g["e:log_bernoulli"], rtol=2e-6)
    (lb * w).sum().backward()
    pc = torch.tensor(g["e:pm"]).requires_grad_(True)
    (O.t_log_bernoulli(torch.tensor(g["e:xb"]), pc) * w.cpu()).sum().backward()
    close(pm.grad, pc.grad, rtol=1e-5, atol=1e-6)
    xm, xlv = dev(g["e:xm"]).requires_grad_(True), dev(g["e:xlv"]).requires_grad_(True)
    ll = ops.log_logistic_256(dev(g["e:xc"]), xm, xlv)
    close(ll, g["e:log_logistic_256"], rtol=2e-5)
    (ll * w).sum().backward()
    xmc, xlc = torch.tensor(g["e:xm"]).requires_grad_(True), torch.tensor(g["e:xlv"]).requires_grad_(True)
    (O.t_log_logistic_256(torch.tensor(g["e:xc"]), xmc, xlc) * w.cpu()).sum().backward()
    close(xm.grad, xmc.grad, rtol=2e-3, atol=1e-4)
    close(xlv.grad, xlc.grad, rtol=2e-3, atol=1e-4)


def test_reparam_elbo_lincomb(ops):
    g = torch.Generator().manual_seed(0)
    mu, lv, eps = torch.randn(33, 40, generator=g), torch.randn(33, 40, generator=g), torch.randn(33, 40, generator=g)
    mc, lc = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    zc = mc + torch.exp(0.5 * lc) * eps
    w = torch.randn(33, 40, generator=g)
    (zc * w).sum().backward()
    mg, lg = mu.cuda().requires_grad_(True), lv.cuda().requires_grad_(True)
    z = ops.reparameterize(mg, lg, eps.cuda())
    close(z, zc, rtol=1e-6, atol=1e-6)
    (z * w.cuda()).sum().backward()
    close(mg.grad, mc.grad, rtol=1e-6); close(lg.grad, lc.grad, rtol=1e-5, atol=1e-6)
    RE, KL = torch.randn(50, generator=g) * 100, torch.randn(50, generator=g) * 10
    rg, kg = RE.cuda().requires_grad_(True), KL.cuda().requires_grad_(True)
    o3 = ops.elbo_reduce(rg, kg, 0.37, True)
    close(o3, torch.stack(((-RE + 0.37 * KL).mean(), RE.mean(), KL.mean())), rtol=1e-5)
    o3[0].backward()
    close(rg.grad, torch.full((50,), -1 / 50), rtol=1e-6); close(kg.grad, torch.full((50,), 0.37 / 50), rtol=1e-6)
    lb = ops.elbo_reduce(RE.cuda(), KL.cuda(), 0.37, False)
    close(lb, -RE + 0.37 * KL, rtol=1e-6)
    a, b = torch.randn(9, generator=g), torch.randn(9, generator=g)
    ag = a.cuda().requires_grad_(True)
    lc2 = ops.lincomb((-1.0, 1.0), ag, b.cuda())
    close(lc2, b - a, rtol=1e-6, atol=1e-7)
    lc2.sum().backward()
    close(ag.grad, -torch.ones(9), rtol=0)


def test_rng_statistics(ops):
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    n = 1 << 20
    e = ops.rng_normal((n,), 1234, c, 3, "cuda")
    assert abs(e.mean().item()) < 5e-3 and abs(e.std().item() - 1) < 5e-3
    assert abs((e ** 4).mean().item() - 3.0) < 0.05
    p = torch.rand(n, device="cuda")
    b = ops.rng_bernoulli(p, 1234, c, 1)
    assert set(b.unique().tolist()) <= {0.0, 1.0}
    assert abs((b - p).mean().item()) < 2e-3
    r = ops.rng_randint(0, 50000, n, 1234, c, 2, "cuda")
    assert r.min().item() >= 0 and r.max().item() < 50000
    assert abs(r.float().mean().item() - 24999.5) < 100
    e1 = ops.rng_normal((64,), 1234, c, 3, "cuda")
    ops.rng_advance_(c, 1)
    e2 = ops.rng_normal((64,), 1234, c, 3, "cuda")
    assert torch.equal(e1, e[:64]) and not torch.equal(e1, e2)
    # in-kernel advance: the drawing kernel bumps counter[0] itself (ticket word counter[1] returns to 0)
    c2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    a0 = ops.rng_normal((n,), 1234, c2, 3, "cuda", advance=True)
    a1 = ops.rng_normal((64,), 1234, c2, 3, "cuda", advance=True)
    b1 = ops.rng_bernoulli(p, 1234, c2, 1, advance=True)
    torch.cuda.synchronize()
    assert c2.tolist() == [3, 0]
    assert torch.equal(a0, e) and torch.equal(a1, e2)            # offsets 0 and 1 of the same stream
    c.fill_(2)
    assert torch.equal(b1, ops.rng_bernoulli(p, 1234, c, 1))


def test_adam_normgrad_vs_oracle():
    from exemplar_vae_b200.optimizer import AdamNormGrad
    g = torch.Generator().manual_seed(1)
    shapes = [(300, 784), (300,), (1,), (40, 300), (7, 3)]
    ps = [torch.randn(*s, generator=g) for s in shapes]
    cpu = {str(i): p.clone().requires_grad_(True) for i, p in enumerate(ps)}
    gpu = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = AdamNormGrad(gpu, lr=5e-4)
    st = {}
    for it in range(3):
        grads = [torch.randn(*s, generator=g) * (10.0 ** (it - 1)) for s in shapes]
        for i, gr in enumerate(grads):
            cpu[str(i)].grad = gr.clone()
            if gpu[i].grad is None:
                gpu[i].grad = gr.clone().cuda()
            else:
                gpu[i].grad.copy_(gr)
        O.adam_normgrad_step(cpu, st, lr=5e-4)
        opt.step()
    for i in range(len(ps)):
        close(gpu[i], cpu[str(i)], rtol=1e-5, atol=1e-7)
    sd = opt.state_dict()
    assert sd["state"][0]["step"] == 3
    close(sd["state"][0]["exp_avg"], st["0"]["exp_avg"], rtol=1e-5, atol=1e-8)


def test_cpu_tensor_is_rejected(ops):
    from exemplar_vae_b200 import ExvaeError
    with pytest.raises(ExvaeError):
        ops.pairwise_distance(torch.randn(3, 4), torch.randn(5, 4))


# ------------------------------------------------------------------ K4 (implicit-GEMM conv; patch matrix for C < 16)
@pytest.mark.parametrize("N,C,H,Cout,k,s,p", [(3, 1, 28, 32, 7, 1, 3), (2, 32, 14, 64, 5, 1, 2), (2, 32, 28, 32, 3, 2, 1),
                                              (2, 64, 7, 6, 3, 1, 1), (2, 3, 8, 48, 3, 2, 1), (2, 64, 28, 1, 1, 1, 0),
                                              (5, 64, 28, 64, 3, 1, 1),     # convhvae decoder layer: implicit fwd + dx
                                              (3, 64, 14, 64, 3, 2, 1),     # stride 2: implicit fwd, col2im dx
                                              (3, 48, 16, 48, 3, 1, 1),     # single_conv block: 48 channels (2 k-blocks/tap)
                                              (3, 96, 8, 96, 3, 1, 1), (2, 48, 16, 96, 3, 2, 1),
                                              (7, 32, 7, 64, 3, 1, 1),      # 7x7 maps: two images per 128-row tile, odd N
                                              (2, 96, 4, 2, 3, 1, 1),       # bottleneck head: 2 output channels
                                              (1, 48, 32, 3, 3, 1, 1), (2, 2, 4, 96, 3, 1, 1)])
def test_gated_conv2d_fwd_bwd(ops, N, C, H, Cout, k, s, p):
    F = torch.nn.functional
    g = torch.Generator().manual_seed(N + C + H + Cout + k)
    x = torch.randn(N, C, H, H, generator=g)
    Wh = torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5
    Wg = torch.randn(Cout, C, k, k, generator=g) / (C * k * k) ** 0.5
    bh, bg = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    ts = [t.double().requires_grad_(True) for t in (x, Wh, bh, Wg, bg)]
    ref = F.conv2d(ts[0], ts[1], ts[2], stride=s, padding=p) * torch.sigmoid(F.conv2d(ts[0], ts[3], ts[4], stride=s, padding=p))
    dout = torch.randn(ref.shape, generator=g)
    ref.backward(dout.double())
    xc = x.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)          # NHWC
    cs = [t.cuda().requires_grad_(True) for t in (Wh, bh, Wg, bg)]
    out = ops.conv2d_gated(xc, cs[0], cs[1], cs[2], cs[3], s, p)
    close(out.permute(0, 3, 1, 2), ref, rtol=2e-5, atol=3e-5)
    out.backward(dout.permute(0, 2, 3, 1).contiguous().cuda())
    close(xc.grad.permute(0, 3, 1, 2), ts[0].grad, rtol=1e-4, atol=1e-4)
    for c, t in zip(cs, ts[1:]):
        close(c.grad, t.grad, rtol=1e-4, atol=2e-4)
    # plain conv + fused sigmoid, forward and backward
    for t in ts:
        t.grad = None
    ref2 = torch.sigmoid(F.conv2d(ts[0], ts[1], ts[2], stride=s, padding=p))
    ref2.backward(dout.double())
    xc2 = xc.detach().clone().requires_grad_(True)
    c2 = [cs[0].detach().clone().requires_grad_(True), cs[1].detach().clone().requires_grad_(True)]
    out2 = ops.conv2d(xc2, c2[0], c2[1], s, p, 1)
    close(out2.permute(0, 3, 1, 2), ref2, rtol=2e-5, atol=3e-5)
    out2.backward(dout.permute(0, 2, 3, 1).contiguous().cuda())
    close(xc2.grad.permute(0, 3, 1, 2), ts[0].grad, rtol=1e-4, atol=1e-4)
    close(c2[0].grad, ts[1].grad, rtol=1e-4, atol=2e-4)
    close(c2[1].grad, ts[2].grad, rtol=1e-4, atol=2e-4)


def test_weight_norm_fwd_bwd(ops):
    """torch.nn.utils.weight_norm(nn.Conv2d) (models/fully_conv.py:17): w = g v / ||v|| over each output channel."""
    g_ = torch.Generator().manual_seed(4)
    v = torch.randn(48, 16, 3, 3, generator=g_)
    gg = torch.rand(48, 1, 1, 1, generator=g_) + 0.5
    dw = torch.randn(48, 16, 3, 3, generator=g_)
    vd, gd = v.double().requires_grad_(True), gg.double().requires_grad_(True)
    w_ref = vd * (gd / vd.flatten(1).norm(dim=1).view(-1, 1, 1, 1))
    w_ref.backward(dw.double())
    vc, gc = v.cuda().requires_grad_(True), gg.cuda().requires_grad_(True)
    w = ops.weight_norm(vc, gc)
    close(w, w_ref, rtol=1e-5, atol=1e-6)
    w.backward(dw.cuda())
    close(vc.grad, vd.grad, rtol=1e-4, atol=1e-5)
    close(gc.grad, gd.grad, rtol=1e-4, atol=1e-5)


def test_elu_upsample(ops):
    F = torch.nn.functional
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 5, 6, 4, generator=g)
    xd = x.clone().requires_grad_(True)
    xc = x.cuda().requires_grad_(True)
    w = torch.randn(3, 5, 6, 4, generator=g)
    (F.elu(xd) * w).sum().backward()
    y = ops.elu(xc); (y * w.cuda()).sum().backward()
    close(y, F.elu(x), rtol=1e-6, atol=1e-7); close(xc.grad, xd.grad, rtol=1e-6, atol=1e-7)
    xd2 = x.permute(0, 3, 1, 2).clone().requires_grad_(True)                       # NCHW for torch
    up = F.interpolate(xd2, scale_factor=2)
    w2 = torch.randn(up.shape, generator=g)
    (up * w2).sum().backward()
    xc2 = x.cuda().requires_grad_(True)
    u = ops.upsample2x(xc2)
    close(u.permute(0, 3, 1, 2), up, rtol=0)
    (u * w2.permute(0, 2, 3, 1).contiguous().cuda()).sum().backward()
    close(xc2.grad.permute(0, 3, 1, 2), xd2.grad, rtol=1e-6, atol=1e-6)
