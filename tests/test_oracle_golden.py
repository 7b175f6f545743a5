"""The oracle (oracle/exvae_oracle.py) against fixtures produced by the reference itself
(oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import exvae_oracle as O

TAGS = ("s", "r", "d24")


def _close(a, b, rtol, atol=0.0):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def test_pairwise_distance_bit_exact(golden):
    g = golden("prior")
    for t in TAGS:
        d = O.pairwise_distance_np(g[f"{t}:z"], g[f"{t}:mu"])
        # fp64 accumulate then round to fp32: identical up to (rare) rounding-boundary flips
        assert np.mean(d != g[f"{t}:pairwise_distance"]) < 1e-3
        _close(d, g[f"{t}:pairwise_distance"], rtol=3e-7)


def test_log_normal_diag_vectorized(golden):
    g = golden("prior")
    for t in TAGS:
        C = g[f"{t}:mu"].shape[0]
        lv = (g[f"{t}:lv"] * np.ones((1, g[f"{t}:mu"].shape[1]), dtype=np.float32)).astype(np.float32)
        ln, pd = O.log_normal_diag_vectorized_np(g[f"{t}:z"], g[f"{t}:mu"], lv)
        _close(pd, g[f"{t}:pair_dist_scaled"], rtol=2e-6, atol=1e-4)
        _close(ln, g[f"{t}:log_normal"], rtol=2e-6, atol=1e-4)


def test_log_p_z_exemplar_matrix_and_lse(golden):
    g = golden("prior")
    for t in TAGS:
        C, D = g[f"{t}:mu"].shape
        bank_lv = (g[f"{t}:lv"] * np.ones((C, D), dtype=np.float32)).astype(np.float32)
        prob = O.log_p_z_exemplar_np(g[f"{t}:z"], g[f"{t}:z_idx"], g[f"{t}:mu"], bank_lv, g[f"{t}:mu_idx"], test=False)
        ref = g[f"{t}:prob_train"]
        assert np.array_equal(np.isneginf(prob), np.isneginf(ref))
        fin = np.isfinite(ref)
        _close(prob[fin], ref[fin], rtol=2e-6, atol=1e-4)
        lse = O.lse_rows_np(prob)
        _close(lse, g[f"{t}:lse_train"], rtol=1e-5)
        lse_t = O.log_p_z_exemplar_lse_np(g[f"{t}:z"], None, g[f"{t}:mu"], bank_lv, g[f"{t}:mu_idx"], test=True)
        _close(lse_t, g[f"{t}:lse_test"], rtol=1e-5)
        # both sit within 1e-5 of the exact fp64 value
        f64 = O.log_p_z_exemplar_lse_f64(g[f"{t}:z"], g[f"{t}:z_idx"], g[f"{t}:mu"], bank_lv[0], g[f"{t}:mu_idx"])
        _close(g[f"{t}:lse_train"], f64, rtol=1e-5)


def test_prior_gradients_via_torch_restatement(golden):
    g = golden("prior")
    for t in TAGS:
        z = torch.tensor(g[f"{t}:z"], requires_grad=True)
        mu = torch.tensor(g[f"{t}:mu"], requires_grad=True)
        lv = torch.tensor(g[f"{t}:lv"], requires_grad=True)
        D = mu.shape[1]
        lp = O.t_log_p_z_exemplar(z, torch.tensor(g[f"{t}:z_idx"]), mu, lv * torch.ones(D),
                                  torch.tensor(g[f"{t}:mu_idx"]), masked=True)
        _close(lp.detach().numpy(), g[f"{t}:lse_train"], rtol=1e-5)
        (lp * torch.tensor(g[f"{t}:w"])).sum().backward()
        _close(z.grad.numpy(), g[f"{t}:dz"], rtol=1e-4, atol=1e-5)
        _close(mu.grad.numpy(), g[f"{t}:dmu"], rtol=1e-4, atol=1e-5)
        _close(lv.grad.numpy(), g[f"{t}:dlv"], rtol=1e-4, atol=1e-4)


def test_elementwise_log_densities(golden):
    g = golden("prior")
    _close(O.log_normal_diag_np(g["e:x"], g["e:m"], g["e:lv"]), g["e:log_normal_diag"], rtol=2e-6)
    _close(O.log_normal_standard_np(g["e:x"]), g["e:log_normal_standard"], rtol=2e-6)
    _close(O.log_bernoulli_np(g["e:xb"], g["e:pm"]), g["e:log_bernoulli"], rtol=2e-6)
    _close(O.log_logistic_256_np(g["e:xc"], g["e:xm"], g["e:xlv"]), g["e:log_logistic_256"], rtol=1e-5)


def test_knn_indices_bit_exact(golden):
    g = golden("knn")
    k = int(g["k"])
    assert float(g["min_gap"]) > 0
    uniq, idx = O.nearest_exemplar_positions_np(g["z"], g["bank"], k)
    assert np.array_equal(idx, g["topk_idx"])
    assert np.array_equal(uniq, g["unique"])
    assert np.array_equal(O.find_nearest_neighbors_np(g["z"], g["bank"], 20), g["nn20"])


def _params(g):
    return {k[2:]: torch.tensor(v).requires_grad_(True) for k, v in g.items() if k.startswith("p:")}


def _check_step(g, model_name, prior="exemplar_prior"):
    side = int(g["side"])
    args = O.make_args(model_name=model_name, hidden_size=int(g["hidden"]), number_components=len(g["ex_idx"]),
                       training_set_size=int(g["T"]), input_size=[1, side, side], prior=prior)
    p = _params(g)
    x = torch.tensor(g["x"]); xi = torch.tensor(g["x_idx"]); ex = torch.tensor(g["exemplars"])
    ei = torch.tensor(g["ex_idx"]); beta = float(g["beta"])
    eps = [torch.tensor(g[f"eps{i}"]) for i in range(1 if model_name == "vae" else 2)]
    fn = O.loss_fn(args)
    loss, RE, KL = fn(p, args, x, xi, *eps, ex, ei, beta=beta, average=True)
    _close(loss.item(), g["loss"], rtol=1e-5)
    _close(RE.item(), g["RE"], rtol=1e-5)
    _close(KL.item(), g["KL"], rtol=1e-5)
    loss.backward()
    for k, v in g.items():
        if k.startswith("g:"):
            _close(p[k[2:]].grad.numpy(), v, rtol=2e-3, atol=1e-6)
    O.adam_normgrad_step(p, {}, lr=float(g["lr"]))
    for k, v in g.items():
        if k.startswith("n:"):
            _close(p[k[2:]].detach().numpy(), v, rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        p0 = _params(g)
        lb, reb, klb = fn(p0, args, x, xi, *eps, ex, ei, beta=beta, average=False)
    _close(lb.numpy(), g["loss_b"], rtol=1e-5)
    _close(reb.numpy(), g["RE_b"], rtol=1e-5)
    _close(klb.numpy(), g["KL_b"], rtol=1e-5, atol=1e-4)


def test_vae_training_step(golden):
    _check_step(golden("vae_step"), "vae")


def test_hvae_training_step(golden):
    _check_step(golden("hvae_step"), "hvae_2level")


def test_vampprior_training_step(golden):
    """prior == 'vampprior' (per-component mean and log-variance), pinned by the reference."""
    _check_step(golden("vamp_step"), "vae", prior="vampprior")


def test_approximate_prior_selection(golden):
    """models/BaseModel.py:256-271 restated with the numpy primitives."""
    g = golden("approx")
    side = int(g["side"])
    args = O.make_args(model_name="vae", hidden_size=int(g["hidden"]), input_size=[1, side, side],
                       approximate_prior=True, approximate_k=int(g["k"]))
    p = _params(g)
    x = torch.tensor(g["x"])
    with torch.no_grad():
        z_mean, _ = O.vae_q_z(p, args, x)
    cache = g["cache_mean"].copy()
    cache[g["x_idx"].reshape(-1)] = z_mean.numpy()
    sub = cache[g["ex_idx"]]
    uniq, _ = O.nearest_exemplar_positions_np(z_mean.numpy(), sub, int(g["k"]))
    sel = g["ex_idx"][uniq]
    assert np.array_equal(sel, g["sel_indices"])
    with torch.no_grad():
        em, elv = O.vae_q_z(p, args, torch.tensor(g["data"][sel]), prior=True)
    _close(em.numpy(), g["sel_mean"], rtol=1e-5, atol=1e-6)
    loss, RE, KL = O.vae_loss(p, args, x, torch.tensor(g["x_idx"]), torch.tensor(g["eps"]), None, None, beta=1.0,
                              exemplars_embedding=(em, elv, torch.tensor(sel)))
    _close(loss.item(), g["loss"], rtol=1e-5)
    _close(KL.item(), g["KL"], rtol=1e-5)


def test_lse_partial_merge_is_associative():
    rng = np.random.default_rng(0)
    logits = rng.normal(size=(7, 90)).astype(np.float64) * 5
    full = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
    parts = np.split(logits, 3, axis=1)
    m = np.stack([p.max(1) for p in parts]); s = np.stack([np.exp(p - p.max(1, keepdims=True)).sum(1) for p in parts])
    M, S = O.merge_lse_partials_np(m, s)
    np.testing.assert_allclose(M + np.log(S), full, rtol=1e-12)


def _compact_check(g, model_name):
    side, chans, D = int(g["side"]), int(g["chans"]), int(g["D"])
    kw = dict(input_type="continuous", bottleneck=2) if model_name == "single_conv" else {}
    args = O.make_args(model_name=model_name, hidden_size=int(g["hidden"]), number_components=len(g["ex_idx"]),
                       training_set_size=int(g["T"]), input_size=[chans, side, side], z1_size=D, z2_size=D, **kw)
    shapes = {k[6:]: tuple(int(v) for v in g[k]) for k in g if k.startswith("shape:")}
    p = {k: v.requires_grad_(True) for k, v in O.synth_params(shapes, int(g["seed"])).items()}
    x = torch.tensor(g["x"]); xi = torch.tensor(g["x_idx"]); ex = torch.tensor(g["exemplars"])
    ei = torch.tensor(g["ex_idx"]); beta = float(g["beta"])
    eps = [torch.tensor(g[f"eps{i}"]) for i in range(1 if model_name == "single_conv" else 2)]
    loss, RE, KL = O.loss_fn(args)(p, args, x, xi, *eps, ex, ei, beta=beta, average=True)
    _close(loss.item(), g["loss"], rtol=2e-5)
    _close(RE.item(), g["RE"], rtol=2e-5)
    _close(KL.item(), g["KL"], rtol=2e-5)
    loss.backward()
    for k in g:
        if k.startswith("gn:"):
            name = k[3:]
            gr = p[name].grad.numpy()
            _close(np.linalg.norm(gr.astype(np.float64)), g[k], rtol=2e-3)
            _close(gr.reshape(-1)[:48], g["gh:" + name], rtol=5e-3, atol=2e-3 * (np.abs(g["gh:" + name]).max() + 1e-8))
            # every element of the tensor: seeded +-1 projections recorded from the reference's gradient
            proj = O.grad_projections(gr, name)
            assert np.abs(proj - g["gp:" + name]).max() <= 1e-3 * float(g[k]) + 1e-7, (name, proj, g["gp:" + name])


def test_convhvae_training_step(golden):
    _compact_check(golden("convhvae_step"), "convhvae_2level")


def test_single_conv_training_step(golden):
    _compact_check(golden("single_conv_step"), "single_conv")
