"""Parity of the path bench.py TIMES: ``GraphedTrainStep`` = CUDA graph replay + fused batch+exemplar
encoder pass + in-kernel gradient accumulation into the flat buffer + the prior term on a side stream
(parallel graph branch) + fused AdamNormGrad — compared, at the BASELINE.json sizes, with
  (a) the same steps run eagerly through the plain path (two encoder passes, autograd-accumulated
      gradients, no side stream), and
  (b) the CPU oracle's ``train_step`` (utils/training.py:27-46, models/BaseModel.py:65-77 restated),
with every random draw (binarised batch, exemplar indices, eps) injected into all three.

Tolerances: loss / RE / KL <= 1e-4 relative (north_star); parameters after the steps: Adam's first
updates are sign-like (m/sqrt(v) = +-1), so an element whose normalised gradient sits at the eps=1e-8
scale may legitimately differ by up to lr per step; all other elements must agree to 2e-5 absolute."""
import numpy as np
import pytest
import torch

from oracle import exvae_oracle as O

pytestmark = pytest.mark.gpu

LR = 5e-4


def close(a, b, rtol, atol=0.0):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol)


def _fresh(args, p):
    import exemplar_vae_b200 as E
    m = E.importing_model(args)(args).cuda()
    m.load_state_dict({k: v.detach().clone() for k, v in p.items()})
    m.train()
    return m


def _params_agree(got_sd, ref_sd, steps, what):
    for k, v in got_sd.items():
        a = v.detach().cpu().numpy().astype(np.float64)
        r = ref_sd[k]
        b = (r.detach().cpu().numpy() if torch.is_tensor(r) else np.asarray(r)).astype(np.float64)
        d = np.abs(a - b)
        assert d.max() <= steps * 1.1 * LR + 1e-6, (what, k, d.max())           # hard bound: sign-like steps of lr
        assert np.mean(d > 2e-5) < 2e-3, (what, k, float(np.mean(d > 2e-5)))    # all but near-zero-gradient elements


@pytest.mark.parametrize("model_name,B,N,T", [("vae", 512, 25000, 50000),            # BASELINE configs[1]
                                              ("hvae_2level", 256, 11500, 23000)])   # BASELINE configs[3] sizes
def test_graphed_step_matches_eager_and_oracle_at_baseline_size(model_name, B, N, T):
    import exemplar_vae_b200 as E
    steps, beta, D = 2, 0.8, 40
    args = O.make_args(model_name=model_name, hidden_size=300, number_components=N, training_set_size=T,
                       device="cuda")
    args.dynamic_binarization = False            # the binarised batch is injected (utils/training.py:31)
    p0 = O.init_params(args, seed=11)
    data = O.synthetic_dataset(T)
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    n_eps = 1 if model_name == "vae" else 2
    gen = torch.Generator().manual_seed(2024)
    draws = []
    for _ in range(steps):
        bidx = torch.randperm(T, generator=gen)[:B]
        x = torch.bernoulli(data[bidx], generator=gen)
        ex_idx = torch.randint(0, T, (N,), generator=gen)
        ex_idx[:9] = bidx[:9]                    # make the leave-one-out mask fire
        eps = [torch.randn(B, D, generator=gen) for _ in range(n_eps)]
        draws.append((bidx, x, ex_idx, eps))

    # ---- (b) the CPU oracle
    p_ref = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in p0.items()}
    st, ref_losses = {}, []
    for bidx, x, ex_idx, eps in draws:
        ro = dict(x=x, exemplar_indices=ex_idx)
        if model_name == "vae":
            ro["eps"] = eps[0]
        else:
            ro["eps2"], ro["eps1"] = eps[0], eps[1]
        ref_losses.append(O.train_step(p_ref, st, args, data[bidx], bidx.view(-1, 1), data, beta, gen, rng_override=ro))

    # ---- (a) eager, plain path
    m_e = _fresh(args, p0)
    m_e.fuse_exemplar_encoder = False
    opt_e = E.AdamNormGrad(m_e.parameters(), lr=LR)
    eager_losses = []
    for bidx, x, ex_idx, eps in draws:
        m_e.rng_override = {"eps": [e.cuda() for e in eps], "exemplar_indices": ex_idx.cuda()}
        opt_e.zero_grad()
        loss, RE, KL = m_e.calculate_loss((x.cuda(), bidx.view(-1, 1).cuda()), beta, average=True, dataset=dataset)
        loss.backward()
        opt_e.step()
        eager_losses.append((float(loss), float(RE), float(KL)))

    # ---- the benchmarked path: captured graph, fused encoder pass + gradients, prior on the side stream
    m_g = _fresh(args, p0)
    opt_g = E.AdamNormGrad(m_g.parameters(), lr=LR)
    static = {"eps": [torch.zeros(B, D, device="cuda") for _ in range(n_eps)],
              "exemplar_indices": torch.zeros(N, dtype=torch.int64, device="cuda")}
    step = E.GraphedTrainStep(m_g, opt_g, args, dataset, B, beta=0.123, warmup_steps=3, use_graph=True,
                              rng_override=static)
    assert step.graph is not None and m_g.overlap_prior and m_g.fuse_exemplar_encoder
    # constructing the step (eager warm-up on zero buffers + capture) must not have trained anything
    for k, v in m_g.state_dict().items():
        assert torch.equal(v.cpu(), p0[k].detach()), k
    assert int(opt_g._tables[0]["step"].item()) == 0
    step.set_beta(beta)                          # device scalar: no re-capture
    graph_losses = []
    # exemplar prefetch: a step gathers the NEXT step's exemplar rows behind its backward, so the injected index tensor
    # holds step k+1's draw while step k runs; prime_exemplars() loads the first set
    assert step.prefetch
    static["exemplar_indices"].copy_(draws[0][2])
    step.prime_exemplars()
    for k, (bidx, x, ex_idx, eps) in enumerate(draws):
        for dst, e in zip(static["eps"], eps):
            dst.copy_(e)
        if k + 1 < len(draws):
            static["exemplar_indices"].copy_(draws[k + 1][2])
        out = step.step(x.cuda(), bidx.cuda())
        graph_losses.append(tuple(out.tolist()))
    assert int(opt_g._tables[0]["step"].item()) == steps

    for it in range(steps):
        for got, eag, ref in zip(graph_losses[it], eager_losses[it], ref_losses[it]):
            close(got, eag, rtol=1e-4)
            close(got, ref, rtol=1e-4)
            close(eag, ref, rtol=1e-4)
    _params_agree(m_g.state_dict(), m_e.state_dict(), steps, "graph vs eager")
    _params_agree(m_g.state_dict(), p_ref, steps, "graph vs oracle")


def test_iwae_prior_shape_vs_fp64_oracle():
    """``calculate_likelihood`` runs the prior at B=5000 (one image x S samples) against the full
    N=50 000 bank, no mask (utils/evaluation.py:83-95).  CUDA at the full shape; the fp64 oracle on a row sample
    (rows are independent)."""
    from exemplar_vae_b200 import ops
    B, N, D = 5000, 50000, 40
    g = torch.Generator().manual_seed(8)
    mu = torch.randn(N, D, generator=g)
    src = torch.randint(0, N, (B,), generator=g)
    lv = torch.full((D,), -2.4189)
    z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
    lp = ops.prior_lse(z.cuda(), mu.cuda(), lv.cuda(), None, None).cpu().numpy()
    rows = np.arange(0, B, 37)
    ref = O.log_p_z_exemplar_lse_f64(z[rows].numpy(), None, mu.numpy(), lv.numpy(), None, masked=False)
    close(lp[rows], ref, rtol=1e-4)


def test_knn_mode_step_is_graph_capturable_and_matches_eager():
    """BASELINE configs[2] path (approximate_prior, kNN cache): the selection keeps a fixed capacity B*k and a device
    count (no host sync), so the whole step replays from a CUDA graph; replay == eager, incl. the refreshed cache."""
    import exemplar_vae_b200 as E
    T, N, B, k, D = 4000, 2000, 100, 10, 40
    args = O.make_args(model_name="vae", hidden_size=300, number_components=N, training_set_size=T, device="cuda",
                       approximate_prior=True, approximate_k=k)
    args.dynamic_binarization = False
    p0 = O.init_params(args, seed=5)
    data = O.synthetic_dataset(T)
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    gen = torch.Generator().manual_seed(77)
    draws = []
    for _ in range(3):
        bidx = torch.randperm(T, generator=gen)[:B]
        draws.append((bidx, torch.bernoulli(data[bidx], generator=gen), torch.randint(0, T, (N,), generator=gen),
                      torch.randn(B, D, generator=gen)))
    results = []
    for use_graph in (False, True):
        m = _fresh(args, p0)
        assert m.knn_graph_capturable
        opt = E.AdamNormGrad(m.parameters(), lr=LR)
        with torch.no_grad():
            cache = m.cache_z(dataset)
        cache = (cache[0].contiguous().clone(), cache[1].contiguous().clone())
        static = {"eps": [torch.zeros(B, D, device="cuda")], "exemplar_indices": torch.zeros(N, dtype=torch.int64, device="cuda")}
        step = E.GraphedTrainStep(m, opt, args, dataset, B, beta=1.0, warmup_steps=2, use_graph=use_graph,
                                  rng_override=static, cache=cache)
        assert (step.graph is not None) == use_graph
        losses = []
        for bidx, x, ex_idx, eps in draws:
            static["eps"][0].copy_(eps)
            static["exemplar_indices"].copy_(ex_idx)
            losses.append(step.step(x.cuda(), bidx.cuda()).clone())
        results.append((torch.stack(losses).cpu(), cache[0].cpu(), {k_: v.detach().cpu() for k_, v in m.state_dict().items()}))
    (l0, c0, s0), (l1, c1, s1) = results
    close(l1, l0, rtol=1e-5)
    close(c1, c0, rtol=1e-4, atol=1e-5)
    _params_agree(s1, s0, 3, "knn graph vs eager")
