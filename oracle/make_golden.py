"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on CPU with its RNG draws replaced by recorded tensors.

Run in the build container only (the reference does not exist on the GPU box):

    python oracle/make_golden.py

The fixtures are committed; tests never import the reference.  RNG call sites patched
(SURVEY.md §4): models/BaseModel.py:81 (eps), :245/:257 (exemplar indices).  The Bernoulli
draw of utils/training.py:31 is replaced by passing an already-binarised batch.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
from argparse import Namespace

import numpy as np
import torch

REF = os.environ.get("EXVAE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def ref_args(**kw):
    d = dict(model_name="vae", prior="exemplar_prior", input_type="binary", input_size=[1, 28, 28],
             hidden_size=48, z1_size=40, z2_size=40, number_components=64, training_set_size=128,
             approximate_prior=False, approximate_k=10, no_mask=False, no_attention=False,
             same_variational_var=False, use_logit=False, lambd=1e-4, bottleneck=6,
             dataset_name="dynamic_mnist", device="cpu", dynamic_binarization=True, warmup=100,
             batch_size=12, lr=5e-4, continuous=False)
    d.update(kw)
    return Namespace(**d)


class RngTape:
    """Replays recorded eps / randint draws inside the reference."""

    def __init__(self, eps_list, randint_list):
        self.eps = list(eps_list)
        self.ri = list(randint_list)

    def __enter__(self):
        from models import BaseModel as BM
        self._BM = BM
        self._orig_rep = BM.BaseModel.reparameterize
        self._orig_ri = torch.randint
        tape = self

        def reparameterize(self_m, mu, logvar):          # models/BaseModel.py:79-82 with eps injected
            std = logvar.mul(0.5).exp_()
            eps = tape.eps.pop(0)
            assert eps.shape == std.shape
            return eps.mul(std).add_(mu)

        def randint(*a, **k):
            return tape.ri.pop(0).clone()

        BM.BaseModel.reparameterize = reparameterize
        torch.randint = randint
        return self

    def __exit__(self, *exc):
        self._BM.BaseModel.reparameterize = self._orig_rep
        torch.randint = self._orig_ri


def build_ref_model(args, seed):
    from utils.utils import importing_model
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = importing_model(args)(args)
    return model


def sd_np(model, prefix="p:"):
    return {prefix + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def golden_prior(rng):
    """a1-a4: pairwise_distance, log_normal_diag_vectorized, log_p_z_exemplar, log_p_z."""
    from utils.distributions import pairwise_distance, log_normal_diag_vectorized, log_normal_diag, \
        log_bernoulli, log_normal_standard, log_logistic_256
    out = {}
    for tag, (B, C, D, T) in {"s": (16, 200, 40, 150), "r": (37, 333, 40, 5000), "d24": (9, 130, 24, 100)}.items():
        args = ref_args(z1_size=D, z2_size=D, number_components=C, training_set_size=T)
        model = build_ref_model(args, 1)
        g = torch.Generator().manual_seed(100 + B)
        mu = torch.randn(C, D, generator=g)
        lv = torch.tensor([-2.4189], dtype=torch.float32) if tag != "d24" else torch.tensor([-0.7])
        src = torch.randperm(C, generator=g)[:B]
        z = mu[src] + torch.exp(0.5 * lv) * torch.randn(B, D, generator=g)
        mu_idx = torch.randint(0, T, (C,), generator=g)
        z_idx = mu_idx[src].clone().view(-1, 1)          # every row hits >= 1 exemplar (leave-one-out fires)
        z_idx[0, 0] = T + 5                               # ... except row 0 (no hit)
        logvar_bank = lv * torch.ones(C, D)
        emb = (mu, logvar_bank, mu_idx)
        model.train()
        with torch.no_grad():
            pd = pairwise_distance(z, mu)
            ln, pd2 = log_normal_diag_vectorized(z, mu, logvar_bank[0:1])
            prob_train = model.log_p_z((z, z_idx), emb, sum=False, test=False)
            lse_train = model.log_p_z((z, z_idx), emb, sum=True, test=False)
            lse_test = model.log_p_z((z, None), emb, sum=True, test=True)
        # gradients of sum(w * log_p) wrt z, mu, prior_log_variance-like scalar
        zg = z.clone().requires_grad_(True)
        mug = mu.clone().requires_grad_(True)
        lvg = lv.clone().requires_grad_(True)
        w = torch.randn(B, generator=g)
        lp = model.log_p_z((zg, z_idx), (mug, lvg * torch.ones(C, D), mu_idx), sum=True, test=False)
        (lp * w).sum().backward()
        out.update({f"{tag}:z": z, f"{tag}:mu": mu, f"{tag}:lv": lv, f"{tag}:z_idx": z_idx, f"{tag}:mu_idx": mu_idx,
                    f"{tag}:pairwise_distance": pd, f"{tag}:log_normal": ln, f"{tag}:pair_dist_scaled": pd2,
                    f"{tag}:prob_train": prob_train, f"{tag}:lse_train": lse_train, f"{tag}:lse_test": lse_test,
                    f"{tag}:w": w, f"{tag}:dz": zg.grad, f"{tag}:dmu": mug.grad, f"{tag}:dlv": lvg.grad})
    # a17 elementwise log-densities
    g = torch.Generator().manual_seed(7)
    x = torch.randn(11, 40, generator=g); m = torch.randn(11, 40, generator=g)
    lvv = torch.rand(11, 40, generator=g) * 8 - 6
    xb = torch.bernoulli(torch.rand(11, 784, generator=g), generator=g)
    pm = torch.rand(11, 784, generator=g); pm[0, :5] = 0.0; pm[1, :5] = 1.0
    xc = torch.rand(11, 784, generator=g); xm = torch.rand(11, 784, generator=g)
    xlv = torch.rand(11, 784, generator=g) * 4.5 - 4.5
    out.update({"e:x": x, "e:m": m, "e:lv": lvv, "e:log_normal_diag": log_normal_diag(x, m, lvv, dim=1),
                "e:log_normal_standard": log_normal_standard(x, dim=1),
                "e:xb": xb, "e:pm": pm, "e:log_bernoulli": log_bernoulli(xb, pm, dim=1),
                "e:xc": xc, "e:xm": xm, "e:xlv": xlv, "e:log_logistic_256": log_logistic_256(xc, xm, xlv, dim=1)})
    return {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}


def golden_knn():
    """a6/a20: top-k positions (tie-free by construction) + unique, and find_nearest_neighbors."""
    from utils.distributions import pairwise_distance
    from utils.knn_on_latent import find_nearest_neighbors
    g = torch.Generator().manual_seed(11)
    B, N, D, k = 24, 700, 40, 10
    bank = torch.randn(N, D, generator=g)
    z = bank[torch.randperm(N, generator=g)[:B]] + 0.3 * torch.randn(B, D, generator=g)
    d = pairwise_distance(z, bank)
    vals, idx = d.topk(k=k, largest=False, dim=1)
    srt = torch.sort(d, dim=1).values
    gap = (srt[:, 1:k + 1] - srt[:, :k]).min().item()
    assert gap > 0, "golden must be tie-free"
    uniq = torch.unique(idx.view(-1))
    nn20 = find_nearest_neighbors(z, bank, None)
    dd = ((z.unsqueeze(1) - bank.unsqueeze(0)) ** 2).sum(2) ** 0.5
    s2 = torch.sort(dd, dim=1).values
    assert (s2[:, 1:21] - s2[:, :20]).min().item() > 0
    return {"z": z.numpy(), "bank": bank.numpy(), "k": np.int64(k), "dist": d.numpy(), "topk_val": vals.numpy(),
            "topk_idx": idx.numpy(), "unique": uniq.numpy(), "min_gap": np.float32(gap), "nn20": nn20.numpy()}


def _step(model, args, x, x_idx, eps_list, ex_idx, dataset, beta, cache=None):
    from utils.optimizer import AdamNormGrad
    import warnings
    model.train()
    opt = AdamNormGrad(model.parameters(), lr=args.lr)
    with RngTape(eps_list, [ex_idx]):
        opt.zero_grad()
        loss, RE, KL = model.calculate_loss((x, x_idx), beta, average=True, cache=cache, dataset=dataset)
        loss.backward()
    grads = {"g:" + n: p.grad.detach().numpy().copy() for n, p in model.named_parameters() if p.grad is not None}
    with RngTape([e.clone() for e in eps_list], [ex_idx]):
        with torch.no_grad():
            l2, re2, kl2 = model.calculate_loss((x, x_idx), beta, average=False, cache=cache, dataset=dataset)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.step()
    new = {"n:" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    return loss, RE, KL, grads, new, (l2, re2, kl2)


def golden_model_step(model_name, hidden, seed, side=28, T=128, N=64, B=12, chans=1, D=40, **extra):
    """a5, a8-a19 (+ AdamNormGrad): one training step of the exact exemplar prior."""
    P = chans * side * side
    args = ref_args(model_name=model_name, hidden_size=hidden, number_components=N, training_set_size=T, batch_size=B,
                    input_size=[chans, side, side], z1_size=D, z2_size=D, **extra)
    model = build_ref_model(args, seed)
    with torch.no_grad():
        if args.prior == "exemplar_prior":
            model.prior_log_variance.fill_(-1.3)
    compact = extra.pop("_compact", False) if False else model_name in ("convhvae_2level", "single_conv")
    if compact:     # big models: parameters are regenerated from seeds by the tests (oracle.synth_params)
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
        from oracle.exvae_oracle import synth_params
        sp = synth_params({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed)
        model.load_state_dict(sp, strict=False)
    g = torch.Generator().manual_seed(seed + 1)
    data = torch.rand(T, P, generator=g)
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    bidx = torch.randperm(T, generator=g)[:B]
    x = torch.bernoulli(data[bidx], generator=g) if args.input_type == "binary" else data[bidx].clone()
    x_idx = bidx.view(-1, 1)
    ex_idx = torch.randint(0, T, (N,), generator=g)
    ex_idx[:3] = bidx[:3]                                  # guarantee leave-one-out hits
    n_eps = 1 if model_name in ("vae", "single_conv") else 2
    eps_list = [torch.randn(B, D, generator=g) for _ in range(n_eps)]
    beta = 0.37
    before = sd_np(model)
    import warnings
    warnings.simplefilter("ignore")
    loss, RE, KL, grads, new, per = _step(model, args, x, x_idx, [e.clone() for e in eps_list], ex_idx, dataset, beta)
    if compact:   # keep norms + the first 48 entries of every gradient / updated tensor
        head = lambda a: np.asarray(a).reshape(-1)[:48].copy()
        out = {"shape:" + k[2:]: np.asarray(v.shape, dtype=np.int64) for k, v in before.items()}
        out.update({"gh:" + k[2:]: head(v) for k, v in grads.items()})
        out.update({"gn:" + k[2:]: np.float64(np.linalg.norm(np.asarray(v, dtype=np.float64))) for k, v in grads.items()})
        from oracle.exvae_oracle import grad_projections          # 8 seeded +-1 projections of the WHOLE gradient
        out.update({"gp:" + k[2:]: grad_projections(v, k[2:]) for k, v in grads.items()})
        out.update({"nh:" + k[2:]: head(v) for k, v in new.items() if v.dtype == np.float32})
        out["seed"] = np.int64(seed)
    else:
        out = dict(before)
        out.update(grads); out.update(new)
    out.update({"x": x.numpy(), "x_idx": x_idx.numpy(), "ex_idx": ex_idx.numpy(), "exemplars": data[ex_idx].numpy(),
                "beta": np.float32(beta), "lr": np.float32(args.lr), "hidden": np.int64(hidden),
                "loss": loss.detach().numpy(), "RE": RE.detach().numpy(), "KL": KL.detach().numpy(),
                "loss_b": per[0].numpy(), "RE_b": per[1].numpy(), "KL_b": per[2].numpy(),
                "T": np.int64(T), "side": np.int64(side), "chans": np.int64(chans), "D": np.int64(D)})
    for i, e in enumerate(eps_list):
        out[f"eps{i}"] = e.numpy()
    return out


def golden_approx():
    """a6/a7: cache_z + get_approximate_nearest_exemplars + loss in kNN mode (vae, k=3)."""
    T, N, B, k = 160, 96, 10, 3
    side = 14
    args = ref_args(model_name="vae", hidden_size=32, number_components=N, training_set_size=T, batch_size=B,
                    approximate_prior=True, approximate_k=k, input_size=[1, side, side])
    model = build_ref_model(args, 5)
    with torch.no_grad():
        model.prior_log_variance.fill_(-0.9)
    g = torch.Generator().manual_seed(55)
    data = torch.rand(T, side * side, generator=g)
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    bidx = torch.randperm(T, generator=g)[:B]
    x = torch.bernoulli(data[bidx], generator=g)
    x_idx = bidx.view(-1, 1)
    ex_idx = torch.randint(0, T, (N,), generator=g)
    eps = torch.randn(B, 40, generator=g)
    model.train()
    with torch.no_grad():
        cache = model.cache_z(dataset)
    cache0 = (cache[0].clone(), cache[1].clone())
    before = sd_np(model)
    with RngTape([eps.clone()], [ex_idx]):
        z_mean, z_logvar = model.q_z(x)
        sel = model.get_approximate_nearest_exemplars((z_mean, z_logvar, x_idx), (cache[0].clone(), cache[1].clone()), dataset)
    cache_run = (cache[0].clone(), cache[1].clone())
    with RngTape([eps.clone()], [ex_idx]):
        loss, RE, KL = model.calculate_loss((x, x_idx), 1.0, average=True, cache=cache_run, dataset=dataset)
    out = dict(before)
    out.update({"data": data.numpy(), "x": x.numpy(), "x_idx": x_idx.numpy(), "ex_idx": ex_idx.numpy(), "eps": eps.numpy(),
                "k": np.int64(k), "cache_mean": cache0[0].numpy(), "cache_logvar": cache0[1].numpy(),
                "sel_indices": sel[2].numpy(), "sel_mean": sel[0].detach().numpy(),
                "cache_after": cache_run[0].detach().numpy(),
                "loss": loss.detach().numpy(), "RE": RE.detach().numpy(), "KL": KL.detach().numpy(), "hidden": np.int64(32), "side": np.int64(side)})
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}; goldens are generated in the build container only")
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    np.savez_compressed(os.path.join(OUT, "prior.npz"), **golden_prior(None))
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **golden_knn())
    np.savez_compressed(os.path.join(OUT, "vae_step.npz"), **golden_model_step("vae", 48, 3))
    np.savez_compressed(os.path.join(OUT, "hvae_step.npz"), **golden_model_step("hvae_2level", 24, 4, side=14))
    np.savez_compressed(os.path.join(OUT, "approx.npz"), **golden_approx())
    # f4: the same step under the VampPrior (20 pseudo-inputs spread over [0,1] so that Hardtanh(0,1) clips some)
    np.savez_compressed(os.path.join(OUT, "vamp_step.npz"),
                        **golden_model_step("vae", 32, 9, side=14, T=64, N=20, B=10, prior="vampprior",
                                            use_training_data_init=False, pseudoinputs_mean=0.4,
                                            pseudoinputs_std=0.35))
    np.savez_compressed(os.path.join(OUT, "convhvae_step.npz"),
                        **golden_model_step("convhvae_2level", 300, 6, side=28, T=24, N=6, B=4))
    np.savez_compressed(os.path.join(OUT, "single_conv_step.npz"),
                        **golden_model_step("single_conv", 300, 8, side=8, T=24, N=6, B=4, chans=3, D=8,
                                            input_type="continuous", bottleneck=2))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
