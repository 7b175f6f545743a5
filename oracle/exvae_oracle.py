"""CPU oracle for the Exemplar-VAE hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (numpy for the index/selection arithmetic and the
exemplar-prior primitives, torch-CPU where the restatement needs autograd) of the
algorithms on the reference's training hot path.  It is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  The product path
(``exemplar_vae_b200``) never imports anything from here and raises when its CUDA
library is missing.

Parity pinning: the reference ships no tests and no golden vectors, so
this oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the build container and
writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those fixtures.

Every function cites the reference ``file:line`` it restates (paths relative to the
reference repository root).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)
MIN_EPS = 1e-5          # utils/distributions.py:6
MAX_EPS = 1.0 - 1e-5    # utils/distributions.py:7


# --------------------------------------------------------------------------------------
# numpy primitives (a1-a4, a17, a20)
# --------------------------------------------------------------------------------------
def pairwise_distance_np(z: np.ndarray, means: np.ndarray) -> np.ndarray:
    """utils/distributions.py:12-18 — ||z||^2 + ||mu||^2 - 2 z.mu, fp64 inside, fp32 out."""
    zd = z.astype(np.float64)
    md = means.astype(np.float64)
    d1 = (zd * zd).sum(axis=1)[:, None]
    d2 = (md * md).sum(axis=1)[None, :]
    d3 = zd @ md.T
    return ((d1 + d2) + (-2.0 * d3)).astype(np.float32)


def log_normal_diag_vectorized_np(x, mean, log_var):
    """utils/distributions.py:21-25 — log_var is [1, D]; returns (log_normal [B,C], pair_dist [B,C])."""
    x = x.astype(np.float32)
    mean = mean.astype(np.float32)
    log_var = log_var.astype(np.float32)
    sd = np.exp(np.float32(0.5) * log_var).astype(np.float32)
    pd = pairwise_distance_np((x / sd).astype(np.float32), (mean / sd).astype(np.float32))
    const = np.float32(-0.5) * np.sum(log_var + np.float32(LOG_2PI), axis=1, dtype=np.float32)
    log_normal = (const[:, None] - np.float32(0.5) * pd).astype(np.float32)
    return log_normal, pd


def log_p_z_exemplar_np(z, z_indices, centers, center_log_variance, center_indices, test, no_mask=False):
    """models/BaseModel.py:98-109 — per-pair log-density with leave-one-out mask, minus
    log(C - #masked).  ``center_log_variance`` is the [C, D] bank of which only row 0 is used."""
    C = centers.shape[0]
    B = z.shape[0]
    denom = np.full((B,), float(C), dtype=np.float32)
    lv = center_log_variance[0:1, :]
    prob, _ = log_normal_diag_vectorized_np(z, centers, lv)
    if (test is False) and (no_mask is False):
        mask = z_indices.reshape(-1, 1) == center_indices.reshape(1, -1)
        prob = np.where(mask, np.float32(-np.inf), prob)
        denom = denom - mask.sum(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        prob = prob - np.log(denom)[:, None]
    return prob.astype(np.float32)


def lse_rows_np(prob: np.ndarray) -> np.ndarray:
    """models/BaseModel.py:124-125 — max-shifted log-sum-exp over dim 1, in fp32."""
    prob = prob.astype(np.float32)
    pmax = prob.max(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (pmax + np.log(np.exp(prob - pmax[:, None]).sum(axis=1, dtype=np.float32))).astype(np.float32)


def log_p_z_exemplar_lse_np(z, z_indices, centers, center_log_variance, center_indices, test, no_mask=False):
    """models/BaseModel.py:111-128 with prior == 'exemplar_prior', sum=True."""
    return lse_rows_np(log_p_z_exemplar_np(z, z_indices, centers, center_log_variance,
                                           center_indices, test, no_mask))


def log_p_z_exemplar_lse_f64(z, z_indices, centers, log_var_row, center_indices, masked=True):
    """fp64 'truth' for the same quantity with direct differences (no expansion); used to
    show how far the fp32 paths (reference, oracle, CUDA) sit from the exact value."""
    zd = z.astype(np.float64)
    md = centers.astype(np.float64)
    lv = log_var_row.astype(np.float64).reshape(1, -1)
    out = np.empty((zd.shape[0],), dtype=np.float64)
    const = -0.5 * np.sum(lv + LOG_2PI)
    inv = np.exp(-lv)
    for b in range(zd.shape[0]):
        d = ((zd[b:b + 1] - md) ** 2 * inv).sum(axis=1)
        logit = const - 0.5 * d
        n = md.shape[0]
        if masked:
            m = center_indices.reshape(-1) == int(np.asarray(z_indices).reshape(-1)[b])
            logit = np.where(m, -np.inf, logit)
            n = n - int(m.sum())
        mx = logit.max()
        out[b] = mx + np.log(np.exp(logit - mx).sum()) - np.log(n)
    return out


def log_normal_diag_np(x, mean, log_var):
    """utils/distributions.py:28-33 with dim=1, average=False."""
    x, mean, log_var = (a.astype(np.float32) for a in (x, mean, log_var))
    t = np.float32(-0.5) * (log_var + np.float32(LOG_2PI) + (x - mean) ** 2 / np.exp(log_var))
    return t.sum(axis=1, dtype=np.float32)


def log_normal_standard_np(x):
    """utils/distributions.py:36-41 with dim=1."""
    x = x.astype(np.float32)
    t = np.float32(-0.5) * x * x - np.float32(0.5 * LOG_2PI)
    return t.sum(axis=1, dtype=np.float32)


def log_bernoulli_np(x, mean):
    """utils/distributions.py:44-51 with dim=1 (probabilities clamped to [1e-5, 1-1e-5])."""
    x = x.astype(np.float32)
    p = np.clip(mean.astype(np.float32), np.float32(MIN_EPS), np.float32(MAX_EPS))
    t = x * np.log(p) + (np.float32(1.0) - x) * np.log(np.float32(1.0) - p)
    return t.sum(axis=1, dtype=np.float32)


def log_logistic_256_np(x, mean, logvar):
    """utils/distributions.py:54-66 with dim=1."""
    x, mean, logvar = (a.astype(np.float32) for a in (x, mean, logvar))
    bin_size = np.float32(1.0 / 256.0)
    scale = np.exp(logvar)
    xs = (np.floor(x / bin_size) * bin_size - mean) / scale
    sig = lambda v: np.float32(1.0) / (np.float32(1.0) + np.exp(-v))
    cdf_plus = sig(xs + bin_size / scale)
    cdf_minus = sig(xs)
    return np.log(cdf_plus - cdf_minus + np.float32(1e-7)).sum(axis=1, dtype=np.float32)


def topk_smallest_np(dist: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """torch.topk(k, largest=False, dim=1, sorted=True) as used at models/BaseModel.py:263-264
    and utils/knn_on_latent.py:8.  Tie-break (unspecified in torch) is fixed here to the
    LOWEST position, which is the contract of the CUDA kernel; goldens are tie-free."""
    order = np.argsort(dist, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(dist, order, axis=1)
    return vals, order.astype(np.int64)


def nearest_exemplar_positions_np(z_mean, sub_cache, k):
    """models/BaseModel.py:263-265 — unscaled pairwise distance -> top-k smallest -> sorted unique positions."""
    d = pairwise_distance_np(z_mean, sub_cache)
    _, idx = topk_smallest_np(d, k)
    return np.unique(idx.reshape(-1)), idx


def find_nearest_neighbors_np(z_val, z_train, k=20):
    """utils/knn_on_latent.py:4-9 — direct-difference fp32 Euclidean distance, sqrt, top-k smallest sorted."""
    z_val = z_val.astype(np.float32)
    z_train = z_train.astype(np.float32)
    out = np.empty((z_val.shape[0], k), dtype=np.int64)
    for b in range(z_val.shape[0]):
        diff = z_val[b:b + 1, :] - z_train
        d = np.sqrt((diff * diff).sum(axis=1, dtype=np.float32))
        out[b] = np.argsort(d, kind="stable")[:k]
    return out


def merge_lse_partials_np(m: np.ndarray, s: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Associative merge of per-shard (max, sum-of-exp) partials.
    m, s: [G, B].  Returns (M [B], S [B]) with  lse = M + log S."""
    M = m.max(axis=0)
    with np.errstate(invalid="ignore"):
        w = np.where(np.isneginf(m), 0.0, np.exp(m - M[None, :]))
    return M, (s * w).sum(axis=0)


# --------------------------------------------------------------------------------------
# torch-CPU functional restatement of the models (a5-a19); autograd gives the backward.
# Parameters are a flat dict keyed exactly like the reference ``state_dict``.
# --------------------------------------------------------------------------------------
def make_args(**kw) -> SimpleNamespace:
    """The subset of the reference's argparse Namespace the hot path reads
    (density_estimation.py:27-93)."""
    d = dict(model_name="vae", prior="exemplar_prior", input_type="binary", input_size=[1, 28, 28],
             hidden_size=300, z1_size=40, z2_size=40, number_components=1000, training_set_size=2000,
             approximate_prior=False, approximate_k=10, no_mask=False, no_attention=False,
             same_variational_var=False, use_logit=False, lambd=1e-4, bottleneck=6,
             dataset_name="dynamic_mnist", device="cpu", dynamic_binarization=True, warmup=100,
             batch_size=100, lr=5e-4, continuous=False)
    d.update(kw)
    return SimpleNamespace(**d)


def t_pairwise_distance(z: torch.Tensor, means: torch.Tensor) -> torch.Tensor:
    """utils/distributions.py:12-18 (torch, differentiable)."""
    zd, md = z.double(), means.double()
    d1 = (zd * zd).sum(1, keepdim=True)
    d2 = (md * md).sum(1).unsqueeze(0)
    d3 = zd @ md.t()
    return ((d1 + d2) + (-2.0 * d3)).float()


def t_log_normal_diag(x, mean, log_var):
    """utils/distributions.py:28-33, dim=1."""
    return (-0.5 * (log_var + LOG_2PI + (x - mean) ** 2 / torch.exp(log_var))).sum(1)


def t_log_bernoulli(x, mean):
    """utils/distributions.py:44-51, dim=1."""
    p = torch.clamp(mean, min=MIN_EPS, max=MAX_EPS)
    return (x * torch.log(p) + (1.0 - x) * torch.log(1.0 - p)).sum(1)


def t_log_logistic_256(x, mean, logvar):
    """utils/distributions.py:54-66, dim=1."""
    bin_size = 1.0 / 256.0
    scale = torch.exp(logvar)
    xs = (torch.floor(x / bin_size) * bin_size - mean) / scale
    return torch.log(torch.sigmoid(xs + bin_size / scale) - torch.sigmoid(xs) + 1e-7).sum(1)


def t_log_p_z_exemplar(z, z_indices, centers, log_var_row, center_indices, masked: bool):
    """models/BaseModel.py:98-128 (exemplar prior, sum=True).  log_var_row: [D]."""
    C = centers.shape[0]
    sd = torch.exp(0.5 * log_var_row).unsqueeze(0)
    pd = t_pairwise_distance(z / sd, centers / sd)
    prob = -0.5 * torch.sum(log_var_row + LOG_2PI) - 0.5 * pd
    denom = torch.full((z.shape[0],), float(C))
    if masked:
        mask = z_indices.reshape(-1, 1) == center_indices.reshape(1, -1)
        prob = prob.masked_fill(mask, float("-inf"))
        denom = denom - mask.sum(1).float()
    prob = prob - torch.log(denom).unsqueeze(1)
    pmax = prob.max(1).values
    return pmax + torch.log(torch.exp(prob - pmax.unsqueeze(1)).sum(1))


def _lin(p, name, x):
    return torch.nn.functional.linear(x, p[name + ".weight"], p.get(name + ".bias"))


def t_gated_dense(p: Dict[str, torch.Tensor], name: str, x):
    """utils/nn.py:44-69 with activation=None, no_attention=False: h(x) * sigmoid(g(x))."""
    return _lin(p, name + ".h", x) * torch.sigmoid(_lin(p, name + ".g", x))


def t_gated_conv(p, name, x, stride, padding):
    """utils/nn.py:72-95 with activation=None: conv_h(x) * sigmoid(conv_g(x))."""
    F = torch.nn.functional
    h = F.conv2d(x, p[name + ".h.weight"], p[name + ".h.bias"], stride=stride, padding=padding)
    g = F.conv2d(x, p[name + ".g.weight"], p[name + ".g.bias"], stride=stride, padding=padding)
    return h * torch.sigmoid(g)


def t_hardtanh_linear(p, name, x, lo=-6.0, hi=2.0):
    """utils/nn.py:29-41 NonLinear with Hardtanh activation (logvar heads, models/VAE.py:25-26)."""
    return torch.clamp(_lin(p, name + ".linear", x), lo, hi)


# ---- model_name == 'vae' (models/VAE.py:15-30, models/AbsModel.py) --------------------
def vae_q_z(p, args, x, prior=False):
    """models/BaseModel.py:205-221 for the MLP VAE."""
    h = t_gated_dense(p, "q_z_layers.0", x)
    h = t_gated_dense(p, "q_z_layers.1", h)
    mean = _lin(p, "q_z_mean", h)
    if prior and args.prior == "exemplar_prior":
        logvar = p["prior_log_variance"] * torch.ones((x.shape[0], args.z1_size))
    else:
        logvar = t_hardtanh_linear(p, "q_z_logvar", h)
    return mean, logvar


def vae_p_x(p, args, z):
    """models/AbsModel.py:31-42 for binary inputs: decoder trunk + sigmoid head."""
    h = t_gated_dense(p, "p_x_layers.0", z)
    h = t_gated_dense(p, "p_x_layers.1", h)
    return torch.sigmoid(_lin(p, "p_x_mean.linear", h))


def vae_loss(p, args, x, x_indices, eps, exemplars, exemplar_indices, beta=1.0, average=True,
             masked=True, exemplars_embedding=None):
    """models/BaseModel.py:65-77 + models/AbsModel.py:13-19,44-49 with the RNG draws injected:
    ``eps`` replaces BaseModel.py:81, ``exemplar_indices``/``exemplars`` replace :245-247."""
    mean, logvar = vae_q_z(p, args, x)
    z = mean + torch.exp(0.5 * logvar) * eps
    x_mean = vae_p_x(p, args, z)
    RE = t_log_bernoulli(x, x_mean)
    if args.prior == "exemplar_prior":
        if exemplars_embedding is None:
            ex_mean, ex_logvar = vae_q_z(p, args, exemplars, prior=True)
            ex_idx = exemplar_indices
        else:
            ex_mean, ex_logvar, ex_idx = exemplars_embedding
        log_p = t_log_p_z_exemplar(z, x_indices, ex_mean, ex_logvar[0], ex_idx, masked)
    elif args.prior == "vampprior":
        log_p = t_log_p_z_vampprior(p, args, z, exemplars_embedding)
    else:
        log_p = (-0.5 * z * z - 0.5 * LOG_2PI).sum(1)
    log_q = t_log_normal_diag(z, mean, logvar)
    KL = -(log_p - log_q)
    loss = -RE + beta * KL
    if average:
        return loss.mean(), RE.mean(), KL.mean()
    return loss, RE, KL


def t_vamp_logprob_matrix(z, pm, plv):
    """models/BaseModel.py:90-96: log_normal_diag(z[:,None], mean[None], logvar[None], dim=2) - log C  -> [B,C]"""
    C = pm.shape[0]
    t = -0.5 * (plv[None] + LOG_2PI + (z[:, None, :] - pm[None]) ** 2 / torch.exp(plv[None]))
    return t.sum(2) - math.log(C)


def t_log_p_z_vampprior(p, args, z, exemplars_embedding=None):
    """models/BaseModel.py:84-96 + :123-125 — pseudo-inputs = Hardtanh(0,1)(I @ W^T) (BaseModel.py:130-139), encoded
    with q_z(prior=True) (both heads: the prior is not the exemplar prior), then max-shifted log-sum-exp."""
    if exemplars_embedding is None:
        X = torch.clamp(p["means.linear.weight"].t(), 0.0, 1.0)
        pm, plv = vae_q_z(p, args, X, prior=True)
    else:
        pm, plv = exemplars_embedding[0], exemplars_embedding[1]
    prob = t_vamp_logprob_matrix(z, pm, plv)
    pmax = prob.max(1)[0]
    return pmax + torch.log(torch.exp(prob - pmax[:, None]).sum(1))


# ---- model_name == 'hvae_2level' (models/HVAE_2level.py:15-66, models/AbsHModel.py) ---
def hvae_q_z(p, args, x, prior=False):
    """models/BaseModel.py:205-221: q(z2|x) of the 2-level MLP model."""
    h = t_gated_dense(p, "q_z_layers.0", x)
    h = t_gated_dense(p, "q_z_layers.1", h)
    mean = _lin(p, "q_z_mean", h)
    if prior and args.prior == "exemplar_prior":
        logvar = p["prior_log_variance"] * torch.ones((x.shape[0], args.z1_size))
    else:
        logvar = t_hardtanh_linear(p, "q_z_logvar", h)
    return mean, logvar


def hvae_loss(p, args, x, x_indices, eps2, eps1, exemplars, exemplar_indices, beta=1.0, average=True,
              masked=True, exemplars_embedding=None):
    """models/AbsHModel.py:13-29,45-106 for hvae_2level, binary input, RNG injected."""
    z2_mean, z2_logvar = hvae_q_z(p, args, x)
    z2 = z2_mean + torch.exp(0.5 * z2_logvar) * eps2
    # q(z1 | x, z2)  AbsHModel.py:45-54
    hx = t_gated_dense(p, "q_z1_layers_x.0", x)
    hz = t_gated_dense(p, "q_z1_layers_z2.0", z2)
    hj = t_gated_dense(p, "q_z1_layers_joint.0", torch.cat((hx, hz), 1))
    z1_mean = _lin(p, "q_z1_mean", hj)
    z1_logvar = t_hardtanh_linear(p, "q_z1_logvar", hj)
    z1 = z1_mean + torch.exp(0.5 * z1_logvar) * eps1
    # p(z1 | z2)  AbsHModel.py:39-43
    hp = t_gated_dense(p, "p_z1_layers_z2.0", z2)
    hp = t_gated_dense(p, "p_z1_layers_z2.1", hp)
    z1_p_mean = _lin(p, "p_z1_mean", hp)
    z1_p_logvar = t_hardtanh_linear(p, "p_z1_logvar", hp)
    # p(x | z1, z2)  AbsHModel.py:56-94
    d1 = t_gated_dense(p, "p_x_layers_z1.0", z1)
    d2 = t_gated_dense(p, "p_x_layers_z2.0", z2)
    dj = t_gated_dense(p, "p_x_layers_joint.0", torch.cat((d1, d2), 1))
    x_mean = torch.sigmoid(_lin(p, "p_x_mean.linear", dj))
    RE = t_log_bernoulli(x, x_mean)
    if exemplars_embedding is None:
        ex_mean, ex_logvar = hvae_q_z(p, args, exemplars, prior=True)
        ex_idx = exemplar_indices
    else:
        ex_mean, ex_logvar, ex_idx = exemplars_embedding
    log_p_z1 = t_log_normal_diag(z1, z1_p_mean, z1_p_logvar)
    log_q_z1 = t_log_normal_diag(z1, z1_mean, z1_logvar)
    log_p_z2 = t_log_p_z_exemplar(z2, x_indices, ex_mean, ex_logvar[0], ex_idx, masked)
    log_q_z2 = t_log_normal_diag(z2, z2_mean, z2_logvar)
    KL = -(log_p_z1 + log_p_z2 - log_q_z1 - log_q_z2)
    loss = -RE + beta * KL
    if average:
        return loss.mean(), RE.mean(), KL.mean()
    return loss, RE, KL


# ---- parameter construction (models/BaseModel.py:25-44, utils/nn.py:12-14) -------------
def _linear_shapes_vae(args):
    P = int(np.prod(args.input_size)); H = args.hidden_size; D = args.z1_size
    return [("q_z_layers.0.h", P, H), ("q_z_layers.0.g", P, H), ("q_z_layers.1.h", H, H),
            ("q_z_layers.1.g", H, H), ("q_z_mean", H, D), ("q_z_logvar.linear", H, D),
            ("p_x_layers.0.h", D, H), ("p_x_layers.0.g", D, H), ("p_x_layers.1.h", H, H),
            ("p_x_layers.1.g", H, H), ("p_x_mean.linear", H, P)]


def _linear_shapes_hvae(args):
    P = int(np.prod(args.input_size)); H = args.hidden_size; D1 = args.z1_size; D2 = args.z2_size
    s = []
    for n, i, o in [("q_z_layers.0", P, H), ("q_z_layers.1", H, H), ("q_z1_layers_x.0", P, H),
                    ("q_z1_layers_z2.0", D2, H), ("q_z1_layers_joint.0", 2 * H, H),
                    ("p_z1_layers_z2.0", D2, H), ("p_z1_layers_z2.1", H, H), ("p_x_layers_z1.0", D1, H),
                    ("p_x_layers_z2.0", D2, H), ("p_x_layers_joint.0", 2 * H, H)]:
        s += [(n + ".h", i, o), (n + ".g", i, o)]
    s += [("q_z_mean", H, D2), ("q_z_logvar.linear", H, D2), ("q_z1_mean", H, D1),
          ("q_z1_logvar.linear", H, D1), ("p_z1_mean", H, D1), ("p_z1_logvar.linear", H, D1),
          ("p_x_mean.linear", H, P)]
    return s


def init_params(args, seed=0) -> Dict[str, torch.Tensor]:
    """He-normal weights (utils/nn.py:12-14 applied to every nn.Linear, BaseModel.py:39-44),
    uniform(-1/sqrt(in), 1/sqrt(in)) biases (nn.Linear default), prior_log_variance ~ N(0,1)
    (BaseModel.py:25-26).  Deterministic from ``seed``; not bit-identical to the reference's
    own init (different RNG consumption order) — parity tests load identical weights instead."""
    g = torch.Generator().manual_seed(seed)
    shapes = _linear_shapes_vae(args) if args.model_name == "vae" else _linear_shapes_hvae(args)
    p: Dict[str, torch.Tensor] = {}
    if args.prior == "exemplar_prior":
        p["prior_log_variance"] = torch.randn(1, generator=g)
    if args.prior == "vampprior":       # BaseModel.py:130-139: N(pseudoinputs_mean, pseudoinputs_std)
        P = int(np.prod(args.input_size))
        p["means.linear.weight"] = -0.05 + 0.01 * torch.randn(P, args.number_components, generator=g)
    for name, fin, fout in shapes:
        p[name + ".weight"] = torch.randn(fout, fin, generator=g) * math.sqrt(2.0 / fin)
        bound = 1.0 / math.sqrt(fin)
        p[name + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound
    for v in p.values():
        v.requires_grad_(True)
    return p


def loss_fn(args):
    return {"vae": vae_loss, "hvae_2level": hvae_loss, "convhvae_2level": convhvae_loss,
            "single_conv": single_conv_loss}[args.model_name]


# ---- AdamNormGrad (utils/optimizer.py:32-80) -------------------------------------------
def adam_normgrad_step(params: Dict[str, torch.Tensor], state: Dict[str, dict], lr=5e-4,
                       betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """utils/optimizer.py:32-80: per-tensor grad / (||grad||_2 + 1e-7), then Adam."""
    b1, b2 = betas
    with torch.no_grad():
        for name, prm in params.items():
            if prm.grad is None:
                continue
            g = prm.grad / (torch.norm(prm.grad, 2) + 1e-7)
            st = state.setdefault(name, {"step": 0, "exp_avg": torch.zeros_like(prm),
                                         "exp_avg_sq": torch.zeros_like(prm)})
            st["step"] += 1
            if weight_decay != 0:
                g = g + weight_decay * prm
            st["exp_avg"].mul_(b1).add_(g, alpha=1 - b1)
            st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = st["exp_avg_sq"].sqrt().add_(eps)
            step_size = lr * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
            prm.addcdiv_(st["exp_avg"], denom, value=-step_size)


def train_step(params, opt_state, args, data, indices, dataset_x, beta, gen: torch.Generator,
               lr=5e-4, rng_override: Optional[dict] = None):
    """Loop body of utils/training.py:27-46 (exact exemplar prior): dynamic binarisation,
    exemplar re-sampling with replacement, loss, backward, AdamNormGrad step.
    ``rng_override`` may carry 'x', 'eps' (or 'eps2'/'eps1'), 'exemplar_indices' to replay
    the draws of another implementation."""
    ro = rng_override or {}
    x = ro["x"] if "x" in ro else torch.bernoulli(data, generator=gen)
    N = args.number_components
    ex_idx = ro["exemplar_indices"] if "exemplar_indices" in ro else \
        torch.randint(0, args.training_set_size, (N,), generator=gen)
    exemplars = dataset_x[ex_idx]
    for prm in params.values():
        prm.grad = None
    if args.model_name == "vae":
        eps = ro["eps"] if "eps" in ro else torch.randn(x.shape[0], args.z1_size, generator=gen)
        loss, RE, KL = vae_loss(params, args, x, indices, eps, exemplars, ex_idx, beta=beta)
    else:
        eps2 = ro["eps2"] if "eps2" in ro else torch.randn(x.shape[0], args.z2_size, generator=gen)
        eps1 = ro["eps1"] if "eps1" in ro else torch.randn(x.shape[0], args.z1_size, generator=gen)
        loss, RE, KL = hvae_loss(params, args, x, indices, eps2, eps1, exemplars, ex_idx, beta=beta)
    loss.backward()
    adam_normgrad_step(params, opt_state, lr=lr)
    return float(loss.detach()), float(RE.detach()), float(KL.detach())


def synthetic_dataset(T: int, P: int = 784, seed: int = 1234) -> torch.Tensor:
    """SURVEY.md §8d synthetic train set: P ~ U(0,1)^{T x 784}, CPU generator seed 1234
    (layout of utils/load_data/base_load_data.py:55-59: x float32 [T,P], indices int64 [T,1])."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(T, P, generator=g)


# ---- model_name == 'convhvae_2level' (models/convHVAE_2level.py:13-97, models/AbsHModel.py) ----
_CONV_QZ = [(7, 1, 3), (3, 2, 1), (5, 1, 2), (3, 2, 1), (3, 1, 1)]      # (kernel, stride, padding) of q_z_layers
_CONV_QZ1 = [(3, 1, 1), (3, 2, 1), (3, 1, 1), (3, 2, 1), (3, 1, 1)]     # q_z1_layers_x


def _gated_stack(p, name, x, spec):
    for i, (k, s, pd) in enumerate(spec):
        x = t_gated_conv(p, f"{name}.{i}", x, s, pd)
    return x


def convhvae_q_z(p, args, x, prior=False):
    """models/BaseModel.py:205-221 for convhvae_2level: 5 gated convs -> flatten (C,H,W) -> linear heads."""
    C, H, W = args.input_size
    h = _gated_stack(p, "q_z_layers", x.view(-1, C, H, W), _CONV_QZ).reshape(x.shape[0], -1)
    mean = _lin(p, "q_z_mean.linear", h)
    if prior and args.prior == "exemplar_prior":
        logvar = p["prior_log_variance"] * torch.ones((x.shape[0], args.z1_size))
    else:
        logvar = t_hardtanh_linear(p, "q_z_logvar", h)
    return mean, logvar


def convhvae_loss(p, args, x, x_indices, eps2, eps1, exemplars, exemplar_indices, beta=1.0, average=True,
                  masked=True, exemplars_embedding=None):
    """models/AbsHModel.py:13-29,45-106 for convhvae_2level, binary input, RNG injected."""
    C, H, W = args.input_size
    z2_mean, z2_logvar = convhvae_q_z(p, args, x)
    z2 = z2_mean + torch.exp(0.5 * z2_logvar) * eps2
    hx = _gated_stack(p, "q_z1_layers_x", x.view(-1, C, H, W), _CONV_QZ1).reshape(x.shape[0], -1)
    hz = t_gated_dense(p, "q_z1_layers_z2.0", z2)
    hj = t_gated_dense(p, "q_z1_layers_joint.0", torch.cat((hx, hz), 1))
    z1_mean = _lin(p, "q_z1_mean.linear", hj)
    z1_logvar = t_hardtanh_linear(p, "q_z1_logvar", hj)
    z1 = z1_mean + torch.exp(0.5 * z1_logvar) * eps1
    hp = t_gated_dense(p, "p_z1_layers_z2.1", t_gated_dense(p, "p_z1_layers_z2.0", z2))
    z1_p_mean = _lin(p, "p_z1_mean.linear", hp)
    z1_p_logvar = t_hardtanh_linear(p, "p_z1_logvar", hp)
    d = torch.cat((t_gated_dense(p, "p_x_layers_z1.0", z1), t_gated_dense(p, "p_x_layers_z2.0", z2)), 1)
    d = t_gated_dense(p, "p_x_layers_joint_pre.0", d).view(-1, C, H, W)
    d = _gated_stack(p, "p_x_layers_joint", d, [(3, 1, 1)] * 4)
    x_mean = torch.sigmoid(torch.nn.functional.conv2d(d, p["p_x_mean.conv.weight"], p["p_x_mean.conv.bias"]))
    RE = t_log_bernoulli(x, x_mean.reshape(x.shape[0], -1))
    if exemplars_embedding is None:
        ex_mean, ex_logvar = convhvae_q_z(p, args, exemplars, prior=True)
        ex_idx = exemplar_indices
    else:
        ex_mean, ex_logvar, ex_idx = exemplars_embedding
    KL = -(t_log_normal_diag(z1, z1_p_mean, z1_p_logvar) + t_log_p_z_exemplar(z2, x_indices, ex_mean, ex_logvar[0], ex_idx, masked)
           - t_log_normal_diag(z1, z1_mean, z1_logvar) - t_log_normal_diag(z2, z2_mean, z2_logvar))
    loss = -RE + beta * KL
    if average:
        return loss.mean(), RE.mean(), KL.mean()
    return loss, RE, KL


# ---- model_name == 'single_conv' (models/fully_conv.py:12-81, models/AbsModel.py) -------------
def _wn_conv(p, name, x, stride=1):
    """torch.nn.utils.weight_norm(nn.Conv2d(k=3, padding=1)): w = g * v / ||v|| per output channel."""
    v, g = p[name + ".weight_v"], p[name + ".weight_g"]
    w = v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1, 1))
    return torch.nn.functional.conv2d(x, w, p.get(name + ".bias"), stride=stride, padding=1)


def _res_blocks(p, name, x, first):
    elu = torch.nn.functional.elu
    for i in range(first, first + 6):                      # block: x + conv1(ELU(x))   fully_conv.py:13-23
        x = x + _wn_conv(p, f"{name}.{i}.conv1", elu(x))
    return x


def single_conv_q_z(p, args, x, prior=False):
    elu = torch.nn.functional.elu
    C, H, W = args.input_size
    h = elu(_wn_conv(p, "q_z_layers.0", x.view(-1, C, H, W), stride=2))
    h = _res_blocks(p, "q_z_layers", h, 2)
    h = elu(_wn_conv(p, "q_z_layers.8", h, stride=2))
    h = _res_blocks(p, "q_z_layers", h, 10)
    mean = _wn_conv(p, "q_z_mean", h).reshape(-1, args.z1_size)
    if prior and args.prior == "exemplar_prior":
        logvar = p["prior_log_variance"] * torch.ones((x.shape[0], args.z1_size))
    else:
        logvar = _wn_conv(p, "q_z_logvar", h).reshape(-1, args.z1_size)
    return mean, logvar


def single_conv_loss(p, args, x, x_indices, eps, exemplars, exemplar_indices, beta=1.0, average=True, masked=True,
                     exemplars_embedding=None):
    """models/AbsModel.py:13-49 with the fully-conv architecture; binary or continuous (logistic-256) input."""
    elu = torch.nn.functional.elu
    up = lambda t: torch.nn.functional.interpolate(t, scale_factor=2)
    C, H, W = args.input_size
    mean, logvar = single_conv_q_z(p, args, x)
    z = mean + torch.exp(0.5 * logvar) * eps
    h = z.reshape(-1, args.bottleneck, H // 4, W // 4)
    h = elu(_wn_conv(p, "p_x_layers.1", up(h)))
    h = _res_blocks(p, "p_x_layers", h, 3)
    h = elu(_wn_conv(p, "p_x_layers.10", up(h)))
    h = _res_blocks(p, "p_x_layers", h, 12)
    P = C * H * W
    if args.input_type == "binary":
        x_mean = torch.sigmoid(torch.nn.functional.conv2d(h, p["p_x_mean.0.weight"], p["p_x_mean.0.bias"], padding=1))
        RE = t_log_bernoulli(x, x_mean.reshape(-1, P))
    else:
        x_mean = torch.clamp(_wn_conv(p, "p_x_mean", h), min=1. / 512., max=1. - 1. / 512.).reshape(-1, P)
        x_logvar = p["decoder_logstd"] * torch.ones_like(x_mean)
        RE = t_log_logistic_256(x, x_mean, x_logvar)
    if exemplars_embedding is None:
        ex_mean, ex_logvar = single_conv_q_z(p, args, exemplars, prior=True)
        ex_idx = exemplar_indices
    else:
        ex_mean, ex_logvar, ex_idx = exemplars_embedding
    log_p = t_log_p_z_exemplar(z, x_indices, ex_mean, ex_logvar[0], ex_idx, masked)
    KL = -(log_p - t_log_normal_diag(z, mean, logvar))
    loss = -RE + beta * KL
    if average:
        return loss.mean(), RE.mean(), KL.mean()
    return loss, RE, KL


# ---- deterministic synthetic parameters (large conv models: fixtures store seeds, not tensors) ----
def grad_projections(arr, name: str, k: int = 8) -> np.ndarray:
    """k dot products (float64) of the WHOLE flattened tensor with Rademacher (+-1) vectors seeded by the parameter name.
    The compact conv goldens keep these instead of full gradients: every element enters every projection, so an error
    anywhere in a tensor (a transposed filter, a swapped channel, a wrong tap order) moves them by O(||g||)."""
    import zlib
    a = np.asarray(arr, dtype=np.float64).reshape(-1)
    rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    out = np.empty(k, dtype=np.float64)
    for i in range(k):
        sgn = rs.randint(0, 2, size=a.size).astype(np.float64) * 2.0 - 1.0
        out[i] = float(np.dot(sgn, a))
    return out


def synth_params(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Reproducible parameter values keyed by state_dict name: N(0, 1/fan_in) weights, small biases,
    positive weight-norm gains.  BatchNorm tensors of the never-applied ``block.normalization``
    (models/fully_conv.py:16) and integer buffers are left out (callers keep their own)."""
    import zlib
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if "normalization" in key or "num_batches_tracked" in key:
            continue
        # fully_conv blocks expose conv1 twice in the state_dict (``conv1.*`` and ``f.1.*``, one tensor)
        canon = key.replace(".f.1.", ".conv1.")
        g = torch.Generator().manual_seed((zlib.crc32(canon.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if key == "prior_log_variance":
            v = torch.full(shape, -1.3)
        elif key == "decoder_logstd":
            v = torch.full(shape, -0.4)
        elif key.endswith("weight_g"):
            v = 0.5 + torch.rand(shape, generator=g)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        else:
            v = 0.1 * torch.randn(shape, generator=g)
        out[key] = v
    return out
