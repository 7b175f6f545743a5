"""Evaluation loops — mirror of utils/evaluation.py:11-33,56-69,72-103 (validation/test ELBO over the
full-train exemplar bank, IWAE log-likelihood).  Both reuse the fused exemplar-prior kernel at
larger shapes (B=100..5000 rows against N=|train| exemplars) without materialising any [B,N] matrix;
sums are accumulated on the device and read back once."""
from __future__ import annotations

import math

import torch

from . import ops


def load_all_pseudo_input(args, model, dataset):
    """utils/evaluation.py:56-69 — exemplar bank = embeddings of the whole training set."""
    if args.prior == 'exemplar_prior':
        with torch.no_grad():
            exemplars_z, exemplars_log_var = model.cache_z(dataset)
        return (exemplars_z, exemplars_log_var, torch.arange(len(exemplars_z), device=exemplars_z.device))
    if args.prior == 'standard':
        return None
    if args.prior == 'vampprior':      # utils/evaluation.py:60-64: embed the pseudo-inputs once
        with torch.no_grad():
            return model.pseudo_embedding()
    raise Exception('Wrong name of the prior!')


@torch.no_grad()
def evaluate_loss(args, model, loader, dataset=None, exemplars_embedding=None):
    """utils/evaluation.py:11-33 — returns (ELBO, -RE, KL) averaged over ``loader.dataset``."""
    model.eval()
    if exemplars_embedding is None:
        exemplars_embedding = load_all_pseudo_input(args, model, dataset)
    dev = next(model.parameters()).device
    acc = torch.zeros(3, dtype=torch.float64, device=dev)
    for data in loader:
        x = data[0].to(dev, non_blocking=True)
        loss, RE, KL = model.calculate_loss((x, None), average=False, exemplars_embedding=exemplars_embedding)
        acc += torch.stack((loss.sum(), -RE.sum(), KL.sum())).double()
    out = (acc / len(loader.dataset)).tolist()
    return out[0], out[1], out[2]


@torch.no_grad()
def calculate_likelihood(args, model, loader, S=5000, exemplars_embedding=None):
    """utils/evaluation.py:72-103 — importance-weighted log-likelihood with S samples per image:
    each image is expanded to S rows, one fused loss evaluation, logsumexp(-loss) - log S.
    Returns the negative mean log-likelihood (nats)."""
    model.eval()
    dev = next(model.parameters()).device
    lls = []
    for data in torch.utils.data.DataLoader(loader.dataset, batch_size=1):
        x = data[0].to(dev).expand(S, -1).contiguous()
        prob, _, _ = model.calculate_loss((x, None), exemplars_embedding=exemplars_embedding)
        ll = torch.logsumexp(-prob.double(), dim=0) - math.log(S)
        if getattr(args, "use_logit", False):
            lambd = float(args.lambd)
            xs = x[0].double()
            ll = ll - (-torch.nn.functional.softplus(-xs) - torch.nn.functional.softplus(xs)
                       - math.log((1 - 2 * lambd) / 256)).sum()
        lls.append(ll)
    return float(-torch.stack(lls).mean())
