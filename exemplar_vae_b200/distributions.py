"""Host-side mirror of the reference's ``utils/distributions.py`` (same function names and
argument meaning) on top of the exvae_b200 kernels.  2-D inputs reduced over ``dim=1`` only —
the only way the hot path calls them."""
from __future__ import annotations

import math

import torch

from . import ops

min_epsilon = 1e-5
max_epsilon = 1. - 1e-5
log_2_pi = math.log(2 * math.pi)


def pairwise_distance(z, means):
    """utils/distributions.py:12-18 — [B,C] fp32, fp64 inside.  Selection primitive: no autograd."""
    return ops.pairwise_distance(z, means)


def log_normal_diag_vectorized(x, mean, log_var):
    """utils/distributions.py:21-25 — ``log_var`` is [1, D]; returns (log_normal, pair_dist), both [B,C].
    Materialising variant (no autograd); training uses the fused ``ops.prior_lse``."""
    return ops.log_normal_diag_vectorized(x, mean, log_var.reshape(-1))


def _reduce_args(t, average, dim):
    if t.dim() != 2 or dim not in (1, -1):
        raise NotImplementedError("exvae_b200 log-densities reduce 2-D tensors over dim=1")
    return (1.0 / t.shape[1]) if average else 1.0


def log_normal_diag(x, mean, log_var, average=False, dim=None):
    """utils/distributions.py:28-33"""
    scale = _reduce_args(x, average, dim)
    out = ops.log_normal_diag(x, mean.expand_as(x), log_var.expand_as(x))
    return out if scale == 1.0 else ops.lincomb((scale,), out)


def log_normal_standard(x, average=False, dim=None):
    """utils/distributions.py:36-41"""
    scale = _reduce_args(x, average, dim)
    out = ops.log_normal_standard(x)
    return out if scale == 1.0 else ops.lincomb((scale,), out)


def log_bernoulli(x, mean, average=False, dim=None):
    """utils/distributions.py:44-51"""
    scale = _reduce_args(mean, average, dim)
    out = ops.log_bernoulli(x, mean)
    return out if scale == 1.0 else ops.lincomb((scale,), out)


def log_logistic_256(x, mean, logvar, average=False, reduce=True, dim=None):
    """utils/distributions.py:54-66"""
    scale = _reduce_args(mean, average, dim)
    out = ops.log_logistic_256(x, mean, logvar.expand_as(mean))
    return out if scale == 1.0 else ops.lincomb((scale,), out)
