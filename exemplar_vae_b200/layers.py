"""Layer zoo of the in-scope architectures — host-side mirror of the reference's
``utils/nn.py:1-114`` (same class names, constructor arguments and state_dict keys), with the
arithmetic dispatched to the exvae_b200 CUDA kernels."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_HARDTANH, ACT_NONE, ACT_RELU, ACT_SIGMOID


def he_init(m):
    """utils/nn.py:12-14"""
    s = np.sqrt(2. / m.in_features)
    m.weight.data.normal_(0, s)


def xavier_init(m):
    """utils/nn.py:7-9"""
    s = np.sqrt(2. / (m.in_features + m.out_features))
    m.weight.data.normal_(0, s)


def normal_init(m, mean=0., std=0.01):
    """utils/nn.py:17-18"""
    m.weight.data.normal_(mean, std)


def _act_code(activation):
    if activation is None:
        return ACT_NONE, 0.0, 0.0
    if isinstance(activation, nn.Sigmoid):
        return ACT_SIGMOID, 0.0, 0.0
    if isinstance(activation, nn.Hardtanh):
        return ACT_HARDTANH, float(activation.min_val), float(activation.max_val)
    if isinstance(activation, nn.ReLU):
        return ACT_RELU, 0.0, 0.0
    raise NotImplementedError(
        f"activation {activation!r} has no fused exvae_b200 epilogue (supported: None, Sigmoid, Hardtanh, ReLU)")


class NonLinear(nn.Module):
    """utils/nn.py:29-41 — ``activation(linear(x))`` as one GEMM with a fused epilogue."""

    def __init__(self, input_size, output_size, bias=True, activation=None):
        super().__init__()
        self.activation = activation
        self.linear = nn.Linear(int(input_size), int(output_size), bias=bias)
        self._act = _act_code(activation)

    def forward(self, x):
        act, lo, hi = self._act
        return ops.linear(x, self.linear.weight, self.linear.bias, act, lo, hi)


class Linear(nn.Linear):
    """torch.nn.Linear with the exvae_b200 GEMM (models/VAE.py:21 uses a bare nn.Linear)."""

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class GatedDense(nn.Module):
    """utils/nn.py:44-69 — ``h(x) * sigmoid(g(x))``; both branches in one fused GEMM."""

    def __init__(self, input_size, output_size, activation=None, no_attention=False):
        super().__init__()
        self.activation = activation
        self.no_attention = no_attention
        self.sigmoid = nn.Sigmoid()
        self.h = nn.Linear(input_size, output_size)
        if no_attention is False:
            self.g = nn.Linear(input_size, output_size)
        else:
            self.activation = torch.nn.ReLU()
        if self.no_attention is False and self.activation is not None:
            raise NotImplementedError("GatedDense with an inner activation is not used by the in-scope models")

    def forward(self, x):
        if self.no_attention is False:
            return ops.gated_dense(x, self.h.weight, self.h.bias, self.g.weight, self.g.bias)
        return ops.linear(x, self.h.weight, self.h.bias, ACT_RELU)
