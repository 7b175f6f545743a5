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
        self._pack()

    def _pack(self):
        """Keep ``h.weight`` and ``g.weight`` ADJACENT in one [2*O, K] buffer (the two parameters become views of it;
        names, shapes and ``state_dict`` are unchanged): the fused GEMM reads [Wh ; Wg] as one operand, and with the
        weights stored that way the per-call concat copy (one launch per layer per step) disappears.  A module whose
        weights are not adjacent (e.g. after ``copy.deepcopy``) still works: the C side copies then."""
        if self.no_attention is not False:
            return
        h, g = self.h.weight, self.g.weight
        if h.device != g.device or h.dtype != g.dtype or h.shape != g.shape:
            return
        if g.data_ptr() == h.data_ptr() + h.numel() * h.element_size():
            return
        O, K = h.shape
        with torch.no_grad():
            buf = torch.empty((2 * O, K), dtype=h.dtype, device=h.device)
            buf[:O].copy_(h.data)
            buf[O:].copy_(g.data)
            h.data = buf[:O]
            g.data = buf[O:]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)       # .cuda() / .to(): every parameter moves on its own
        self._pack()
        return out

    def forward(self, x):
        if self.no_attention is False:
            return ops.gated_dense(x, self.h.weight, self.h.bias, self.g.weight, self.g.bias)
        return ops.linear(x, self.h.weight, self.h.bias, ACT_RELU)


# ----------------------------------------------------------------------------------------------
# convolutional layers.  Activations between these modules are NHWC ([N, H, W, C]); the model
# classes convert at the stack boundaries (free when C == 1).
# ----------------------------------------------------------------------------------------------
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class GatedConv2d(nn.Module):
    """utils/nn.py:72-95 — ``act(h(x)) * sigmoid(g(x))`` as one im2col + fused GEMM (h and g share the
    patch matrix)."""

    def __init__(self, input_channels, output_channels, kernel_size, stride, padding, dilation=1, activation=None,
                 no_attention=False):
        super().__init__()
        if dilation != 1 or activation is not None or no_attention:
            raise NotImplementedError("GatedConv2d: only dilation=1, activation=None, no_attention=False are used "
                                      "by the in-scope models (the reference's no_attention branch does not run)")
        self.no_attention = no_attention
        self.activation = activation
        self.sigmoid = nn.Sigmoid()
        self.h = nn.Conv2d(input_channels, output_channels, kernel_size, stride, padding, dilation)
        self.g = nn.Conv2d(input_channels, output_channels, kernel_size, stride, padding, dilation)
        self._stride, self._pad = _pair(stride)[0], _pair(padding)[0]

    def forward(self, x):
        return ops.conv2d_gated(x, self.h.weight, self.h.bias, self.g.weight, self.g.bias, self._stride, self._pad)


class Conv2d(nn.Module):
    """utils/nn.py:100-114 — ``activation(conv(x))`` with the activation fused in the GEMM epilogue."""

    def __init__(self, input_channels, output_channels, kernel_size, stride, padding, dilation=1, activation=None,
                 bias=True):
        super().__init__()
        if dilation != 1:
            raise NotImplementedError("dilation != 1")
        self.activation = activation
        self.conv = nn.Conv2d(input_channels, output_channels, kernel_size, stride, padding, dilation, bias=bias)
        self._stride, self._pad = _pair(stride)[0], _pair(padding)[0]
        self._act = _act_code(activation)

    def forward(self, x):
        act, lo, hi = self._act
        return ops.conv2d(x, self.conv.weight, self.conv.bias, self._stride, self._pad, act, lo, hi)


class PlainConv2d(nn.Conv2d):
    """A bare ``nn.Conv2d`` (models/fully_conv.py:77,81) evaluated by the exvae kernels (NHWC)."""

    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0])


class WNConv2d(nn.Module):
    """``torch.nn.utils.weight_norm(nn.Conv2d(...))`` (models/fully_conv.py:16-17,28,...): parameters
    ``weight_g`` [Cout,1,1,1], ``weight_v`` [Cout,Cin,kh,kw], ``bias`` — same state_dict keys;
    w = g * v / ||v||_2 (norm over each output channel)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        ref = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        # registration order = torch.nn.utils.weight_norm(nn.Conv2d): bias, weight_g, weight_v.  Optimizer state is
        # positional, so a reference checkpoint's Adam moments only line up with this order.
        self.bias = nn.Parameter(ref.bias.detach().clone()) if bias else None
        self.weight_g = nn.Parameter(ref.weight.detach().flatten(1).norm(dim=1).view(-1, 1, 1, 1).clone())
        self.weight_v = nn.Parameter(ref.weight.detach().clone())
        self._stride, self._pad = _pair(stride)[0], _pair(padding)[0]

    def weight(self):
        return ops.weight_norm(self.weight_v, self.weight_g)

    def forward(self, x, act=ACT_NONE, lo=0.0, hi=0.0):
        return ops.conv2d(x, self.weight(), self.bias, self._stride, self._pad, act, lo, hi)


class ELU(nn.Module):
    """torch.nn.ELU (alpha=1) on the exvae kernels."""

    def forward(self, x):
        return ops.elu(x)


class Upsample2x(nn.Module):
    """nn.Upsample(scale_factor=2) (nearest), NHWC."""

    def forward(self, x):
        return ops.upsample2x(x)


class Sigmoid(nn.Module):
    """Marker for a sigmoid that the preceding conv fuses (kept so that nn.Sequential indices — and
    therefore state_dict keys — match the reference)."""

    def forward(self, x):
        return x
