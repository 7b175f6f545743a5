"""``model_name='single_conv'`` — mirror of models/fully_conv.py:8-81: weight-normed 3x3 convs, ELU,
residual blocks ``x + conv(ELU(x))``, 2x nearest upsampling; latent = bottleneck x H/4 x W/4.  The
``block.normalization`` BatchNorm of the reference is constructed (its tensors are part of the
state_dict) but never applied (fully_conv.py:16,20-23) — same here."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_SIGMOID
from .base_model import AbsModel
from .layers import ELU, PlainConv2d, Sigmoid, Upsample2x, WNConv2d


class _SigmoidConv2d(nn.Conv2d):
    def forward(self, x):
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0], ACT_SIGMOID)


class block(nn.Module):
    def __init__(self, input_size, output_size, stride=1, kernel=3, padding=1):
        super().__init__()
        self.normalization = nn.BatchNorm2d(input_size)
        self.conv1 = WNConv2d(input_size, output_size, kernel, stride, padding, bias=True)
        self.activation = ELU()
        self.f = torch.nn.Sequential(self.activation, self.conv1)

    def forward(self, x):
        return ops.lincomb((1.0, 1.0), x, self.f(x))


class VAE(AbsModel):
    def __init__(self, args):
        super().__init__(args)

    def create_model(self, args, train_data_size=None):
        self.train_data_size = train_data_size
        self.cs = cs = 48
        self.bottleneck = b = self.args.bottleneck
        C = self.args.input_size[0]
        blocks = lambda ch: [block(ch, ch) for _ in range(6)]
        self.q_z_layers = nn.Sequential(WNConv2d(C, cs, 3, 2, 1), ELU(), *blocks(cs),
                                        WNConv2d(cs, cs * 2, 3, 2, 1), ELU(), *blocks(cs * 2))
        self.q_z_mean = WNConv2d(cs * 2, b, 3, 1, 1)
        self.q_z_logvar = WNConv2d(cs * 2, b, 3, 1, 1)
        self.p_x_layers = nn.Sequential(Upsample2x(), WNConv2d(b, cs * 2, 3, 1, 1), ELU(), *blocks(cs * 2),
                                        Upsample2x(), WNConv2d(cs * 2, cs, 3, 1, 1), ELU(), *blocks(cs))
        if self.args.input_type == 'binary':
            self.p_x_mean = nn.Sequential(_SigmoidConv2d(cs, C, 3, 1, 1), Sigmoid())
        elif self.args.input_type in ('gray', 'continuous'):
            self.p_x_mean = WNConv2d(cs, C, 3, 1, 1)
            self.p_x_logvar = PlainConv2d(cs, C, 3, 1, 1)
