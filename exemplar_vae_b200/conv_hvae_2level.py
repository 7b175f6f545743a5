"""``model_name='convhvae_2level'`` — mirror of models/convHVAE_2level.py:9-97 (gated-conv encoders,
gated-dense + gated-conv decoder, exemplar prior on z2).  Conv stacks run NHWC on the exvae kernels."""
from __future__ import annotations

import numpy as np
import torch.nn as nn

from .base_model import BaseHModel
from .layers import Conv2d, GatedConv2d, GatedDense, NonLinear


class VAE(BaseHModel):
    def __init__(self, args):
        super().__init__(args)

    def create_model(self, args):
        if args.dataset_name == 'freyfaces':
            self.h_size = 210
        elif args.dataset_name in ('cifar10', 'svhn'):
            self.h_size = 384
        else:
            self.h_size = 294
        fc_size = 300
        C = self.args.input_size[0]
        na = args.no_attention
        ht = lambda: nn.Hardtanh(min_val=-6., max_val=2.)
        # encoder: q(z2 | x)
        self.q_z_layers = nn.Sequential(
            GatedConv2d(C, 32, 7, 1, 3, no_attention=na), GatedConv2d(32, 32, 3, 2, 1, no_attention=na),
            GatedConv2d(32, 64, 5, 1, 2, no_attention=na), GatedConv2d(64, 64, 3, 2, 1, no_attention=na),
            GatedConv2d(64, 6, 3, 1, 1, no_attention=na))
        self.q_z_mean = NonLinear(self.h_size, self.args.z2_size, activation=None)
        self.q_z_logvar = NonLinear(self.h_size, self.args.z2_size, activation=ht())
        # encoder: q(z1 | x, z2)
        self.q_z1_layers_x = nn.Sequential(
            GatedConv2d(C, 32, 3, 1, 1, no_attention=na), GatedConv2d(32, 32, 3, 2, 1, no_attention=na),
            GatedConv2d(32, 64, 3, 1, 1, no_attention=na), GatedConv2d(64, 64, 3, 2, 1, no_attention=na),
            GatedConv2d(64, 6, 3, 1, 1, no_attention=na))
        self.q_z1_layers_z2 = nn.Sequential(GatedDense(self.args.z2_size, self.h_size))
        self.q_z1_layers_joint = nn.Sequential(GatedDense(2 * self.h_size, fc_size))
        self.q_z1_mean = NonLinear(fc_size, self.args.z1_size, activation=None)
        self.q_z1_logvar = NonLinear(fc_size, self.args.z1_size, activation=ht())
        # decoder: p(z1 | z2)
        self.p_z1_layers_z2 = nn.Sequential(GatedDense(self.args.z2_size, fc_size, no_attention=na),
                                            GatedDense(fc_size, fc_size, no_attention=na))
        self.p_z1_mean = NonLinear(fc_size, self.args.z1_size, activation=None)
        self.p_z1_logvar = NonLinear(fc_size, self.args.z1_size, activation=ht())
        # decoder: p(x | z1, z2)
        self.p_x_layers_z1 = nn.Sequential(GatedDense(self.args.z1_size, fc_size, no_attention=na))
        self.p_x_layers_z2 = nn.Sequential(GatedDense(self.args.z2_size, fc_size, no_attention=na))
        self.p_x_layers_joint_pre = nn.Sequential(
            GatedDense(2 * fc_size, int(np.prod(self.args.input_size)), no_attention=na))
        self.p_x_layers_joint = nn.Sequential(
            GatedConv2d(C, 64, 3, 1, 1, no_attention=na), GatedConv2d(64, 64, 3, 1, 1, no_attention=na),
            GatedConv2d(64, 64, 3, 1, 1, no_attention=na), GatedConv2d(64, 64, 3, 1, 1, no_attention=na))
        if self.args.input_type == 'binary':
            self.p_x_mean = Conv2d(64, 1, 1, 1, 0, activation=nn.Sigmoid())
        elif self.args.input_type in ('gray', 'continuous'):
            self.p_x_mean = Conv2d(64, C, 1, 1, 0)
            self.p_x_logvar = Conv2d(64, C, 1, 1, 0, activation=nn.Hardtanh(min_val=-4.5, max_val=0.))

    def forward(self, x, zq=None):
        return super().forward(x, zq=zq)
