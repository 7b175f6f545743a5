"""``model_name='hvae_2level'`` — mirror of models/HVAE_2level.py:11-66 (two-level MLP VAE,
exemplar prior on z2)."""
from __future__ import annotations

import numpy as np
import torch.nn as nn

from .base_model import BaseHModel
from .layers import GatedDense, Linear, NonLinear


class VAE(BaseHModel):
    def __init__(self, args):
        super().__init__(args)

    def create_model(self, args):
        self.args = args
        P = int(np.prod(self.args.input_size))
        H, D1, D2 = self.args.hidden_size, self.args.z1_size, self.args.z2_size
        ht = lambda: nn.Hardtanh(min_val=-6., max_val=2.)
        # encoder: q(z2 | x)
        self.q_z_layers = nn.Sequential(GatedDense(P, H), GatedDense(H, H))
        self.q_z_mean = Linear(H, D2)
        self.q_z_logvar = NonLinear(H, D2, activation=ht())
        # encoder: q(z1 | x, z2)
        self.q_z1_layers_x = nn.Sequential(GatedDense(P, H))
        self.q_z1_layers_z2 = nn.Sequential(GatedDense(D2, H))
        self.q_z1_layers_joint = nn.Sequential(GatedDense(2 * H, H))
        self.q_z1_mean = Linear(H, D1)
        self.q_z1_logvar = NonLinear(H, D1, activation=ht())
        # decoder: p(z1 | z2)
        self.p_z1_layers_z2 = nn.Sequential(GatedDense(D2, H), GatedDense(H, H))
        self.p_z1_mean = Linear(H, D1)
        self.p_z1_logvar = NonLinear(H, D1, activation=ht())
        # decoder: p(x | z1, z2)
        self.p_x_layers_z1 = nn.Sequential(GatedDense(D1, H))
        self.p_x_layers_z2 = nn.Sequential(GatedDense(D2, H))
        self.p_x_layers_joint = nn.Sequential(GatedDense(2 * H, H))
