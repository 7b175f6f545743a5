"""``model_name='vae'`` — mirror of models/VAE.py:11-30 (MLP encoder/decoder, hidden 300, D=40)."""
from __future__ import annotations

import numpy as np
import torch.nn as nn

from .base_model import AbsModel
from .layers import GatedDense, Linear, NonLinear


class VAE(AbsModel):
    def __init__(self, args):
        super().__init__(args)

    def create_model(self, args, train_data_size=None):
        if getattr(args, "same_variational_var", False):
            raise NotImplementedError("same_variational_var=True is not callable in the reference either (VAE.py:22-23)")
        self.train_data_size = train_data_size
        P = int(np.prod(self.args.input_size))
        H, D = self.args.hidden_size, self.args.z1_size
        na = self.args.no_attention
        self.q_z_layers = nn.Sequential(GatedDense(P, H, no_attention=na), GatedDense(H, H, no_attention=na))
        self.q_z_mean = Linear(H, D)
        self.q_z_logvar = NonLinear(H, D, activation=nn.Hardtanh(min_val=-6., max_val=2.))
        self.p_x_layers = nn.Sequential(GatedDense(D, H, no_attention=na), GatedDense(H, H, no_attention=na))
