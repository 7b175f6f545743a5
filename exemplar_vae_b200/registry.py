"""Model registry and checkpoint helpers — mirror of utils/utils.py:4-32."""
from __future__ import annotations

import os

import torch


def importing_model(args):
    """utils/utils.py:4-19 — returns the class registered under ``args.model_name``."""
    if args.model_name == 'vae':
        from .vae import VAE
    elif args.model_name == 'hvae_2level':
        from .hvae_2level import VAE
    elif args.model_name == 'convhvae_2level':
        from .conv_hvae_2level import VAE
    elif args.model_name == 'single_conv':
        from .fully_conv import VAE
    elif args.model_name == 'pixelcnn':
        raise NotImplementedError("model_name='pixelcnn' is out of scope (SURVEY.md §2)")
    else:
        raise Exception('Wrong name of the model!')
    return VAE


def save_model(save_path, load_path, content):
    """utils/utils.py:22-24 — atomic: write temp, then rename."""
    torch.save(content, save_path)
    os.rename(save_path, load_path)


def load_model(load_path, model, optimizer=None):
    """utils/utils.py:27-32"""
    checkpoint = torch.load(load_path, map_location=next(model.parameters()).device)
    model.load_state_dict(checkpoint['state_dict'])
    if optimizer is not None:
        optimizer.load_state_dict(checkpoint['optimizer'])
    return checkpoint
