"""exemplar_vae_b200 — B200-native (sm_100a) implementation of the Exemplar-VAE training hot
path behind the reference's own Python plug points (model_name / prior / calculate_loss /
log_p_z).  Importing the package loads ``csrc/libexvae_b200.so``; there is no CPU fallback."""
from ._lib import ExvaeError, lib

lib()  # fail loudly at import time if the CUDA library is missing or stale

from . import ops  # noqa: E402
from .registry import importing_model, load_model, save_model  # noqa: E402
from .optimizer import AdamNormGrad  # noqa: E402
from .training import GraphedTrainStep, set_beta, train_one_epoch  # noqa: E402

__all__ = ["ExvaeError", "lib", "ops", "importing_model", "load_model", "save_model", "AdamNormGrad",
           "GraphedTrainStep", "set_beta", "train_one_epoch"]
