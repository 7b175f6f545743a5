"""Flag defaults of the reference's CLI (density_estimation.py:27-93) for the fields the hot-path
model classes read, and the synthetic train-set layout used by bench.py
(utils/load_data/base_load_data.py:55-59: x float32 [T,P], indices int64 [T,1], labels)."""
from __future__ import annotations

from argparse import Namespace

import torch


def default_args(**kw) -> Namespace:
    d = dict(model_name="vae", prior="exemplar_prior", input_type="binary", input_size=[1, 28, 28],
             hidden_size=300, z1_size=40, z2_size=40, number_components=25000, training_set_size=50000,
             approximate_prior=False, approximate_k=10, no_mask=False, no_attention=False,
             same_variational_var=False, use_logit=False, lambd=1e-4, bottleneck=6,
             dataset_name="dynamic_mnist", device="cuda", dynamic_binarization=True, warmup=100,
             batch_size=100, test_batch_size=100, lr=5e-4, continuous=False, seed=14, epochs=2000,
             early_stopping_epochs=50)
    d.update(kw)
    return Namespace(**d)


def synthetic_train_set(T: int, P: int = 784, seed: int = 1234) -> torch.utils.data.TensorDataset:
    """P ~ U(0,1)^{T x P} from a CPU generator, wrapped like the reference's train set."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(T, P, generator=g)
    return torch.utils.data.TensorDataset(x, torch.arange(T).view(-1, 1), torch.zeros(T))
