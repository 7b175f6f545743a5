"""Mirror of utils/knn_on_latent.py:4-9 on the K2 kernel."""
from . import ops


def find_nearest_neighbors(z_val, z_train, z_train_log_var=None):
    """20 nearest training latents (sqrt-Euclidean, ascending) for every row of ``z_val`` -> [B,20] int64."""
    idx, _ = ops.knn_topk(z_val, z_train, 20, metric=1)
    return idx
