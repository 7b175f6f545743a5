"""Runs the UNMODIFIED reference training loop on CPU from ``oracle/_ref`` (see build_ref.py).

TEST / BASELINE INFRASTRUCTURE ONLY — used by ``bench.py --impl reference`` and the ``cpu_baseline``
leg.  The reference's own ``train_one_epoch`` (utils/training.py:15-51), model classes
(models/*.py via utils/utils.py:4-19 ``importing_model``) and ``AdamNormGrad`` (utils/optimizer.py)
run as they are; this file only builds the argparse Namespace they expect (field list of
SURVEY.md §5, defaults of density_estimation.py:27-93), a synthetic ``TensorDataset`` in the layout of
utils/load_data/base_load_data.py:55-59, and a loader that yields K batches.
"""
from __future__ import annotations

import os
import sys
import time
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "utils", "training.py"))


def _import_ref():
    """Import the vendored reference packages ``models`` / ``utils`` (top-level names, as the reference expects)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("models", "utils"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF):
            raise RuntimeError(f"a different top-level package '{name}' is already imported")
    import utils.optimizer as ref_opt          # noqa: E402
    import utils.training as ref_training      # noqa: E402
    import utils.utils as ref_utils            # noqa: E402
    return ref_utils, ref_training, ref_opt


def ref_args(**kw) -> Namespace:
    d = dict(model_name="vae", prior="exemplar_prior", input_type="binary", input_size=[1, 28, 28],
             hidden_size=300, z1_size=40, z2_size=40, number_components=25000, training_set_size=50000,
             approximate_prior=False, approximate_k=10, no_mask=False, no_attention=False,
             same_variational_var=False, use_logit=False, lambd=1e-4, bottleneck=6,
             dataset_name="dynamic_mnist", device="cpu", dynamic_binarization=True, warmup=100,
             batch_size=100, lr=5e-4, continuous=False, seed=14, use_training_data_init=False,
             pseudoinputs_mean=-0.05, pseudoinputs_std=0.01)
    d.update(kw)
    return Namespace(**d)


def run_steps(args: Namespace, data, batch: int, steps: int, warmup: int, threads: int, budget_s: float = 30.0):
    """``warmup`` untimed + up to ``steps`` timed training steps of the reference on CPU.  Returns a dict with
    ms_per_step (mean), the steps actually timed, and cache_z seconds (kNN mode: per-epoch cost, untimed)."""
    import torch
    ref_utils, ref_training, ref_opt = _import_ref()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    T = data.shape[0]
    dataset = torch.utils.data.TensorDataset(data, torch.arange(T).view(-1, 1), torch.zeros(T))
    model = ref_utils.importing_model(args)(args)
    opt = ref_opt.AdamNormGrad(model.parameters(), lr=args.lr)
    gen = torch.Generator().manual_seed(1)

    def epoch(k):
        """the reference's own epoch function over k batches of the full train set"""
        idx = torch.randint(0, T, (k * batch,), generator=gen).tolist()
        loader = torch.utils.data.DataLoader(dataset, batch_size=batch, sampler=idx)
        t0 = time.perf_counter()
        ref_training.train_one_epoch(1, args, loader, model, opt)
        return time.perf_counter() - t0

    cache_s = 0.0
    if args.approximate_prior:
        t0 = time.perf_counter()
        with torch.no_grad():
            model.cache_z(dataset)
        cache_s = time.perf_counter() - t0
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):       # the reference prints beta every epoch
        w = max(1, warmup)
        tw = epoch(w) - cache_s
        per = max(tw / w, 1e-3)
        k = max(1, min(steps, int(budget_s / per)))
        tt = epoch(k) - cache_s
    return {"ms_per_step": 1e3 * tt / k, "steps": k, "warmup": w, "cache_z_s": cache_s, "threads": threads}
