"""Sample-path factory and small helpers (the hot-path part of
``/root/reference/src/utils.py``)."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .sample_paths import (MultiOutputSamplePathBatch, SamplePath, SamplePathBatch, draw_dkgp_linear_paths,
                           draw_dkgp_rff_paths, draw_matheron_paths, draw_reference_paths, posterior_mean_path)

NUM_RFF_FEATURES = 1000  # hard-coded in the reference (src/utils.py:215,227,240)


def get_function_sample_batch(model, num_samples: int, **kwargs) -> SamplePathBatch:
    """S posterior sample paths of ``model`` resident on the GPU.

    kwargs: ``sampler`` = "matheron" (default: shared basis + exact-kernel update, one
    n x n Cholesky for all S) or "reference" (botorch ``get_gp_samples`` scheme, one basis
    and one L x L posterior per sample); ``num_rff_features`` (default 1000);
    ``generator``; ``device``; ``engine``."""
    sampler = kwargs.get("sampler", "matheron")
    L = kwargs.get("num_rff_features", NUM_RFF_FEATURES)
    gen = kwargs.get("generator", None)
    device = kwargs.get("device", None)
    from .models.deep_kernel_gp import DKGP
    from .models.gp import ModelListGP
    if isinstance(model, DKGP):
        if model.kernel == "linear":  # exact weight-space draw (utils.py:186-206)
            return draw_dkgp_linear_paths(model, num_samples, gen, device, kwargs.get("engine", "auto"))
        return draw_dkgp_rff_paths(model, num_samples, L, gen, device, kwargs.get("engine", "auto"))  # :207-221
    if isinstance(model, ModelListGP):  # one batch of paths per output model (utils.py:231-248)
        return MultiOutputSamplePathBatch([get_function_sample_batch(m, num_samples, **kwargs) for m in model.models])
    if sampler == "matheron":
        return draw_matheron_paths(model, num_samples, L, gen, device, kwargs.get("engine", "auto"))
    if sampler == "reference":
        return draw_reference_paths(model, num_samples, L, gen, device)
    raise ValueError(f"unknown sampler {sampler!r}")


def get_function_samples(model, **kwargs) -> SamplePath:
    """One posterior function sample as a callable — drop-in for
    ``src/utils.py:185-250``: ``f(X[b, 1, d]) -> [b]`` for a ``SingleTaskGP``, ``-> [b, 1, m]`` for a
    ``ModelListGP`` of m outputs (``:231-248``)."""
    return get_function_sample_batch(model, 1, **kwargs).path(0)


def posterior_mean_function(model, **kwargs) -> SamplePath:
    """``PosteriorMean(model)`` as a GPU path (``src/performance_metrics.py:117-122``)."""
    return posterior_mean_path(model, kwargs.get("device", None)).path(0)


def seed_torch(seed, verbose=True):
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_mesh(dim, steps):
    """``src/utils.py:273-286``: `dim` axes of linspace(0, 1, steps), 'ij' order."""
    axes = [torch.linspace(0, 1, steps) for _ in range(dim)]
    return torch.meshgrid(*axes, indexing="ij")


def reshape_mesh(xx):
    """``src/utils.py:264-271``: list of mesh tensors -> [num_points, dim]."""
    return torch.hstack([x.reshape(-1, 1) for x in xx])


def generate_random_points(num_points, input_dim, seed=None):
    """``src/utils.py:74-78``."""
    if seed is not None:
        old = torch.random.get_rng_state()
        torch.manual_seed(seed)
        X = torch.rand(num_points, input_dim)
        torch.random.set_rng_state(old)
        return X
    return torch.rand(num_points, input_dim)
