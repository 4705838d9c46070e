"""Shared builders for the tests: seeded synthetic GP states / path parameters."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import psbax_oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "algorithms_golden.npz")


def himmelblau(X):
    Z = 12.0 * X - 6.0
    a, b = Z[:, 0], Z[:, 1]
    return -((a ** 2 + b - 11) ** 2 + (a + b ** 2 - 7) ** 2)


def smooth_objective(X):
    w = torch.arange(1, X.shape[1] + 1, dtype=X.dtype)
    return torch.sin(3.0 * (X * w).sum(-1)) + 0.5 * torch.cos(5.0 * X[:, 0])


def make_gp(n, d, kernel="matern52", seed=0, noise=1e-3, ls=None):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, d, dtype=torch.float64, generator=g)
    y = himmelblau(X) if d == 2 else smooth_objective(X)
    if ls is None:
        ls = 0.1 + 0.4 * torch.rand(d, dtype=torch.float64, generator=g)
    return O.GPState(train_X=X, train_Y=y, lengthscale=ls, outputscale=1.0, noise=noise,
                     kernel=kernel, mean_const=0.0)


def random_params(S, L, d, n, G=1, kernel="matern52", seed=0, vscale=1.0):
    """Arbitrary (not posterior-consistent) parameters: exercises the evaluator only."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, dtype=torch.float64, generator=g)  # noqa: E731
    omega = 6.0 * r(G, L, d)
    bias = 2 * math.pi * torch.rand(G, L, dtype=torch.float64, generator=g)
    W = math.sqrt(2.0 / max(L, 1)) * r(S, L)
    p = O.PathParams(omega=omega, bias=bias, W=W, kernel=kernel, mean_const=0.1, y_mean=-0.7, y_std=1.9)
    if n > 0:
        ls = 0.15 + 0.4 * torch.rand(d, dtype=torch.float64, generator=g)
        p.Xtrain = torch.rand(n, d, dtype=torch.float64, generator=g) / ls
        p.inv_ls = 1.0 / ls
        p.V = vscale * r(S, n)
    return p.cast(torch.float32)


def to_batch(p, device="cuda", engine="auto"):
    from psbax_b200 import SamplePathBatch

    return SamplePathBatch(p.omega, p.bias, p.W, p.kernel, p.Xtrain, p.inv_ls, p.V, p.mean_const,
                           p.y_mean, p.y_std, device=device, engine=engine)


def candidates(N, d, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(N, d, dtype=torch.float32, generator=g)


def unpack_mask(words: torch.Tensor, N: int) -> np.ndarray:
    w = words.cpu().numpy().astype(np.uint32)
    bits = ((w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[0], -1)
    return bits[:, :N].astype(bool)
