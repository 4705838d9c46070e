"""make(): random sample-path batches of a given shape (shared by the profiling scripts)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import SamplePathBatch

dev = torch.device("cuda:0")


def make(S, L, d, n, engine, G=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    omega = 6 * torch.randn(G, L, d, generator=g)
    bias = 6.28 * torch.rand(G, L, generator=g)
    W = torch.randn(S, L, generator=g) * (2.0 / max(L, 1)) ** 0.5
    kw = {}
    if n:
        kw = dict(Xtrain=torch.rand(n, d, generator=g) / 0.3, inv_ls=torch.full((d,), 1 / 0.3),
                  V=torch.randn(S, n, generator=g))
    return SamplePathBatch(omega, bias, W, "matern52", device=dev, engine=engine, **kw)


