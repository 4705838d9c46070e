"""Shared problem builder of the projection-variant probes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psbax_b200.sample_paths import SamplePathBatch

dev = torch.device("cuda:0")


def make(S, L, d, n, engine, kernel="matern52", seed=0):
    g = torch.Generator().manual_seed(seed)
    omega = (torch.randn(1, L, d, generator=g) * 3.0).to(dev)
    bias = (torch.rand(1, L, generator=g) * 6.2831853).to(dev)
    W = (torch.randn(S, L, generator=g) * (2.0 / max(L, 1)) ** 0.5).to(dev)
    kw = {}
    if n:
        kw = dict(Xtrain=torch.rand(n, d, generator=g).to(dev), inv_ls=torch.full((d,), 2.0).to(dev),
                  V=torch.randn(S, n, generator=g).to(dev) * 0.3)
    return SamplePathBatch(omega, bias, W, kernel, device=dev, engine=engine, **kw)


