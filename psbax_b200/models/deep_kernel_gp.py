"""Deep-kernel GP with a linear kernel on a feed-forward embedding — the model object of
``/root/reference/src/models/deep_kernel_gp.py:18-188`` reduced to what the sample-path hot path
reads (``src/utils.py:153-206``): ``embedding(X)``, ``train_inputs``, ``train_targets``,
``covar_module.base_kernel.variance``, ``likelihood.noise``, ``mean_module.constant``,
``feature_extractor``, ``architecture``.  botorch/gpytorch-free; training (outside the accelerated
path) is a plain Adam loop on the exact marginal likelihood of the linear-kernel GP.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

DT = torch.float64

_ACT = {"relu": torch.nn.ReLU, "leaky_relu": torch.nn.LeakyReLU, "swish": torch.nn.SiLU,
        "sigmoid": torch.nn.Sigmoid, "tanh": torch.nn.Tanh}


class FFNN(torch.nn.Sequential):
    """Linear + activation stack over ``architecture`` widths; no activation after the last layer."""

    def __init__(self, architecture, activation="leaky_relu"):
        super().__init__()
        self.architecture = list(architecture)
        self.activation = activation
        for i in range(len(architecture) - 1):
            self.add_module(f"linear{i + 1}", torch.nn.Linear(architecture[i], architecture[i + 1]).double())
            if i + 2 < len(architecture):
                self.add_module(f"{activation}{i + 1}", _ACT[activation.lower()]())


class DKGP:
    """``architecture`` = [d_in, ..., h, 1] as in the reference: the network maps d_in -> h and a
    GP with kernel  v * <e, e'>  (+ constant mean, Gaussian noise >= 1e-4) sits on the embedding."""

    def __init__(self, train_X, train_Y, architecture, activation="leaky_relu"):
        if len(architecture) <= 2:
            raise AttributeError("DKGP needs a hidden architecture (reference: deep_kernel_gp.py:100-101)")
        self.architecture = list(architecture)
        self.dkl = True
        self.feature_extractor = FFNN(architecture[:-1], activation)
        self.train_inputs = (train_X.to(DT),)
        self.train_targets = train_Y.to(DT).reshape(-1)
        self.covar_module = SimpleNamespace(
            base_kernel=SimpleNamespace(variance=torch.tensor(1.0, dtype=DT), kind="linear"),
            outputscale=torch.tensor(1.0, dtype=DT))
        self.likelihood = SimpleNamespace(noise=torch.tensor([1e-2], dtype=DT))
        self.mean_module = SimpleNamespace(constant=torch.tensor(0.0, dtype=DT))

    def embedding(self, X):
        net = self.feature_extractor
        p = next(net.parameters())
        return net(X.to(p.device, p.dtype))

    def weight_posterior(self, pending=None, jitter: float = 1e-8):
        """Weight-space posterior of the linear-kernel GP on the embedding (``src/utils.py:153-182``):
        Sn^-1 = I / v + E^T E / sigma^2, mn = Sn E^T (y - mu) / sigma^2.  ``pending`` rows [m, d_in] are
        added as noise-free observations (variance ``jitter * v``), which only changes Sn.
        Returns (mn [h], Sn [h, h]) in float64 on the network's device."""
        with torch.no_grad():
            E = self.embedding(self.train_inputs[0]).to(DT)
            h = E.shape[1]
            v = float(self.covar_module.base_kernel.variance)
            sigma2 = float(self.likelihood.noise.reshape(-1)[0])
            mu = float(self.mean_module.constant)
            prec = torch.eye(h, dtype=DT, device=E.device) / v + E.T @ E / sigma2
            mn = torch.linalg.solve(prec, E.T @ (self.train_targets.to(E.device) - mu)) / sigma2
            if pending is not None and pending.numel() > 0:
                Ep = self.embedding(pending.reshape(-1, self.train_inputs[0].shape[1])).to(DT)
                prec = prec + Ep.T @ Ep / (jitter * v)
            return mn, torch.linalg.inv(prec)

    def posterior(self, X):
        """Latent posterior at X [..., q, d_in]: mean e(X) mn + mu, covariance e(X) Sn e(X)^T
        (what the exact GP with kernel v <e, e'> gives; no outcome transform, as in the reference)."""
        from .gp import Posterior

        mn, Sn = self.weight_posterior()
        with torch.no_grad():
            E = self.embedding(X.reshape(-1, X.shape[-1])).to(DT).reshape(*X.shape[:-1], -1)
        mean = E @ mn + float(self.mean_module.constant)
        return Posterior(mean, E @ Sn @ E.transpose(-1, -2))

    def train_model(self, X, Y, lr=1e-2, num_iter=100):
        """Adam on the negative exact MLL of y ~ N(mu, v E E^T + sigma^2 I) over the network weights,
        log v, log(sigma^2 - 1e-4) and mu."""
        X, Y = X.to(DT), Y.to(DT).reshape(-1)
        n = X.shape[0]
        raw = torch.zeros(3, dtype=DT, requires_grad=True)  # log v, log(noise - 1e-4), mu
        with torch.no_grad():
            raw[1] = math.log(1e-2)
        params = list(self.feature_extractor.parameters()) + [raw]
        opt = torch.optim.Adam(params, lr=lr)
        self.feature_extractor.train()
        for _ in range(num_iter):
            opt.zero_grad()
            E = self.feature_extractor(X)
            v, noise, mu = raw[0].exp(), raw[1].exp() + 1e-4, raw[2]
            K = v * (E @ E.T) + noise * torch.eye(n, dtype=DT)
            Lk = torch.linalg.cholesky(K)
            r = (Y - mu)[:, None]
            loss = 0.5 * (r * torch.cholesky_solve(r, Lk)).sum() + Lk.diagonal().log().sum()
            (loss / n).backward()
            opt.step()
        self.feature_extractor.eval()
        with torch.no_grad():
            self.covar_module.base_kernel.variance = raw[0].exp().detach().clone()
            self.likelihood.noise = (raw[1].exp() + 1e-4).reshape(1).detach().clone()
            self.mean_module.constant = raw[2].detach().clone()
        return None
