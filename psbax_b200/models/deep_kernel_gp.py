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

    def __init__(self, train_X, train_Y, architecture, activation="leaky_relu", kernel="linear"):
        """``kernel``: "linear" (what the reference hard-wires, ``deep_kernel_gp.py:112-116``) or
        "rbf" / "matern12|32|52" on the embedding (the commented-out alternative there, which is what
        the non-linear branch of ``get_function_samples``, ``src/utils.py:207-221``, serves)."""
        if len(architecture) <= 2:
            raise AttributeError("DKGP needs a hidden architecture (reference: deep_kernel_gp.py:100-101)")
        self.architecture = list(architecture)
        self.dkl = True
        self.feature_extractor = FFNN(architecture[:-1], activation)
        self.train_inputs = (train_X.to(DT),)
        self.train_targets = train_Y.to(DT).reshape(-1)
        h = architecture[-2]
        self.covar_module = SimpleNamespace(
            base_kernel=SimpleNamespace(variance=torch.tensor(1.0, dtype=DT), kind=kernel,
                                        lengthscale=torch.ones(1, h, dtype=DT),
                                        nu={"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}.get(kernel)),
            outputscale=torch.tensor(1.0, dtype=DT))
        self.likelihood = SimpleNamespace(noise=torch.tensor([1e-2], dtype=DT))
        self.mean_module = SimpleNamespace(constant=torch.tensor(0.0, dtype=DT))

    def embedding(self, X):
        net = self.feature_extractor
        p = next(net.parameters())
        return net(X.to(p.device, p.dtype))

    @property
    def kernel(self) -> str:
        return self.covar_module.base_kernel.kind

    def _gram(self, E, v, ls=None):
        """Kernel matrix of the embedded rows: v <e, e'> or v kappa(|e - e'| / ls)."""
        if self.kernel == "linear":
            return v * (E @ E.T)
        from .gp import kernel_profile

        a = E / (self.covar_module.base_kernel.lengthscale.reshape(-1).to(E) if ls is None else ls)
        return v * kernel_profile(((a[:, None, :] - a[None, :, :]) ** 2).sum(-1), self.kernel)

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

        if self.kernel != "linear":
            raise NotImplementedError("posterior() of the stand-in is the weight-space form of the linear kernel")
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
        h = self.architecture[-2]
        raw_ls = torch.zeros(h, dtype=DT, requires_grad=self.kernel != "linear")  # log lengthscale (non-linear kernels)
        with torch.no_grad():
            raw[1] = math.log(1e-2)
        params = list(self.feature_extractor.parameters()) + [raw] + ([raw_ls] if self.kernel != "linear" else [])
        opt = torch.optim.Adam(params, lr=lr)
        self.feature_extractor.train()
        for _ in range(num_iter):
            opt.zero_grad()
            E = self.feature_extractor(X)
            v, noise, mu = raw[0].exp(), raw[1].exp() + 1e-4, raw[2]
            K = self._gram(E, v, raw_ls.exp()) + noise * torch.eye(n, dtype=DT)
            Lk = torch.linalg.cholesky(K)
            r = (Y - mu)[:, None]
            loss = 0.5 * (r * torch.cholesky_solve(r, Lk)).sum() + Lk.diagonal().log().sum()
            (loss / n).backward()
            opt.step()
        self.feature_extractor.eval()
        with torch.no_grad():
            self.covar_module.base_kernel.variance = raw[0].exp().detach().clone()
            if self.kernel != "linear":  # ScaleKernel(Matern / RBF): the scale is the outputscale
                self.covar_module.outputscale = raw[0].exp().detach().clone()
                self.covar_module.base_kernel.lengthscale = raw_ls.exp().detach().clone().reshape(1, -1)
            self.likelihood.noise = (raw[1].exp() + 1e-4).reshape(1).detach().clone()
            self.mean_module.constant = raw[2].detach().clone()
        return None
