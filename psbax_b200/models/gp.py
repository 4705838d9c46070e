"""Minimal botorch-free exact GP carrying the attributes the sample-path hot path reads.

botorch / gpytorch are not installable in this environment, so the model object
the reference builds in ``src/fit_model.py:20-70`` (``SingleTaskGP`` +
``Standardize`` with a ``ScaleKernel(Matern-5/2 | RBF)``) is restated here with
the attribute surface listed in SURVEY.md Appendix C:

    model.train_inputs[0], model.train_targets, model.covar_module.outputscale,
    model.covar_module.base_kernel.{lengthscale, nu}, model.likelihood.noise,
    model.mean_module.constant, model.outcome_transform.{means, stdvs},
    model.posterior(X).{mean, variance, covariance_matrix}

Hyper-parameter fitting is outside the accelerated path; ``fit`` is a small
L-BFGS MAP fit with botorch 0.10's default priors so that demo-sized problems
behave like the reference.  Everything here is float64 torch, CPU or CUDA.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import torch

DT = torch.float64


def kernel_profile(r2: torch.Tensor, kernel: str) -> torch.Tensor:
    if kernel == "rbf":
        return torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp(r2, min=1e-30))
    if kernel == "matern12":
        return torch.exp(-r)
    if kernel == "matern32":
        s = math.sqrt(3.0) * r
        return (1.0 + s) * torch.exp(-s)
    if kernel == "matern52":
        s = math.sqrt(5.0) * r
        return (1.0 + s + (5.0 / 3.0) * r2) * torch.exp(-s)
    raise NotImplementedError(f"kernel {kernel!r}: only RBF and Matern have an RFF basis")


class Standardize:
    """botorch ``Standardize(m=1)``: means / stdvs of train_Y (unbiased std, 1.0 if < 1e-8)."""

    def __init__(self, Y: torch.Tensor):
        self.means = Y.mean(dim=-2, keepdim=True)
        std = Y.std(dim=-2, keepdim=True) if Y.shape[-2] > 1 else torch.ones_like(self.means)
        self.stdvs = torch.where(std >= 1e-8, std, torch.ones_like(std))

    def __call__(self, Y):
        return (Y - self.means) / self.stdvs

    def untransform(self, Y):
        return self.means + self.stdvs * Y


class _BaseKernel:
    def __init__(self, kind: str, lengthscale: torch.Tensor):
        self.kind = kind
        self.lengthscale = lengthscale  # [1, d] (ARD) or [1, 1]
        self.nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}.get(kind)


class _ScaleKernel:
    def __init__(self, base: _BaseKernel, outputscale: torch.Tensor):
        self.base_kernel = base
        self.outputscale = outputscale


class Posterior:
    def __init__(self, mean, cov):
        self.mean = mean.unsqueeze(-1)  # [..., q, 1]
        self.covariance_matrix = cov  # [..., q, q]

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-2, dim2=-1).unsqueeze(-1)

    @property
    def stddev(self):
        return self.variance.clamp_min(0).sqrt()


class SingleTaskGP:
    """Exact GP, constant mean, Gaussian noise, ``Standardize`` outcome transform."""

    def __init__(self, train_X: torch.Tensor, train_Y: torch.Tensor, kernel: str = "matern52",
                 lengthscale=None, outputscale: float = 1.0, noise: float = 1e-4,
                 mean_const: float = 0.0, ard: bool = True, standardize: bool = True):
        X = train_X.to(DT)
        Y = train_Y.to(DT).reshape(X.shape[0], -1)
        assert Y.shape[1] == 1, "single-output model"
        self.train_inputs = (X,)
        self._Y_raw = Y
        self.outcome_transform = Standardize(Y) if standardize else None
        Ys = self.outcome_transform(Y) if standardize else Y
        self.train_targets = Ys[:, 0]
        d = X.shape[1]
        if lengthscale is None:
            lengthscale = torch.full((1, d if ard else 1), 0.5, dtype=DT)
        ls = torch.as_tensor(lengthscale, dtype=DT).reshape(1, -1).to(X.device)
        self.covar_module = _ScaleKernel(_BaseKernel(kernel, ls),
                                         torch.tensor(float(outputscale), dtype=DT, device=X.device))
        self.likelihood = SimpleNamespace(noise=torch.tensor([float(noise)], dtype=DT, device=X.device))
        self.mean_module = SimpleNamespace(constant=torch.tensor(float(mean_const), dtype=DT, device=X.device))
        self._chol = None

    # -- convenience ---------------------------------------------------------------
    @property
    def kernel(self) -> str:
        return self.covar_module.base_kernel.kind

    @property
    def device(self):
        return self.train_inputs[0].device

    def lengthscale_vector(self) -> torch.Tensor:
        d = self.train_inputs[0].shape[1]
        ls = self.covar_module.base_kernel.lengthscale.reshape(-1)
        return ls.expand(d) if ls.numel() == 1 else ls

    def y_mean_std(self):
        if getattr(self, "outcome_transform", None) is None:
            return 0.0, 1.0
        return float(self.outcome_transform.means), float(self.outcome_transform.stdvs)

    def kernel_matrix(self, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
        ls = self.lengthscale_vector()
        a, b = A.to(DT) / ls, B.to(DT) / ls
        r2 = ((a.unsqueeze(-2) - b.unsqueeze(-3)) ** 2).sum(-1)
        return self.covar_module.outputscale * kernel_profile(r2, self.kernel)

    def train_cholesky(self) -> torch.Tensor:
        if self._chol is None:
            X = self.train_inputs[0]
            K = self.kernel_matrix(X, X)
            K = K + self.likelihood.noise * torch.eye(X.shape[0], dtype=DT, device=X.device)
            self._chol = torch.linalg.cholesky(K)
        return self._chol

    def invalidate(self):
        self._chol = None

    # -- posterior -------------------------------------------------------------------
    def posterior(self, X: torch.Tensor) -> Posterior:
        """X [..., q, d] -> joint posterior over the q points, original units."""
        X = X.to(DT).to(self.device)
        Xt = self.train_inputs[0]
        Lk = self.train_cholesky()
        mu = self.mean_module.constant
        Ks = self.kernel_matrix(X, Xt)  # [..., q, n]
        alpha = torch.cholesky_solve((self.train_targets - mu).unsqueeze(-1), Lk)[:, 0]
        mean = mu + Ks @ alpha
        A = torch.linalg.solve_triangular(Lk, Ks.transpose(-1, -2), upper=False)
        cov = self.kernel_matrix(X, X) - A.transpose(-1, -2) @ A
        if self.outcome_transform is not None:
            m, s = self.outcome_transform.means.reshape(()), self.outcome_transform.stdvs.reshape(())
            mean, cov = m + s * mean, (s ** 2) * cov
        return Posterior(mean, cov)

    # -- MAP fit (outside the accelerated path) ----------------------------------------
    def fit(self, max_iter: int = 75) -> "SingleTaskGP":
        """L-BFGS on the exact MLL + botorch 0.10 default priors
        (lengthscale Gamma(3, 6), outputscale Gamma(2, 0.15), noise Gamma(1.1, 0.05), noise >= 1e-4)."""
        X, y = self.train_inputs[0], self.train_targets
        n, d = X.shape
        nls = self.covar_module.base_kernel.lengthscale.numel()
        raw = torch.zeros(nls + 3, dtype=DT, device=X.device)
        raw[:nls] = self.covar_module.base_kernel.lengthscale.reshape(-1).log()
        raw[nls] = self.covar_module.outputscale.log()
        raw[nls + 1] = (self.likelihood.noise[0] - 1e-4).clamp_min(1e-6).log()
        raw[nls + 2] = self.mean_module.constant
        raw.requires_grad_(True)
        opt = torch.optim.LBFGS([raw], max_iter=max_iter, line_search_fn="strong_wolfe")

        def gamma_lp(x, a, b):
            return (a - 1) * torch.log(x) - b * x

        def closure():
            opt.zero_grad()
            ls, os_, noise = raw[:nls].exp(), raw[nls].exp(), raw[nls + 1].exp() + 1e-4
            lsv = ls.expand(d) if nls == 1 else ls
            a = X / lsv
            r2 = ((a[:, None, :] - a[None, :, :]) ** 2).sum(-1)
            K = os_ * kernel_profile(r2, self.kernel) + noise * torch.eye(n, dtype=DT, device=X.device)
            try:
                Lk = torch.linalg.cholesky(K)
            except Exception:
                Lk = torch.linalg.cholesky(K + 1e-6 * torch.eye(n, dtype=DT, device=X.device))
            r = (y - raw[nls + 2]).unsqueeze(-1)
            sol = torch.cholesky_solve(r, Lk)
            mll = -0.5 * (r * sol).sum() - Lk.diagonal().log().sum() - 0.5 * n * math.log(2 * math.pi)
            lp = gamma_lp(ls, 3.0, 6.0).sum() + gamma_lp(os_, 2.0, 0.15) + gamma_lp(noise, 1.1, 0.05)
            loss = -(mll + lp) / n
            loss.backward()
            return loss

        try:
            opt.step(closure)
        except Exception:
            pass
        with torch.no_grad():
            if torch.isfinite(raw).all():
                self.covar_module.base_kernel.lengthscale = raw[:nls].exp().reshape(1, -1).detach().clone()
                self.covar_module.outputscale = raw[nls].exp().detach().clone()
                self.likelihood.noise = (raw[nls + 1].exp() + 1e-4).reshape(1).detach().clone()
                self.mean_module.constant = raw[nls + 2].detach().clone()
        self.invalidate()
        return self


class ModelListGP:
    """Independent single-output GPs over the same inputs (botorch ``ModelListGP`` as
    ``src/fit_model.py:59-62`` builds it): ``.models``, per-model ``train_inputs`` / ``train_targets``
    lists, and a joint ``posterior`` whose mean is [..., q, m] (block-diagonal across outputs)."""

    def __init__(self, *models):
        if not models:
            raise ValueError("ModelListGP needs at least one model")
        self.models = list(models)

    @property
    def num_outputs(self) -> int:
        return len(self.models)

    @property
    def train_inputs(self):
        return [m.train_inputs for m in self.models]

    @property
    def train_targets(self):
        return [m.train_targets for m in self.models]

    def posterior(self, X: torch.Tensor):
        posts = [m.posterior(X) for m in self.models]
        covs = [p.covariance_matrix for p in posts]
        q, m = covs[0].shape[-1], len(covs)
        cov = covs[0].new_zeros(*covs[0].shape[:-2], m * q, m * q)  # outputs are independent: block diagonal
        for j, c in enumerate(covs):
            cov[..., j * q:(j + 1) * q, j * q:(j + 1) * q] = c
        return SimpleNamespace(mean=torch.cat([p.mean for p in posts], dim=-1),  # [..., q, m]
                               variance=torch.cat([p.variance for p in posts], dim=-1), covariance_matrix=cov)
