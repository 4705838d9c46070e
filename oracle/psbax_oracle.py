"""CPU oracle for the PS-BAX sample-path hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing under
``psbax_b200/`` imports it: the product path is the CUDA library and fails
loudly without it.

What it restates (all ``file:line`` are relative to the read-only reference
checkout ``/root/reference``):

* the sample-path factory ``get_function_samples`` (``src/utils.py:185-250``),
  whose arithmetic lives in the un-vendored dependency ``botorch==0.10.0``
  (``env.yml:21``; ``botorch/utils/gp_sampling.py``: ``RandomFourierFeatures``,
  ``get_weights_posterior``, ``get_gp_samples``) and ``gpytorch==1.11`` kernels;
* ``eval_in_batch`` (``src/algorithms/topk.py:48-49``,
  ``src/algorithms/levelset.py:55-56``);
* the algorithm reductions ``TopK.run_algorithm_on_f``
  (``src/algorithms/topk.py:23-38``), ``LevelSetEstimator.run_algorithm_on_f``
  (``src/algorithms/levelset.py:23-45``) and ``SubsetSelect.run_algorithm_on_f``
  (``src/algorithms/discobax.py:27-65``);
* the PS-BAX pooling loop ``gen_posterior_sampling_batch``
  (``src/acquisition_functions/posterior_sampling.py:46-83``) and the entropy
  pick (``src/acquisition_functions/entropy.py:33-41``).

PARITY STATUS
-------------
* Reduction layer (top-k / level-set / subset-select): PINNED.  The functions
  below are checked against golden vectors produced by importing the
  reference's own ``src/algorithms/*.py`` in the build container
  (``tests/golden/make_golden.py``; botorch's ``PosteriorMean`` is only used
  there for an ``isinstance`` test and is stubbed).
* Sample-path arithmetic (RFF basis, weight posterior, Matheron update):
  PARITY UNPINNED.  botorch/gpytorch are not installed here and cannot be
  installed (no network), the reference ships no golden vectors, and
  ``src/utils.py`` cannot be imported.  The restatement follows the published
  botorch 0.10.0 algorithm and is anchored by
  self-consistency properties in ``tests/test_oracle.py`` (RFF Gram matrix ->
  exact kernel, posterior samples interpolate the data as noise -> 0, Matheron
  and weight-space posteriors agree in mean).

Everything is float64 on the CPU, like the reference (``src/utils.py:26-29``).
"""
from __future__ import annotations

import math
from argparse import Namespace
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

DTYPE = torch.float64

KERNEL_IDS = {"rbf": 0, "matern12": 1, "matern32": 2, "matern52": 3, "linear": 4}
KERNEL_NU = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}


# ----------------------------------------------------------------------------
# Parameter container shared by both sampling schemes
# ----------------------------------------------------------------------------
@dataclass
class PathParams:
    """Everything needed to evaluate S sample paths at arbitrary X.

    f_s(x) = y_mean + y_std * ( mean_const
                                + sum_l W[s,l] * cos(omega[g,l,:] . x + bias[g,l])
                                + sum_j V[s,j] * kappa(|| x*inv_ls - Xtrain[j] ||) )
    with g = 0 when G == 1 (shared basis) and g = s when G == S (one basis per
    sample, the reference's scheme).  ``omega`` is already divided by the
    lengthscale, ``W`` already carries sqrt(2*outputscale/L), ``Xtrain`` is
    already divided by the lengthscale and ``V`` already carries the
    outputscale; ``kappa`` is the unit-variance kernel profile.
    """

    omega: torch.Tensor  # [G, L, d]
    bias: torch.Tensor  # [G, L]
    W: torch.Tensor  # [S, L]
    kernel: str = "matern52"
    Xtrain: Optional[torch.Tensor] = None  # [n, d] (scaled by 1/lengthscale)
    inv_ls: Optional[torch.Tensor] = None  # [d]
    V: Optional[torch.Tensor] = None  # [S, n]
    mean_const: float = 0.0
    y_mean: float = 0.0
    y_std: float = 1.0

    @property
    def S(self) -> int:
        return self.W.shape[0]

    @property
    def G(self) -> int:
        return self.omega.shape[0]

    @property
    def L(self) -> int:
        return self.omega.shape[1]

    @property
    def d(self) -> int:
        return self.omega.shape[2]

    @property
    def n(self) -> int:
        return 0 if self.V is None else self.V.shape[1]

    def cast(self, dtype) -> "PathParams":
        """Round every array through ``dtype`` and come back to float64.

        Parity tests feed the *same* float32-representable numbers to the GPU
        kernels and to this oracle ("identical Omega, b, w_s, v_s and X*")."""

        def r(t):
            return None if t is None else t.to(dtype).to(DTYPE)

        return PathParams(
            omega=r(self.omega), bias=r(self.bias), W=r(self.W), kernel=self.kernel,
            Xtrain=r(self.Xtrain), inv_ls=r(self.inv_ls), V=r(self.V),
            mean_const=float(torch.tensor(self.mean_const).to(dtype)),
            y_mean=float(torch.tensor(self.y_mean).to(dtype)),
            y_std=float(torch.tensor(self.y_std).to(dtype)),
        )


# ----------------------------------------------------------------------------
# Kernels (gpytorch 1.11 RBFKernel / MaternKernel inside ScaleKernel)
# ----------------------------------------------------------------------------
def kernel_profile(r2: torch.Tensor, kernel: str) -> torch.Tensor:
    """Unit-variance stationary kernel as a function of the squared scaled
    distance r2 = ||(x - x') / lengthscale||^2 (gpytorch ``RBFKernel.forward``,
    ``MaternKernel.forward``)."""
    if kernel == "rbf":
        return torch.exp(-0.5 * r2)
    r = torch.sqrt(torch.clamp(r2, min=0.0))
    if kernel == "matern12":
        return torch.exp(-r)
    if kernel == "matern32":
        s = math.sqrt(3.0) * r
        return (1.0 + s) * torch.exp(-s)
    if kernel == "matern52":
        s = math.sqrt(5.0) * r
        return (1.0 + s + (5.0 / 3.0) * r2) * torch.exp(-s)
    raise NotImplementedError(kernel)  # botorch RandomFourierFeatures raises too


def kernel_matrix(X1, X2, lengthscale, outputscale, kernel) -> torch.Tensor:
    """outputscale * kappa(||(x1-x2)/lengthscale||), shapes [a,d],[b,d] -> [a,b]."""
    A = X1.to(DTYPE) / lengthscale
    B = X2.to(DTYPE) / lengthscale
    r2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
    return outputscale * kernel_profile(r2, kernel)


# ----------------------------------------------------------------------------
# Model state read by the path
# ----------------------------------------------------------------------------
@dataclass
class GPState:
    """The attributes of a fitted ``SingleTaskGP`` that ``get_gp_samples`` reads.

    ``train_Y`` is in original units; ``Standardize`` statistics follow botorch
    (``means = Y.mean(-2)``, ``stdvs = Y.std(-2)`` unbiased, replaced by 1.0
    when < 1e-8); ``train_targets`` are the standardized targets the GP sees
    (``src/fit_model.py:39-44``)."""

    train_X: torch.Tensor  # [n, d]
    train_Y: torch.Tensor  # [n]
    lengthscale: torch.Tensor  # [d] (ARD) or [1]
    outputscale: float = 1.0
    noise: float = 1e-4
    kernel: str = "matern52"
    mean_const: float = 0.0
    standardize: bool = True
    y_mean: float = field(init=False, default=0.0)
    y_std: float = field(init=False, default=1.0)

    def __post_init__(self):
        self.train_X = self.train_X.to(DTYPE)
        self.train_Y = self.train_Y.to(DTYPE).reshape(-1)
        self.lengthscale = torch.as_tensor(self.lengthscale, dtype=DTYPE).reshape(-1)
        if self.standardize:
            self.y_mean = float(self.train_Y.mean())
            std = float(self.train_Y.std()) if self.train_Y.numel() > 1 else 1.0
            self.y_std = std if std >= 1e-8 else 1.0

    @property
    def train_targets(self) -> torch.Tensor:
        return (self.train_Y - self.y_mean) / self.y_std

    @property
    def d(self) -> int:
        return self.train_X.shape[1]


# ----------------------------------------------------------------------------
# botorch 0.10.0 gp_sampling, restated
# ----------------------------------------------------------------------------
def rff_basis_draw(d: int, L: int, kernel: str, generator: Optional[torch.Generator] = None):
    """``RandomFourierFeatures.__init__``: weights = randn(d, L); for Matern the
    columns are multiplied by rsqrt(Gamma(nu, nu)) (one draw per feature);
    bias = 2*pi*rand(L).  RNG order randn -> gamma -> rand (call site
    ``src/utils.py:223-228``).  Returns (weights [d, L], bias [L])."""
    weights = torch.randn(d, L, dtype=DTYPE, generator=generator)
    if kernel != "rbf":
        nu = KERNEL_NU[kernel]
        # Gamma(concentration=nu, rate=nu) = standard_gamma(nu) / nu
        g = torch._standard_gamma(torch.full((1, L), nu, dtype=DTYPE), generator=generator) / nu
        weights = torch.rsqrt(g) * weights
    bias = 2.0 * math.pi * torch.rand(L, dtype=DTYPE, generator=generator)
    return weights, bias


def rff_features(X, weights, bias, lengthscale, outputscale) -> torch.Tensor:
    """``RandomFourierFeatures.forward``:
    sqrt(2*outputscale/L) * cos((X / lengthscale) @ weights + bias)."""
    L = weights.shape[1]
    return math.sqrt(2.0 * outputscale / L) * torch.cos((X.to(DTYPE) / lengthscale) @ weights + bias)


def weights_posterior_draw(Phi, y, sigma2, generator=None, return_moments=False):
    """``get_weights_posterior`` + ``.sample()``: A = Phi^T Phi + sigma2 I,
    A^-1 through two triangular solves against I, m = A^-1 Phi^T y,
    w ~ N(m, sigma2 A^-1)."""
    L = Phi.shape[1]
    A = Phi.T @ Phi + sigma2 * torch.eye(L, dtype=DTYPE)
    LA = torch.linalg.cholesky(A)
    eye = torch.eye(L, dtype=DTYPE)
    Ainv = torch.linalg.solve_triangular(
        LA.T, torch.linalg.solve_triangular(LA, eye, upper=False), upper=True)
    m = Ainv @ (Phi.T @ y)
    cov = sigma2 * Ainv
    cov = 0.5 * (cov + cov.T)
    Lc = torch.linalg.cholesky(cov + 1e-12 * torch.eye(L, dtype=DTYPE))
    z = torch.randn(L, dtype=DTYPE, generator=generator)
    w = m + Lc @ z
    if return_moments:
        return w, m, cov
    return w


def draw_reference_paths(gp: GPState, S: int, L: int = 1000, generator=None) -> PathParams:
    """S independent calls of ``get_function_samples(model)`` for a
    ``SingleTaskGP`` (``src/utils.py:222-229``): a fresh basis and a fresh
    weight-posterior draw per sample (``posterior_sampling.py:68-69``), no
    exact-kernel term, GP constant mean ignored, ``Standardize`` un-transform
    applied by the deterministic model's posterior."""
    d = gp.d
    ls = gp.lengthscale.expand(d) if gp.lengthscale.numel() == 1 else gp.lengthscale
    omegas, biases, Ws = [], [], []
    y = gp.train_targets
    for _ in range(S):
        weights, bias = rff_basis_draw(d, L, gp.kernel, generator)
        Phi = rff_features(gp.train_X, weights, bias, ls, gp.outputscale)
        w = weights_posterior_draw(Phi, y, gp.noise, generator)
        omegas.append((weights / ls[:, None]).T.contiguous())  # [L, d]
        biases.append(bias)
        Ws.append(math.sqrt(2.0 * gp.outputscale / L) * w)
    return PathParams(
        omega=torch.stack(omegas), bias=torch.stack(biases), W=torch.stack(Ws),
        kernel=gp.kernel, mean_const=0.0, y_mean=gp.y_mean, y_std=gp.y_std)


def draw_matheron_paths(gp: GPState, S: int, L: int = 1000, generator=None) -> PathParams:
    """North-star form (BASELINE.json): shared basis, prior weights
    w_s ~ N(0, I), exact-kernel update
    v_s = (K + sigma2 I)^-1 (y - mu - Phi w_s - eps_s), eps_s ~ N(0, sigma2 I)."""
    d = gp.d
    ls = gp.lengthscale.expand(d) if gp.lengthscale.numel() == 1 else gp.lengthscale
    weights, bias = rff_basis_draw(d, L, gp.kernel, generator)
    Phi = rff_features(gp.train_X, weights, bias, ls, gp.outputscale)  # [n, L]
    n = gp.train_X.shape[0]
    w = torch.randn(S, L, dtype=DTYPE, generator=generator)
    eps = math.sqrt(gp.noise) * torch.randn(S, n, dtype=DTYPE, generator=generator)
    K = kernel_matrix(gp.train_X, gp.train_X, ls, gp.outputscale, gp.kernel)
    Lk = torch.linalg.cholesky(K + gp.noise * torch.eye(n, dtype=DTYPE))
    resid = gp.train_targets[None, :] - gp.mean_const - w @ Phi.T - eps  # [S, n]
    v = torch.cholesky_solve(resid.T, Lk).T  # [S, n]
    return PathParams(
        omega=(weights / ls[:, None]).T.contiguous()[None], bias=bias[None],
        W=math.sqrt(2.0 * gp.outputscale / L) * w, kernel=gp.kernel,
        Xtrain=gp.train_X / ls, inv_ls=1.0 / ls, V=gp.outputscale * v,
        mean_const=gp.mean_const, y_mean=gp.y_mean, y_std=gp.y_std)


def draw_dkgp_linear_paths(emb_train: torch.Tensor, y: torch.Tensor, variance: float, noise: float,
                           mean_const: float, S: int, generator=None) -> PathParams:
    """``get_function_samples`` DKGP / LinearKernel branch (``src/utils.py:186-206``) with
    ``get_linear_weight_dist`` (``src/utils.py:153-182``):  Sn^-1 = I / v + E^T E / sigma^2,
    mn = Sn^-1 \\ (E^T (y - mu)) / sigma^2,  w ~ N(mn, Sn)  (``MultivariateNormal(precision_matrix=Sn_inv)``);
    f(X) = emb(X) @ w + mu.  The model is fitted on targets standardized by the caller
    (``src/fit_model.py:76-79``), so no outcome transform is applied.  Returned on the EMBEDDING
    space: evaluate with ``eval_paths(p, emb(X))``."""
    E = emb_train.to(DTYPE)
    y = y.to(DTYPE).reshape(-1)
    h = E.shape[1]
    Sn_inv = torch.eye(h, dtype=DTYPE) / variance + E.T @ E / noise
    mn = torch.linalg.solve(Sn_inv, E.T @ (y - mean_const)) / noise
    Sn = torch.linalg.inv(Sn_inv)
    Lc = torch.linalg.cholesky(0.5 * (Sn + Sn.T))
    W = mn[None, :] + torch.randn(S, h, dtype=DTYPE, generator=generator) @ Lc.T
    return PathParams(omega=torch.zeros(1, 0, h, dtype=DTYPE), bias=torch.zeros(1, 0, dtype=DTYPE),
                      W=torch.zeros(S, 0, dtype=DTYPE), kernel="linear", Xtrain=torch.eye(h, dtype=DTYPE),
                      inv_ls=torch.ones(h, dtype=DTYPE), V=W, mean_const=mean_const, y_mean=0.0, y_std=1.0)


def posterior_mean_params(gp: GPState) -> PathParams:
    """``PosteriorMean(model)`` as the S=1, L=0 special case:
    mu(x) = mean_const + k(x, X) (K + sigma2 I)^-1 (y - mean_const)
    (used by ``src/performance_metrics.py:117-122``)."""
    d = gp.d
    ls = gp.lengthscale.expand(d) if gp.lengthscale.numel() == 1 else gp.lengthscale
    n = gp.train_X.shape[0]
    K = kernel_matrix(gp.train_X, gp.train_X, ls, gp.outputscale, gp.kernel)
    Lk = torch.linalg.cholesky(K + gp.noise * torch.eye(n, dtype=DTYPE))
    alpha = torch.cholesky_solve((gp.train_targets - gp.mean_const)[:, None], Lk)[:, 0]
    return PathParams(
        omega=torch.zeros(1, 0, d, dtype=DTYPE), bias=torch.zeros(1, 0, dtype=DTYPE),
        W=torch.zeros(1, 0, dtype=DTYPE), kernel=gp.kernel, Xtrain=gp.train_X / ls,
        inv_ls=1.0 / ls, V=(gp.outputscale * alpha)[None], mean_const=gp.mean_const,
        y_mean=gp.y_mean, y_std=gp.y_std)


# ----------------------------------------------------------------------------
# Evaluation
# ----------------------------------------------------------------------------
def eval_paths(p: PathParams, X: torch.Tensor, chunk: int = 8192) -> torch.Tensor:
    """All S sample paths at X [N, d] -> [N, S], float64."""
    X = X.to(DTYPE)
    N = X.shape[0]
    out = torch.empty(N, p.S, dtype=DTYPE)
    for a in range(0, N, chunk):
        Xc = X[a:a + chunk]
        if p.L > 0:
            if p.G == 1:
                Phi = torch.cos(Xc @ p.omega[0].T + p.bias[0])  # [c, L]
                f = Phi @ p.W.T
            else:
                # [c, S, L]
                arg = torch.einsum("cd,sld->csl", Xc, p.omega) + p.bias[None]
                f = (torch.cos(arg) * p.W[None]).sum(-1)
        else:
            f = torch.zeros(Xc.shape[0], p.S, dtype=DTYPE)
        if p.n > 0:
            Z = Xc * p.inv_ls
            if p.kernel == "linear":  # gpytorch LinearKernel: k(x, x') = x . x'
                f = f + (Z @ p.Xtrain.T) @ p.V.T
            else:
                r2 = ((Z[:, None, :] - p.Xtrain[None, :, :]) ** 2).sum(-1)
                f = f + kernel_profile(r2, p.kernel) @ p.V.T
        out[a:a + chunk] = p.y_mean + p.y_std * (p.mean_const + f)
    return out


def eval_paths_best_effort(p: PathParams, X: torch.Tensor, block: int = 65536) -> torch.Tensor:
    """The "best-effort CPU" form of the same path: float32, shared
    basis, all S paths at once, one blocked GEMM pair per ``block`` rows on all host threads —
    what a careful CPU implementation would do, so that the GPU / CPU ratio is not just the
    reference's Python overhead.  ``p`` must be a shared-basis (G == 1) parameter set.  [N, S] float32."""
    assert p.G == 1, "best-effort CPU line is the shared-basis form"
    f32 = torch.float32
    X = X.to(f32)
    om = p.omega[0].to(f32).T.contiguous() if p.L > 0 else None  # [d, L]
    b = p.bias[0].to(f32) if p.L > 0 else None
    Wt = p.W.to(f32).T.contiguous() if p.L > 0 else None          # [L, S]
    Xt = p.Xtrain.to(f32) if p.n > 0 else None
    Vt = p.V.to(f32).T.contiguous() if p.n > 0 else None          # [n, S]
    il = p.inv_ls.to(f32) if p.n > 0 else None
    xt2 = (Xt * Xt).sum(-1) if p.n > 0 else None
    out = torch.empty(X.shape[0], p.S, dtype=f32)
    for a in range(0, X.shape[0], block):
        Xc = X[a:a + block]
        f = torch.zeros(Xc.shape[0], p.S, dtype=f32)
        if p.L > 0:
            Phi = torch.addmm(b, Xc, om)
            Phi.cos_()
            f = Phi @ Wt
        if p.n > 0:
            Z = Xc * il
            if p.kernel == "linear":
                f = f + (Z @ Xt.T) @ Vt
            else:  # GEMM form of the squared distances (what a BLAS-bound CPU code would use)
                r2 = ((Z * Z).sum(-1, keepdim=True) + xt2[None, :] - 2.0 * (Z @ Xt.T)).clamp_min_(0.0)
                f = f + kernel_profile(r2, p.kernel) @ Vt
        out[a:a + block] = p.y_mean + p.y_std * (p.mean_const + f)
    return out


def eval_path_grad(p: PathParams, X: torch.Tensor, s: int) -> torch.Tensor:
    """d f_s / d x at X [N, d] -> [N, d] (what ``LBFGSB`` obtains by autograd,
    ``src/algorithms/lbfgsb.py:29-39``)."""
    X = X.to(DTYPE).clone().requires_grad_(True)
    sub = PathParams(
        omega=p.omega[(0 if p.G == 1 else s)][None], bias=p.bias[(0 if p.G == 1 else s)][None],
        W=p.W[s][None], kernel=p.kernel, Xtrain=p.Xtrain, inv_ls=p.inv_ls,
        V=None if p.V is None else p.V[s][None], mean_const=p.mean_const,
        y_mean=p.y_mean, y_std=p.y_std)
    Xc = X
    f = torch.cos(Xc @ sub.omega[0].T + sub.bias[0]) @ sub.W[0] if sub.L > 0 else 0.0 * Xc.sum(-1)
    if sub.n > 0:
        Z = Xc * sub.inv_ls
        if sub.kernel == "linear":
            f = f + (Z @ sub.Xtrain.T) @ sub.V[0]
        else:
            r2 = ((Z[:, None, :] - sub.Xtrain[None, :, :]) ** 2).sum(-1)
            f = f + kernel_profile(r2 + 1e-300, sub.kernel) @ sub.V[0]
    y = sub.y_mean + sub.y_std * (sub.mean_const + f)
    (g,) = torch.autograd.grad(y.sum(), X)
    return g


def path_scale(p: PathParams, X: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-sample magnitude THE 1e-4 PARITY TOLERANCE is relative to (SURVEY.md section 8c:
    "y_std * sqrt(outputscale) * ||w_s|| / sqrt(L)-like path scale"): the magnitude of the path,

        y_std * max(1, ||W_s||_2)              (W already carries sqrt(2 * outputscale / L)).

    There is deliberately NO ||V_s|| term: the Matheron update weights V = (K + sigma^2 I)^-1 (...)
    are large and cancel in k(x, X) . V_s (||V_s|| ~ 80 on the bench problem against a path of
    magnitude ~1), so a scale that contains them is far looser than "1e-4 relative" — see
    ``forward_error_scale`` for that diagnostic.  Linear-kernel rows (the DKGP weight-space form,
    ``src/utils.py:153-206``) are different: there V *is* the weight vector w_s on features
    x . x' that are not bounded by 1, so its norm times the largest |feature| over the candidates
    X is the path magnitude."""
    s = p.W.norm(dim=1) if p.L > 0 else torch.zeros(p.S, dtype=DTYPE)
    if p.n > 0 and p.kernel == "linear":
        m = 1.0
        if X is not None:
            m = max(1.0, float(((X.to(DTYPE) * p.inv_ls) @ p.Xtrain.T).abs().max()))
        s = s + m * p.V.norm(dim=1)
    return abs(p.y_std) * torch.clamp(s, min=1.0)


def forward_error_scale(p: PathParams, X: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DIAGNOSTIC ONLY (not the parity criterion): y_std * max(1, ||W_s||_2 + m ||V_s||_2), the
    forward-error scale of the two dot products.  Round 1 used this as the tolerance scale; on
    Matheron draws it is 20-90x looser than the path magnitude."""
    s = p.W.norm(dim=1) if p.L > 0 else torch.zeros(p.S, dtype=DTYPE)
    if p.n > 0:
        m = 1.0
        if p.kernel == "linear" and X is not None:
            m = max(1.0, float(((X.to(DTYPE) * p.inv_ls) @ p.Xtrain.T).abs().max()))
        s = s + m * p.V.norm(dim=1)
    return abs(p.y_std) * torch.clamp(s, min=1.0)


def make_callable(p: PathParams, s: int) -> Callable[[torch.Tensor], torch.Tensor]:
    """The object ``get_function_samples`` returns for a SingleTaskGP:
    ``PosteriorMean(deterministic_model)``; callers pass X [b, 1, d] and get [b]
    (``src/algorithms/topk.py:25-26``)."""
    sub = PathParams(
        omega=p.omega[(0 if p.G == 1 else s)][None], bias=p.bias[(0 if p.G == 1 else s)][None],
        W=p.W[s][None], kernel=p.kernel, Xtrain=p.Xtrain, inv_ls=p.inv_ls,
        V=None if p.V is None else p.V[s][None], mean_const=p.mean_const,
        y_mean=p.y_mean, y_std=p.y_std)

    def f(X):
        X = torch.as_tensor(X, dtype=DTYPE)
        if X.dim() == 3:  # [b, 1, d] -> [b]
            return eval_paths(sub, X.squeeze(1))[:, 0]
        return eval_paths(sub, X.reshape(-1, p.d))[:, 0]

    return f


def eval_in_batch(f, x_set, max_batch_size: int = 100):
    """``src/algorithms/topk.py:48-49``: cat([f(X_) for X_ in x_set.split(100)])."""
    return torch.cat([f(X_) for X_ in x_set.split(max_batch_size)])


# ----------------------------------------------------------------------------
# Algorithm reductions (pinned by tests/golden/algorithms_golden.npz)
# ----------------------------------------------------------------------------
def topk_reduce(fx: torch.Tensor, k: int):
    """``TopK.run_algorithm_on_f`` reduction (``src/algorithms/topk.py:31-34``):
    indices of the k largest values, largest first.  ``torch.topk`` leaves the
    order of exact ties unspecified; the oracle (and the CUDA path) break ties
    towards the smaller index."""
    fx = torch.as_tensor(fx).reshape(-1)
    order = np.lexsort((np.arange(fx.numel()), -fx.numpy()))  # by -value, then index
    idx = torch.from_numpy(order[:k].copy())
    return idx, fx[idx]


def levelset_reduce(fx, threshold: float):
    """``LevelSetEstimator.run_algorithm_on_f`` reduction
    (``src/algorithms/levelset.py:34-39``): idx = where(fx > threshold); if
    empty -> [argmax fx]."""
    fx = np.asarray(fx).reshape(-1)
    idx = np.where(fx > threshold)[0]
    if len(idx) == 0:
        idx = np.array([int(np.argmax(fx))])
    return idx


def subset_select_reduce(fx, etas, k: int, nonneg: bool = False, multiplicative: bool = False):
    """``SubsetSelect.run_algorithm_on_f`` (``src/algorithms/discobax.py:37-56``)
    with the objective's noise model (``src/problems.py:146-163``): values = fx + etas [B,N], or, for
    ``multiplicative``, values = fx * etas with etas the 0 / 1 Bernoulli masks (``:150-160``);
    first pick = argmax of the column means; then k-1 greedy picks maximising
    mean_b max(cur_max[b], values[b, j]) with already-chosen j scored 0.  The
    reference breaks exact ties at random (``discobax.py:79-83``); the oracle
    takes the smallest index."""
    fx = np.asarray(fx, dtype=np.float64).reshape(-1)
    etas = np.asarray(etas, dtype=np.float64)
    values = fx[None, :] * etas if multiplicative else fx[None, :] + etas
    if nonneg:
        values = np.maximum(0.0, values)
    mean_values = values.mean(axis=0)
    idxes = [int(np.argmax(mean_values))]
    cur = values[:, idxes[0]].copy()
    for _ in range(k - 1):
        e_vals = np.maximum(cur[:, None], values).mean(axis=0)
        e_vals[idxes] = 0.0
        j = int(np.argmax(e_vals))
        idxes.append(j)
        cur = np.maximum(cur, values[:, j])
    return idxes, values


# ----------------------------------------------------------------------------
# Exact GP posterior + PS-BAX acquisition (rows a10/a11 of SURVEY.md section 8)
# ----------------------------------------------------------------------------
def gp_posterior(gp: GPState, X: torch.Tensor):
    """Exact posterior mean / covariance of the *standardized* GP at X [q, d]
    and its un-standardized version (botorch ``SingleTaskGP.posterior`` with
    ``Standardize``): returns (mean [q], cov [q,q]) in original units."""
    d = gp.d
    ls = gp.lengthscale.expand(d) if gp.lengthscale.numel() == 1 else gp.lengthscale
    n = gp.train_X.shape[0]
    K = kernel_matrix(gp.train_X, gp.train_X, ls, gp.outputscale, gp.kernel)
    Lk = torch.linalg.cholesky(K + gp.noise * torch.eye(n, dtype=DTYPE))
    Ks = kernel_matrix(X, gp.train_X, ls, gp.outputscale, gp.kernel)  # [q, n]
    Kss = kernel_matrix(X, X, ls, gp.outputscale, gp.kernel)
    alpha = torch.cholesky_solve((gp.train_targets - gp.mean_const)[:, None], Lk)[:, 0]
    mean = gp.mean_const + Ks @ alpha
    A = torch.linalg.solve_triangular(Lk, Ks.T, upper=False)  # [n, q]
    cov = Kss - A.T @ A
    return gp.y_mean + gp.y_std * mean, (gp.y_std ** 2) * cov


def entropy_acqf(gp: GPState, X: torch.Tensor) -> torch.Tensor:
    """``EntropyAcquisitionFunction.forward`` (``entropy.py:33-41``), X [b,q,d] -> [b]."""
    out = []
    for Xb in X:
        _, cov = gp_posterior(gp, Xb)
        out.append(0.5 * math.log(2 * math.pi) + 0.5 * torch.logdet(cov))
    return torch.stack(out)


def condition_on_observations(gp: GPState, X_new: torch.Tensor, Y_new: torch.Tensor) -> GPState:
    """``BAXAcquisitionFunction.fit_with_old_model`` (``bax_acquisition.py:177-203``) -> botorch
    ``SingleTaskGP.condition_on_observations``: the same hyper-parameters and the SAME (already
    fitted) ``Standardize`` statistics; the new rows are observed with the likelihood noise."""
    new = GPState(train_X=torch.cat([gp.train_X, X_new.to(DTYPE).reshape(-1, gp.d)]),
                  train_Y=torch.cat([gp.train_Y, Y_new.to(DTYPE).reshape(-1)]), lengthscale=gp.lengthscale,
                  outputscale=gp.outputscale, noise=gp.noise, kernel=gp.kernel, mean_const=gp.mean_const,
                  standardize=False)
    new.y_mean, new.y_std = gp.y_mean, gp.y_std
    return new


def bax_eig(gp: GPState, exe_paths, X: torch.Tensor, pending: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``BAXAcquisitionFunction.forward`` with ``acq_str="exe"`` (``bax_acquisition.py:87-122``,
    ``acq_exe_normal`` ``:156-174``, ``entropy`` ``:142-152``): X [b, q, d] -> [b],

        H[f(X)] - mean_s H[f(X) | execution path s],   H = 0.5 log(2 pi) + 0.5 logdet Cov_post(X),

    the conditional models being ``condition_on_observations`` of the current one on each path's
    (x, y) (``initialize`` ``:236-247``).  ``exe_paths`` is a list of (x [k, d], y [k]) pairs;
    ``pending`` rows are appended to every q-batch (``@concatenate_pending_points``)."""
    X = X.to(DTYPE)
    if pending is not None and pending.numel() > 0:
        X = torch.cat([X, pending.to(DTYPE).reshape(1, -1, X.shape[-1]).expand(X.shape[0], -1, -1)], dim=-2)
    models = [condition_on_observations(gp, px, py) for px, py in exe_paths]

    def entropy(g, Xb):
        _, cov = gp_posterior(g, Xb)
        return 0.5 * math.log(2 * math.pi) + 0.5 * torch.logdet(cov)

    out = []
    for Xb in X:
        h_post = entropy(gp, Xb)
        h_samp = torch.stack([entropy(m, Xb) for m in models])
        out.append(h_post - h_samp.mean())
    return torch.stack(out)


def optimize_acqf_discrete(gp: GPState, choices: torch.Tensor, q: int):
    """botorch 0.10 ``optimize_acqf_discrete(unique=True)`` restated for the
    entropy acquisition: greedy over q, the chosen row joins X_pending and is
    removed from the choices (``posterior_sampling.py:78-81``)."""
    choices = choices.to(DTYPE)
    remaining = list(range(choices.shape[0]))
    picked: List[int] = []
    for _ in range(q):
        pend = choices[picked]
        vals = entropy_acqf(gp, torch.stack(
            [torch.cat([choices[i][None], pend]) for i in remaining]))
        j = int(torch.argmax(vals))
        picked.append(remaining.pop(j))
    return choices[picked], picked


def gen_posterior_sampling_batch(gp: GPState, x_set: torch.Tensor, algo: str, batch_size: int,
                                 k: int = 10, threshold: float = 0.0, L: int = 1000,
                                 generator=None):
    """``gen_posterior_sampling_batch`` non-SubsetSelect branch
    (``posterior_sampling.py:66-83``): ``batch_size`` reference-style samples,
    each run through the algorithm in chunks of 100, outputs pooled, then the
    max-entropy pick."""
    pool = []
    x_set = x_set.to(DTYPE)
    for _ in range(batch_size):
        p = draw_reference_paths(gp, 1, L, generator)
        f = make_callable(p, 0)
        fx = eval_in_batch(f, x_set.unsqueeze(1)).detach()
        if algo == "TopK":
            idx, _ = topk_reduce(fx, k)
        else:
            idx = torch.from_numpy(levelset_reduce(fx.numpy(), threshold))
        pool.extend(x_set[i] for i in idx.tolist())
    x_batch = torch.stack(pool)
    return optimize_acqf_discrete(gp, x_batch, batch_size)[0]


# ----------------------------------------------------------------------------
# Grid helpers used to build config inputs (src/utils.py:264-286)
# ----------------------------------------------------------------------------
def get_mesh(dim: int, steps: int):
    axes = [torch.linspace(0, 1, steps, dtype=DTYPE) for _ in range(dim)]
    return torch.meshgrid(*axes, indexing="ij")


def reshape_mesh(xx):
    return torch.hstack([x.reshape(-1, 1) for x in xx])
