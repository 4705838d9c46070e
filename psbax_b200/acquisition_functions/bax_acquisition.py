"""INFO-BAX acquisition: the second caller of the sample-path hot path
(``/root/reference/src/acquisition_functions/bax_acquisition.py:45-85, 221-277``).

``initialize`` / ``run_algorithm_on_f_list`` are on the accelerated path: the reference draws
``exe_paths`` (default 30) sample paths one by one and deep-copies the algorithm (including X*) per
path; here the paths are drawn as one batch and the algorithm runs on all of them in one fused
GPU launch.  ``forward`` (the expected-information-gain estimate, ``bax_acquisition.py:87-174``):
H[y(X)] - mean_s H[y(X) | execution path s] with exact GP conditioning under fixed hyper-parameters
and fixed standardization, in plain torch for any q; ``q1_values`` is the same number for every row
of a candidate matrix through S + 1 fused ``psbax_eval_sumsq`` launches (SURVEY.md section 8f, row
f3), which ``optimize_acqf_discrete`` uses for its greedy steps.
"""
import math

import numpy as np
import torch

from ..sample_paths import SamplePathBatch
from ..utils import get_function_sample_batch


class BAXAcquisitionFunction:
    def __init__(self, model, algo, **kwargs):
        self.model = model
        self.algorithm = algo
        for k, v in {"batch_size": 1, "num_rff_features": 1000, "crop": True, "acq_str": "exe"}.items():
            kwargs.setdefault(k, v)
        self.n_samples = kwargs.get("exe_paths", 30)
        for k, v in kwargs.items():  # the reference sets every kwarg as an attribute (:84-85)
            setattr(self, k, v)
        self.X_pending = None

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending

    # -- accelerated part ------------------------------------------------------------------
    def run_algorithm_on_f_list(self, f_sample_list):
        """``f_sample_list``: a ``SamplePathBatch`` (one fused launch) or a list of callables
        (the reference protocol: one algorithm copy per path)."""
        self.algorithm.initialize()
        if isinstance(f_sample_list, SamplePathBatch):
            results = self.algorithm.get_copy().run_algorithm_on_batch(f_sample_list)
        else:
            results = [self.algorithm.get_copy().run_algorithm_on_f(f) for f in f_sample_list]
        return [r[0] for r in results], [r[1] for r in results]

    def initialize(self, **kwargs):
        kwargs.pop("batch_size", None)
        keep = {k: kwargs[k] for k in ("sampler", "num_rff_features", "generator", "device", "engine") if k in kwargs}
        paths = get_function_sample_batch(self.model, self.n_samples, **keep)
        self.exe_path_list, self.output_list = self.run_algorithm_on_f_list(paths)
        self.comb_model_list = [self.fit_with_old_model(torch.as_tensor(np.asarray(ep.x)),
                                                        torch.as_tensor(np.asarray(ep.y)).reshape(-1))
                                for ep in self.exe_path_list]

    # -- plain torch part (not accelerated) -----------------------------------------------------
    def fit_with_old_model(self, data_x, data_y):
        """The GP conditioned on an execution path (``condition_on_observations``, :177-203):
        same hyper-parameters and the OLD standardization, data = old data + path."""
        from ..models import SingleTaskGP

        m = self.model
        y_mean, y_std = m.y_mean_std()
        X = torch.cat([m.train_inputs[0], data_x.to(torch.float64).reshape(-1, m.train_inputs[0].shape[1])])
        Y = torch.cat([m.train_targets * y_std + y_mean, data_y.to(torch.float64).reshape(-1)])
        new = SingleTaskGP(X, Y, kernel=m.kernel, lengthscale=m.covar_module.base_kernel.lengthscale,
                           outputscale=float(m.covar_module.outputscale), noise=float(m.likelihood.noise.mean()),
                           mean_const=float(m.mean_module.constant), standardize=False)
        new.train_targets = (Y - y_mean) / y_std  # keep the old outcome transform
        new.outcome_transform = m.outcome_transform
        return new

    @staticmethod
    def entropy(cov):
        return 0.5 * math.log(2 * math.pi) + 0.5 * torch.logdet(cov)

    def forward(self, X):
        """X [b, q, d] -> EIG estimate [b] (``acq_exe_normal``, :156-174)."""
        if X.dim() == 2:
            X = X.unsqueeze(0)
        if self.X_pending is not None:
            X = torch.cat([X, self.X_pending.to(X).expand(X.shape[0], -1, -1)], dim=-2)
        h_post = self.entropy(self.model.posterior(X).covariance_matrix)
        h_cond = torch.stack([self.entropy(cm.posterior(X).covariance_matrix) for cm in self.comb_model_list])
        return h_post - h_cond.mean(0)

    __call__ = forward

    MAX_FUSED_ROWS = 2048  # training + pending + execution-path rows one fused launch takes as kernel features

    def q1_values(self, choices: torch.Tensor) -> torch.Tensor:
        """``forward(choices.unsqueeze(-2))`` for every row of ``choices`` [m, d] in ONE fused launch for the
        current model and all S fantasy models (``psbax_eval_sumsq_groups`` over the pseudo-paths of
        ``fantasy_variance_params``: the fantasy models share the Cholesky rows of the current one).

        EIG(x) = 0.5 log var_0(x | P) - mean_s 0.5 log var_s(x | P) (+ the x-independent pending terms),
        variances floored at the float32 resolution of the difference alpha - sumsq (1e-6 alpha): rows
        whose variance the data already explain cannot win the arg max by rounding noise."""
        from ..sample_paths import SamplePathBatch, fantasy_variance_params

        choices = torch.as_tensor(choices)
        d = choices.shape[-1]
        path_xs = [torch.as_tensor(np.asarray(ep.x)).reshape(-1, d) for ep in self.exe_path_list]
        pend = self.X_pending.reshape(-1, d).to(torch.float64) if self.X_pending is not None else None
        n_pend = 0 if pend is None else pend.shape[0]
        rows = self.model.train_inputs[0].shape[0] + n_pend + sum(p.shape[0] for p in path_xs)
        if rows > self.MAX_FUSED_ROWS:
            return self._q1_values_per_model(choices)
        group = 16
        params, g0, gs = fantasy_variance_params(self.model, path_xs, pending=pend, group=group)
        try:
            batch = SamplePathBatch(**params, engine="tensor")  # 3xTF32: alpha - sumsq cancels
        except Exception as ex:
            if "engine needs" not in str(ex):
                raise
            batch = SamplePathBatch(**params, engine="simt")
        sq = batch.sumsq_groups(choices, group).double()  # [g0 + sum(gs), m]
        alpha = float(self.model.covar_module.outputscale)
        floor = 1e-6 * alpha
        var0 = alpha - sq[:g0].sum(0)
        val = 0.5 * torch.log(var0.clamp_min(floor))
        cond = torch.zeros_like(val)
        g = g0
        for gk in gs:
            cond += 0.5 * torch.log((var0 - sq[g:g + gk].sum(0)).clamp_min(floor))
            g += gk
        val = val - cond / len(gs)
        if n_pend:  # the terms that do not depend on x
            P = pend.reshape(1, -1, d)
            const = self.entropy(self.model.posterior(P).covariance_matrix[0]) - torch.stack(
                [self.entropy(cm.posterior(P).covariance_matrix[0]) for cm in self.comb_model_list]).mean()
            val = val + const.to(val.device)
        return val.to(choices.device)

    def _q1_values_per_model(self, choices: torch.Tensor) -> torch.Tensor:
        """The same numbers through S + 1 fused launches, one (n + k) x (n + k) factorisation each (kept as
        the cross-check of ``q1_values`` and for execution paths too long for one launch).

        logdet Cov_M([x, P]) = logdet Cov_M(P) + log var_M(x | P) for the current model M_0 and for
        every fantasy model M_s (data + execution path s, observed with the likelihood noise), P the
        pending rows (noise-free).  Every conditional variance is one ``psbax_eval_sumsq`` over the
        pseudo-paths alpha L^-1 of the corresponding augmented training set."""
        from ..sample_paths import posterior_variance

        choices = torch.as_tensor(choices)
        d = choices.shape[-1]
        noise = float(self.model.likelihood.noise.mean())
        pend = self.X_pending.reshape(-1, d).to(torch.float64) if self.X_pending is not None else None
        n_pend = 0 if pend is None else pend.shape[0]

        def log_var(extra_x):
            rows, rnoise = [], []
            if extra_x is not None and extra_x.numel() > 0:
                rows.append(extra_x.reshape(-1, d).to(torch.float64))
                rnoise.append(torch.full((rows[-1].shape[0],), noise, dtype=torch.float64))
            if n_pend:
                rows.append(pend)
                rnoise.append(torch.zeros(n_pend, dtype=torch.float64))
            if not rows:
                var = posterior_variance(self.model, choices)
            else:
                var = posterior_variance(self.model, choices, pending=torch.cat(rows),
                                         pending_noise=torch.cat(rnoise))
            _, y_std = self.model.y_mean_std()
            return torch.log(var.double().clamp_min(1e-6 * float(self.model.covar_module.outputscale) * float(y_std) ** 2))

        val = 0.5 * log_var(None)
        cond = torch.zeros_like(val)
        for ep in self.exe_path_list:
            cond += 0.5 * log_var(torch.as_tensor(np.asarray(ep.x)))
        val = val - cond / len(self.exe_path_list)
        if n_pend:  # the terms that do not depend on x
            P = pend.reshape(1, -1, d)
            const = self.entropy(self.model.posterior(P).covariance_matrix[0]) - torch.stack(
                [self.entropy(cm.posterior(P).covariance_matrix[0]) for cm in self.comb_model_list]).mean()
            val = val + const.to(val.device)
        return val.to(choices.device)
