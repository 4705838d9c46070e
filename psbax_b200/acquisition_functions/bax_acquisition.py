"""INFO-BAX acquisition: the second caller of the sample-path hot path
(``/root/reference/src/acquisition_functions/bax_acquisition.py:45-85, 221-277``).

``initialize`` / ``run_algorithm_on_f_list`` are on the accelerated path: the reference draws
``exe_paths`` (default 30) sample paths one by one and deep-copies the algorithm (including X*) per
path; here the paths are drawn as one batch and the algorithm runs on all of them in one fused
GPU launch.  ``forward`` (the expected-information-gain estimate, ``bax_acquisition.py:87-174``) is
outside the accelerated scope; a plain torch implementation for q = 1
is provided so that the "bax" policy is usable: H[y(x)] - mean_s H[y(x) | execution path s] with
exact GP conditioning under fixed hyper-parameters and fixed standardization.
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
