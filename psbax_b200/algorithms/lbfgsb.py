"""Local BO algorithm: maximise one sample path by multi-start L-BFGS-B
(interface of ``/root/reference/src/algorithms/lbfgsb.py:9-50``; the reference
delegates to botorch ``optimize_acqf`` with raw_samples = 100 d scoring points and
num_restarts = 5 d ascents, ``src/utils.py:120-150``).

Here the raw-sample scoring is one fused arg-max/top-k launch and every L-BFGS-B
iterate gets f and df/dx from the CUDA gradient kernel.
"""
from argparse import Namespace

import numpy as np
import torch
from scipy.optimize import minimize

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm


class LBFGSB(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "LBFGSB")
        self.params.opt_mode = params.get("opt_mode", "max")
        self.params.n_dim = params.get("n_dim", None)
        self.params.raw_samples = params.get("raw_samples", None)
        self.params.num_restarts = params.get("num_restarts", None)
        self.params.maxiter = params.get("maxiter", 100)

    def initialize(self):
        self.exe_path = Namespace(x=[], y=[])

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePathBatch):
            f = f.path(0)
        if not isinstance(f, SamplePath):
            raise TypeError("LBFGSB needs a SamplePath (value + gradient on the GPU)")
        d = self.params.n_dim or f.batch.d
        raw = self.params.raw_samples or 100 * d
        restarts = self.params.num_restarts or 5 * d
        one = f._one()
        X0 = torch.rand(raw, d, dtype=torch.float64)
        _, idx = one.topk(X0, min(restarts, raw))
        starts = X0[idx[0].cpu()]
        best_x, best_y = None, -np.inf

        def neg(xv):
            val, g = one.grad(torch.from_numpy(xv.reshape(1, d)))
            return -float(val[0]), -g[0].double().cpu().numpy()

        for x0 in starts.numpy():
            res = minimize(neg, x0, jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * d,
                           options={"maxiter": self.params.maxiter})
            if -res.fun > best_y:
                best_x, best_y = res.x.copy(), -res.fun
        self.exe_path.x = best_x.reshape(1, d)
        self.exe_path.y = best_y
        return self.exe_path, self.exe_path.x

    def execute(self, f):
        self.run_algorithm_on_f(f)
        return self.exe_path.x

    def get_output(self):
        return self.exe_path.x
