"""Local BO algorithm: maximise a sample path by multi-start box-constrained quasi-Newton ascent
(interface of ``/root/reference/src/algorithms/lbfgsb.py:9-50``; the reference delegates to botorch
``optimize_acqf`` with raw_samples = 100 d scoring points and num_restarts = 5 d L-BFGS-B ascents,
maxiter 100, ``src/utils.py:120-150``).

Here the raw-sample scoring is one fused top-k launch for all S paths and ALL restarts x samples
ascents run inside one ``psbax_ascend`` launch (one CTA per ascent, projected L-BFGS on the device);
``optimizer="scipy"`` keeps the serial scipy L-BFGS-B loop over the CUDA gradient kernel as a
cross-check.
"""
from argparse import Namespace

import numpy as np
import torch

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm


class LBFGSB(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "LBFGSB")
        self.params.opt_mode = params.get("opt_mode", "max")
        self.params.n_dim = params.get("n_dim", None)
        self.params.bounds = params.get("bounds", None)
        self.params.raw_samples = params.get("raw_samples", None)
        self.params.num_restarts = params.get("num_restarts", None)
        self.params.maxiter = params.get("maxiter", 100)
        self.params.optimizer = params.get("optimizer", "device")  # "device" | "scipy"

    def initialize(self):
        self.exe_path = Namespace(x=[], y=[])

    def _starts(self, paths: SamplePathBatch):
        """raw_samples uniform points of [0, 1]^d scored on every path in one fused launch; the
        num_restarts best per path are the starting points: ([S, R, d] float32, R)."""
        d = self.params.n_dim or paths.d
        raw = self.params.raw_samples or 100 * d
        restarts = min(self.params.num_restarts or 5 * d, raw)
        X0 = torch.rand(raw, d, dtype=torch.float64).to(paths.device, torch.float32)
        _, idx = paths.topk(X0, restarts)
        return X0[idx], restarts

    def run_algorithm_on_batch(self, paths: SamplePathBatch):
        """All S paths at once: S x num_restarts ascents in one launch -> list of (exe_path, output)."""
        if self.params.optimizer == "scipy":
            return [self.get_copy().run_algorithm_on_f(p) for p in paths]
        starts, R = self._starts(paths)
        S, d = paths.S, paths.d
        sidx = torch.arange(S, device=paths.device, dtype=torch.int32).repeat_interleave(R)
        x, f, _ = paths.ascend(starts.reshape(S * R, d), sidx, maxiter=self.params.maxiter)
        f, x = f.reshape(S, R), x.reshape(S, R, d)
        best = f.argmax(1)
        bx = x[torch.arange(S, device=x.device), best].double().cpu().numpy()
        by = f.max(1).values.double().cpu().numpy()
        out = []
        for s in range(S):
            algo = self.get_copy()
            algo.exe_path = Namespace(x=bx[s].reshape(1, d), y=float(by[s]))
            out.append((algo.exe_path, algo.exe_path.x))
        return out

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePathBatch):
            f = f.path(0)
        if not isinstance(f, SamplePath):
            raise TypeError("LBFGSB needs a SamplePath (value + gradient on the GPU)")
        one = f._one()
        if self.params.optimizer != "scipy":
            self.exe_path, out = self.run_algorithm_on_batch(one)[0]
            return self.exe_path, out
        from scipy.optimize import minimize

        d = one.d
        starts, _ = self._starts(one)
        best_x, best_y = None, -np.inf

        def neg(xv):
            val, g = one.grad(torch.from_numpy(xv.reshape(1, d)))
            return -float(val[0]), -g[0].double().cpu().numpy()

        for x0 in starts[0].double().cpu().numpy():
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
