"""DiscoBAX subset selection (interface and results of
``/root/reference/src/algorithms/discobax.py:9-83``)."""
from argparse import Namespace

import numpy as np
import torch

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm


class SubsetSelect(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "SubsetSelect")
        self.params.k = params.get("k", 3)
        self._dev_x = None
        self._dev_etas = None

    def initialize(self):
        self.exe_path = Namespace(x=[], y=[], idxes=[], values=[])
        self.x_torch = self.obj_func.get_x()

    def index_to_x(self, idx):
        return self.obj_func.get_x(idx)

    def set_obj_func(self, obj_func):
        self.obj_func = obj_func
        self._dev_x = self._dev_etas = None

    def _device_inputs(self, device):
        if self._dev_x is None or self._dev_x.device != torch.device(device):
            self._dev_x = torch.as_tensor(self.x_torch).to(device, torch.float32).contiguous()
            self._dev_etas = torch.as_tensor(np.asarray(self.obj_func.etas)).to(device, torch.float32).contiguous()
        return self._dev_x, self._dev_etas

    def _finish(self, idxes, fx, values):
        idxes = [int(i) for i in idxes]
        labels = [self.obj_func.df.index[i] for i in idxes]
        self.exe_path.x = torch.as_tensor(self.x_torch)[idxes].numpy()
        self.exe_path.values = values
        self.exe_path.idxes = idxes
        self.exe_path.y = fx[idxes]
        return self.exe_path, labels

    def _greedy_host(self, fx: np.ndarray):
        """Greedy of discobax.py:37-56 for values that are already on the host
        (true objective / foreign callable); ties to the smaller index."""
        values = self.obj_func.get_noisy_f_lst(fx)
        idxes = [int(np.argmax(values.mean(axis=0)))]
        cur = values[:, idxes[0]].copy()
        for _ in range(self.params.k - 1):
            e_vals = np.maximum(cur[:, None], values).mean(axis=0)
            e_vals[idxes] = 0.0
            j = int(np.argmax(e_vals))
            idxes.append(j)
            cur = np.maximum(cur, values[:, j])
        return self._finish(idxes, fx, values)

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePathBatch):
            if f.S != 1:
                raise ValueError("run_algorithm_on_f takes one path; use run_algorithm_on_batch")
            f = f.path(0)
        if isinstance(f, SamplePath):
            return self.run_algorithm_on_batch(f._one())[0]
        if isinstance(f, np.ndarray):
            return self._greedy_host(f)
        x = torch.as_tensor(self.x_torch)
        fx = f(x.unsqueeze(1) if getattr(f, "expects_q_dim", False) else x)
        return self._greedy_host(torch.as_tensor(fx).detach().reshape(-1).numpy())

    def run_algorithm_on_batch(self, paths: SamplePathBatch):
        """Values for all S paths in one launch, then the greedy for all S on the GPU.

        Additive noise: the objective's ``etas`` [B, N] serve every path.  Multiplicative noise
        (``src/problems.py:150-160``): the objective draws a fresh Bernoulli mask in every
        ``get_noisy_f_lst`` call, i.e. per path; the S masks are drawn here in path order through the
        objective itself (``get_noisy_f_lst(ones)``: same RNG consumption as S reference calls) and the
        greedy runs on ``fx * mask`` with one mask per path."""
        self.initialize()
        noise_type = getattr(self.obj_func, "noise_type", "additive")
        X, etas = self._device_inputs(paths.device)
        nonneg = bool(getattr(self.obj_func, "nonneg", False))
        fx_dev = paths.values(X)
        fx_all = fx_dev.double().cpu().numpy()  # exe_path.y / .values bookkeeping
        N = fx_all.shape[0]
        if noise_type == "additive":
            idx = paths.subset_select(X, etas, self.params.k, nonneg).cpu().numpy()
            values = [None] * paths.S
        elif noise_type == "multiplicative":
            masks = np.stack([np.asarray(self.obj_func.get_noisy_f_lst(np.ones(N)), dtype=np.float64)
                              for _ in range(paths.S)])  # [S, B, N] of 0 / 1
            idx = paths.subset_select(X, torch.from_numpy(masks), self.params.k, nonneg, noise="multiplicative").cpu().numpy()
            values = [np.maximum(0, fx_all[:, s] * masks[s]) if nonneg else fx_all[:, s] * masks[s] for s in range(paths.S)]
        else:
            raise NotImplementedError(f"noise_type {noise_type!r}")
        results = []
        for s in range(paths.S):
            self.initialize()
            fx = fx_all[:, s]
            results.append(self._finish(idx[s], fx, values[s] if values[s] is not None else self.obj_func.get_noisy_f_lst(fx)))
        return results

    def get_values_and_selected_indices(self):
        return self.exe_path.values, self.exe_path.idxes

    def execute(self, f):
        _, output = self.run_algorithm_on_f(f)
        return output

    @staticmethod
    def random_argmax(vals):
        max_val = np.max(vals)
        idxes = np.where(vals == max_val)[0]
        return np.random.choice(idxes)
