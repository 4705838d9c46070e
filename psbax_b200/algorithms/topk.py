"""Top-k BAX algorithm (interface and results of ``/root/reference/src/algorithms/topk.py:9-49``)."""
from argparse import Namespace

import numpy as np
import torch

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm, eval_in_batch


def _stable_topk(fx: torch.Tensor, k: int) -> torch.Tensor:
    """k largest, largest first, exact ties towards the smaller index (``torch.topk``
    leaves tie order unspecified; the GPU path uses the same rule)."""
    order = torch.sort(fx, descending=True, stable=True).indices
    return order[:k]


class TopK(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "TopK")
        self.params.k = params.get("k", 10)
        x_path = params.get("x_path", None)
        if isinstance(x_path, np.ndarray):
            x_path = torch.from_numpy(x_path)
        self.params.x_path = x_path
        self.params.no_copy = params.get("no_copy", None)
        self._dev_x = None

    def initialize(self):
        self.exe_path = Namespace()

    def device_candidates(self, device) -> torch.Tensor:
        """X* as float32 on ``device``, uploaded once and kept resident."""
        if self._dev_x is None or self._dev_x.device != torch.device(device):
            self._dev_x = self.params.x_path.to(device, torch.float32).contiguous()
        return self._dev_x

    def _finish(self, idx: torch.Tensor, y: torch.Tensor):
        x_path = self.params.x_path
        self.exe_path.idx = idx.tolist()
        self.exe_path.x = x_path[idx]
        self.exe_path.y = y.to(x_path.dtype).reshape(-1, 1)
        return self.exe_path, self.exe_path.x

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePathBatch):
            if f.S != 1:
                raise ValueError("run_algorithm_on_f takes one path; use run_algorithm_on_batch")
            f = f.path(0)
        if isinstance(f, SamplePath):
            batch = f._one()
            val, idx = batch.topk(self.device_candidates(batch.device), self.params.k)
            return self._finish(idx[0].cpu(), val[0].cpu())
        # foreign callable: the reference protocol (topk.py:25-29)
        x_path = self.params.x_path
        xin = x_path.unsqueeze(1) if getattr(f, "expects_q_dim", False) else x_path
        fx = eval_in_batch(f, xin).detach().reshape(-1)
        idx = _stable_topk(fx, self.params.k)
        return self._finish(idx, fx[idx])

    def run_algorithm_on_batch(self, paths: SamplePathBatch):
        """One fused launch for all S paths -> [(exe_path, output)] in sample order."""
        val, idx = paths.topk(self.device_candidates(paths.device), self.params.k)
        val, idx = val.cpu(), idx.cpu()
        results = []
        for s in range(paths.S):
            self.initialize()
            results.append(self._finish(idx[s], val[s]))
        return results

    def execute(self, f):
        self.run_algorithm_on_f(f)
        return self.exe_path.x

    def get_output(self):
        return self.exe_path.x
