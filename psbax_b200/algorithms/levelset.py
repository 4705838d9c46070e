"""Level-set BAX algorithm (interface and results of
``/root/reference/src/algorithms/levelset.py:10-56``)."""
from argparse import Namespace

import numpy as np
import torch

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm, eval_in_batch


class LevelSetEstimator(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "SimpleLevelSet")
        self.params.threshold = params.get("threshold", 10)
        self.params.x_set = params.get("x_set", None)  # numpy [N, d], as in the reference
        self.params.no_copy = params.get("no_copy", None)
        self._dev_x = None

    def initialize(self):
        self.exe_path = Namespace()

    def device_candidates(self, device) -> torch.Tensor:
        if self._dev_x is None or self._dev_x.device != torch.device(device):
            self._dev_x = torch.as_tensor(self.params.x_set).to(device, torch.float32).contiguous()
        return self._dev_x

    def _finish(self, idx, fx_at_idx):
        x_set = self.params.x_set
        self.exe_path.idx = list(idx)
        self.exe_path.x = np.atleast_2d(np.asarray(x_set)[idx])
        self.exe_path.y = np.atleast_1d(fx_at_idx)
        return self.exe_path, self.exe_path.x

    def _from_values(self, fx: np.ndarray):
        idx = np.where(fx > self.params.threshold)[0]
        if len(idx) == 0:  # empty super-level set -> the arg max (levelset.py:35-37)
            idx = np.array([int(np.argmax(fx))])
        return self._finish(idx, fx[idx])

    def _from_result(self, res, s, y_of):
        idx = res.indices(s).cpu().numpy()
        return self._finish(idx, y_of(idx))

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePathBatch):
            if f.S != 1:
                raise ValueError("run_algorithm_on_f takes one path; use run_algorithm_on_batch")
            f = f.path(0)
        if isinstance(f, SamplePath):
            batch = f._one()
            X = self.device_candidates(batch.device)
            res = batch.levelset(X, self.params.threshold)
            idx = res.indices(0)
            y = batch.values(X[idx])[:, 0].double().cpu().numpy()
            return self._finish(idx.cpu().numpy(), y)
        if getattr(f, "expects_q_dim", False):
            x_set = torch.from_numpy(np.asarray(self.params.x_set)).unsqueeze(1)
            fx = eval_in_batch(f, x_set).detach().numpy().flatten()
        else:
            fx = np.asarray(f(self.params.x_set)).flatten()
        return self._from_values(fx)

    def run_algorithm_on_batch(self, paths: SamplePathBatch, with_values: bool = True, chunk_rows: int = 1 << 18):
        """One fused launch for all S paths -> [(exe_path, output)] in sample order.

        ``exe_path.y`` comes from the SAME batch and engine that decided membership: the values of the
        union of the S super-level sets are evaluated once (row chunks of the union, all S paths per
        launch) and every path gathers its own rows — no per-path repack / launch, and y > threshold
        holds for every reported row up to the engine's rounding of one and the same number."""
        X = self.device_candidates(paths.device)
        N = X.shape[0]
        res = paths.levelset(X, self.params.threshold, want_mask=True, want_pooled=True)
        idxs = [res.indices(s) for s in range(paths.S)]  # ascending row indices per path (or [arg max])
        ys = [torch.full((i.numel(),), float("nan"), dtype=torch.float64, device=paths.device) for i in idxs]
        if with_values:
            shifts = torch.arange(32, device=paths.device, dtype=torch.int32)
            pool = ((res.pooled.unsqueeze(1) >> shifts) & 1).bool().reshape(-1)[:N].clone()
            empty = res.count == 0
            if bool(empty.any()):
                pool[res.argmax[empty]] = True
            union = pool.nonzero().reshape(-1)
            for rows in union.split(chunk_rows):
                V = paths.values(X[rows]).double()  # [c, S]
                lo, hi = rows[0], rows[-1]
                for s, idx in enumerate(idxs):
                    a, b = int(torch.searchsorted(idx, lo)), int(torch.searchsorted(idx, hi, right=True))
                    if b > a:
                        ys[s][a:b] = V[torch.searchsorted(rows, idx[a:b]), s]
        sizes = [i.numel() for i in idxs]
        idx_h = torch.cat(idxs).cpu().numpy()
        y_h = torch.cat(ys).cpu().numpy()
        results, o = [], 0
        for s in range(paths.S):
            self.initialize()
            results.append(self._finish(idx_h[o:o + sizes[s]], y_h[o:o + sizes[s]]))
            o += sizes[s]
        return results

    def pooled_mask(self, paths: SamplePathBatch) -> torch.Tensor:
        """bool [N] on the GPU: candidate i belongs to the output of at least one path — the pool
        ``gen_posterior_sampling_batch`` builds (``posterior_sampling.py:71-78``) without its
        duplicates and without ever materialising the S x |super-level set| coordinate rows: one
        fused launch, an OR over the S bit masks, plus the arg max row of every path whose set is
        empty (``levelset.py:35-37``)."""
        X = self.device_candidates(paths.device)
        N = X.shape[0]
        # the OR over the S masks is formed inside the kernel; the S x N / 8 byte masks are never written
        res = paths.levelset(X, self.params.threshold, want_mask=False, want_pooled=True)
        shifts = torch.arange(32, device=res.pooled.device, dtype=torch.int32)
        bits = ((res.pooled.unsqueeze(1) >> shifts) & 1).bool().reshape(-1)[:N].clone()
        empty = res.count == 0
        if bool(empty.any()):
            bits[res.argmax[empty]] = True
        return bits

    def pooled_union(self, paths: SamplePathBatch) -> torch.Tensor:
        """Row indices (ascending, int64, CPU) of ``pooled_mask``."""
        return self.pooled_mask(paths).nonzero().reshape(-1).cpu()

    def execute(self, f):
        self.run_algorithm_on_f(f)
        return self.exe_path.x

    def get_output(self):
        return self.exe_path.x
