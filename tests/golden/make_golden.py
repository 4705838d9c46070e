#!/usr/bin/env python3
"""Generate ``tests/golden/algorithms_golden.npz`` by running the REFERENCE's own
algorithm classes (``/root/reference/src/algorithms/{topk,levelset,discobax}.py``)
in the build container.  Run once, here; the .npz is committed, the reference is
never read at test time (it does not exist on the GPU box).

    python tests/golden/make_golden.py

The reference modules import ``botorch.acquisition.analytic.PosteriorMean`` for
an ``isinstance`` check only (``src/algorithms/topk.py:4,25``,
``src/algorithms/levelset.py:4,26``).  botorch is not installed in this image,
so a stub module providing a ``PosteriorMean`` class with botorch's calling
convention (X [b, 1, d] -> [b]) is injected before the import.  Nothing else of
the reference is stubbed: the reductions that produce the golden outputs are the
reference's code, unmodified.

``SubsetSelect`` needs an objective exposing ``get_x()``, ``df.index`` and
``get_noisy_f_lst(fx)``; ``src/problems.py`` imports gpytorch at module level, so
a stand-in with the additive-noise arithmetic of ``src/problems.py:146-163``
(``values = fx + etas``; optional ``max(0, .)``) is used, plus one with the multiplicative arithmetic
of ``:150-160`` (``values = fx * mask``) whose Bernoulli mask is drawn once and stored.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "algorithms_golden.npz")


def _install_botorch_stub():
    class PosteriorMean:  # botorch.acquisition.analytic.PosteriorMean calling convention
        def __init__(self, fn):
            self.fn = fn

        def __call__(self, X):
            # @t_batch_mode_transform(expected_q=1): [b, 1, d] -> [b]; a bare [1, d]
            # (q=1, no t-batch; discobax.py:34) gets a t-batch added and removed -> []
            if X.dim() == 2:
                assert X.shape[0] == 1, "PosteriorMean expects q == 1"
                return self.fn(X)[0]
            assert X.dim() == 3 and X.shape[1] == 1, "PosteriorMean expects [b, 1, d]"
            return self.fn(X.squeeze(1))

    botorch = types.ModuleType("botorch")
    acq = types.ModuleType("botorch.acquisition")
    analytic = types.ModuleType("botorch.acquisition.analytic")
    analytic.PosteriorMean = PosteriorMean
    botorch.acquisition = acq
    acq.analytic = analytic
    sys.modules["botorch"] = botorch
    sys.modules["botorch.acquisition"] = acq
    sys.modules["botorch.acquisition.analytic"] = analytic
    return PosteriorMean


def smooth_fn(A, c, w):
    """A deterministic float64 test function f(X) = cos(X A + c) w, X [b, d] -> [b]."""
    A, c, w = (torch.as_tensor(t, dtype=torch.float64) for t in (A, c, w))

    def f(X):
        X = torch.as_tensor(X, dtype=torch.float64)
        return torch.cos(X @ A + c) @ w

    return f


class _FakeDiscoObjective:
    """Stand-in for ``DiscoBAXObjective`` (``src/problems.py:55-173``)."""

    def __init__(self, x, etas, nonneg):
        self._x = torch.as_tensor(x, dtype=torch.float64)
        self.etas = np.asarray(etas)
        self.nonneg = nonneg
        self.df = types.SimpleNamespace(index=[f"g{i}" for i in range(self._x.shape[0])])

    def get_x(self, idx=None):
        return self._x

    def get_noisy_f_lst(self, fx):  # src/problems.py:146-163, additive branch
        values = fx.squeeze() + self.etas
        if self.nonneg:
            values = np.maximum(0, values)
        return values


class _FakeDiscoObjectiveMul(_FakeDiscoObjective):
    """Multiplicative branch of ``get_noisy_f_lst`` (``src/problems.py:150-160``): values = fx * mask with
    mask ~ Bernoulli(sigmoid(etas)).  The reference draws the mask from the global NumPy RNG in every
    call; here it is drawn ONCE (seeded) and stored, so that the golden case is reproducible — the
    greedy that consumes ``values`` is still the reference's."""

    def __init__(self, x, etas, nonneg, seed):
        super().__init__(x, etas, nonneg)
        p = 1.0 / (1.0 + np.exp(-self.etas))
        self.mask = np.random.default_rng(seed).binomial(1, p).astype(np.float64)

    def get_noisy_f_lst(self, fx):
        values = np.multiply(fx.squeeze(), self.mask)
        if self.nonneg:
            values = np.maximum(0, values)
        return values


def main():
    PosteriorMean = _install_botorch_stub()
    # every reference entry point runs in float64 (demos/topk.py:11, src/problems.py:12)
    torch.set_default_dtype(torch.float64)
    sys.path.insert(0, REF)
    from src.algorithms.topk import TopK
    from src.algorithms.levelset import LevelSetEstimator
    from src.algorithms.discobax import SubsetSelect

    rng = np.random.default_rng(20241020)
    out = {}
    case = 0
    # ---- TopK: plain callable and PosteriorMean-style callable --------------------
    for (N, d, k, as_pm) in [(100, 3, 4, False), (100, 3, 4, True), (1000, 5, 10, True),
                             (257, 2, 1, False), (3375, 3, 4, True)]:
        A = rng.normal(size=(d, 16)) * 3.0
        c = rng.uniform(0, 2 * np.pi, size=16)
        w = rng.normal(size=16)
        x = rng.uniform(size=(N, d))
        f = smooth_fn(A, c, w)
        algo = TopK({"x_path": x.copy(), "k": k})
        exe_path, output = algo.run_algorithm_on_f(PosteriorMean(f) if as_pm else f)
        pre = f"topk{case}_"
        out.update({pre + "A": A, pre + "c": c, pre + "w": w, pre + "x": x,
                    pre + "k": np.int64(k), pre + "as_pm": np.bool_(as_pm),
                    pre + "idx": np.asarray(exe_path.idx, dtype=np.int64),
                    pre + "out_x": np.asarray(output), pre + "out_y": np.asarray(exe_path.y)})
        case += 1
    out["n_topk"] = np.int64(case)

    # ---- LevelSetEstimator, including the empty-set fallback -----------------------
    case = 0
    for (N, d, thr, as_pm) in [(2500, 2, 0.3, True), (2500, 2, 0.3, False), (500, 3, 1e9, True),
                               (128, 2, -1e9, True), (1000, 4, 1.1, True)]:
        A = rng.normal(size=(d, 16)) * 2.0
        c = rng.uniform(0, 2 * np.pi, size=16)
        w = rng.normal(size=16) * 0.5
        x = rng.uniform(size=(N, d))
        f = smooth_fn(A, c, w)
        algo = LevelSetEstimator({"name": "SimpleLevelSet", "threshold": thr, "x_set": x.copy()})
        exe_path, output = algo.run_algorithm_on_f(PosteriorMean(f) if as_pm else f)
        pre = f"ls{case}_"
        out.update({pre + "A": A, pre + "c": c, pre + "w": w, pre + "x": x,
                    pre + "thr": np.float64(thr), pre + "as_pm": np.bool_(as_pm),
                    pre + "idx": np.asarray(exe_path.idx, dtype=np.int64),
                    pre + "out_x": np.asarray(output), pre + "out_y": np.asarray(exe_path.y)})
        case += 1
    out["n_ls"] = np.int64(case)

    # ---- SubsetSelect (DiscoBAX greedy), fx given as ndarray and as callable --------
    case = 0
    for (N, d, B, k, nonneg, as_array) in [(60, 3, 7, 4, False, True), (60, 3, 7, 4, True, True),
                                           (40, 5, 5, 3, False, False), (200, 4, 20, 10, True, True)]:
        A = rng.normal(size=(d, 16)) * 2.0
        c = rng.uniform(0, 2 * np.pi, size=16)
        w = rng.normal(size=16) * 0.5
        x = rng.uniform(size=(N, d))
        etas = rng.normal(size=(B, N)) * 0.7
        f = smooth_fn(A, c, w)
        obj = _FakeDiscoObjective(x, etas, nonneg)
        algo = SubsetSelect({"k": k})
        algo.set_obj_func(obj)
        if as_array:
            arg = f(x).numpy()
        else:
            # exercises the per-point loop of discobax.py:32-34
            arg = PosteriorMean(f)
        exe_path, indices = algo.run_algorithm_on_f(arg)
        pre = f"ss{case}_"
        out.update({pre + "A": A, pre + "c": c, pre + "w": w, pre + "x": x, pre + "etas": etas,
                    pre + "k": np.int64(k), pre + "nonneg": np.bool_(nonneg),
                    pre + "as_array": np.bool_(as_array),
                    pre + "idxes": np.asarray(exe_path.idxes, dtype=np.int64),
                    pre + "labels": np.asarray(indices), pre + "out_y": np.asarray(exe_path.y),
                    pre + "values": np.asarray(exe_path.values)})
        case += 1
    out["n_ss"] = np.int64(case)

    # ---- SubsetSelect with the multiplicative noise model (src/problems.py:150-160) --------
    case = 0
    for (N, d, B, k, nonneg) in [(80, 3, 9, 4, False), (150, 4, 16, 4, True)]:  # k small enough that no greedy step is an exact tie
        A = rng.normal(size=(d, 16)) * 2.0
        c = rng.uniform(0, 2 * np.pi, size=16)
        w = rng.normal(size=16) * 0.5
        x = rng.uniform(size=(N, d))
        etas = rng.normal(size=(B, N))
        f = smooth_fn(A, c, w)
        obj = _FakeDiscoObjectiveMul(x, etas, nonneg, seed=100 + case)
        algo = SubsetSelect({"k": k})
        algo.set_obj_func(obj)
        exe_path, indices = algo.run_algorithm_on_f(f(x).numpy())
        pre = f"sm{case}_"
        out.update({pre + "A": A, pre + "c": c, pre + "w": w, pre + "x": x, pre + "mask": obj.mask,
                    pre + "k": np.int64(k), pre + "nonneg": np.bool_(nonneg),
                    pre + "idxes": np.asarray(exe_path.idxes, dtype=np.int64),
                    pre + "labels": np.asarray(indices), pre + "values": np.asarray(exe_path.values)})
        case += 1
    out["n_sm"] = np.int64(case)

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
