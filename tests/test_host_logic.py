"""CPU tests of the host-side mirror of the reference interface: algorithm classes on foreign
callables (checked against the golden vectors the reference's own classes produced), the GP
stand-in, the entropy pick, parameter draws (run on the CPU in float64 and evaluated by the
oracle), and the missing-library failure mode."""
import math
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, make_gp

from psbax_b200.acquisition_functions.entropy import EntropyAcquisitionFunction, optimize_acqf_discrete
from psbax_b200.algorithms import LevelSetEstimator, SubsetSelect, TopK
from psbax_b200.models import SingleTaskGP
from psbax_b200.sample_paths import matheron_params, posterior_mean_params, reference_params


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN, allow_pickle=False)


class _PM:
    """A foreign callable with botorch's PosteriorMean convention: X [b, 1, d] -> [b]."""
    expects_q_dim = True

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, X):
        if X.dim() == 2:
            return self.fn(X)[0]
        assert X.dim() == 3 and X.shape[1] == 1
        return self.fn(X.squeeze(1))


def _fn(g, pre):
    A, c, w = (torch.from_numpy(g[pre + k]) for k in ("A", "c", "w"))
    return lambda X: torch.cos(torch.as_tensor(X, dtype=torch.float64) @ A + c) @ w


def test_topk_class_matches_reference_golden(golden):
    for i in range(int(golden["n_topk"])):
        pre = f"topk{i}_"
        f = _fn(golden, pre)
        algo = TopK({"x_path": golden[pre + "x"].copy(), "k": int(golden[pre + "k"])})
        exe_path, out = algo.run_algorithm_on_f(_PM(f) if bool(golden[pre + "as_pm"]) else f)
        assert exe_path.idx == golden[pre + "idx"].tolist()
        np.testing.assert_array_equal(out.numpy(), golden[pre + "out_x"])
        np.testing.assert_allclose(exe_path.y.numpy(), golden[pre + "out_y"], atol=1e-12)
        assert algo.execute(f).shape == out.shape


def test_levelset_class_matches_reference_golden(golden):
    for i in range(int(golden["n_ls"])):
        pre = f"ls{i}_"
        f = _fn(golden, pre)
        algo = LevelSetEstimator({"name": "SimpleLevelSet", "threshold": float(golden[pre + "thr"]),
                                  "x_set": golden[pre + "x"].copy()})
        exe_path, out = algo.run_algorithm_on_f(_PM(f) if bool(golden[pre + "as_pm"]) else f)
        assert [int(j) for j in exe_path.idx] == golden[pre + "idx"].tolist()
        np.testing.assert_array_equal(out, golden[pre + "out_x"])
        np.testing.assert_allclose(exe_path.y, golden[pre + "out_y"], atol=1e-12)


class _Obj:
    def __init__(self, x, etas, nonneg):
        self._x, self.etas, self.nonneg = torch.as_tensor(x), np.asarray(etas), nonneg
        self.noise_type = "additive"
        self.df = types.SimpleNamespace(index=[f"g{i}" for i in range(self._x.shape[0])])

    def get_x(self, idx=None):
        return self._x

    def get_noisy_f_lst(self, fx):
        v = fx.squeeze() + self.etas
        return np.maximum(0, v) if self.nonneg else v


def test_subset_select_class_matches_reference_golden(golden):
    for i in range(int(golden["n_ss"])):
        pre = f"ss{i}_"
        f = _fn(golden, pre)
        x = golden[pre + "x"]
        algo = SubsetSelect({"k": int(golden[pre + "k"])})
        algo.set_obj_func(_Obj(x, golden[pre + "etas"], bool(golden[pre + "nonneg"])))
        arg = f(x).numpy() if bool(golden[pre + "as_array"]) else f
        exe_path, labels = algo.run_algorithm_on_f(arg)
        assert exe_path.idxes == golden[pre + "idxes"].tolist()
        assert list(labels) == golden[pre + "labels"].tolist()
        np.testing.assert_allclose(exe_path.values, golden[pre + "values"], atol=1e-12)
        assert algo.get_values_and_selected_indices()[1] == exe_path.idxes


def test_algorithm_params_and_copy():
    x = np.random.default_rng(0).uniform(size=(10, 2))
    algo = TopK({"x_path": x, "k": 3})
    assert algo.params.name == "TopK" and isinstance(algo.params.x_path, torch.Tensor)
    cp = algo.get_copy()
    assert cp is not algo and torch.equal(cp.params.x_path, algo.params.x_path)
    ls = LevelSetEstimator({"x_set": x, "threshold": 0.5})
    assert ls.params.name == "SimpleLevelSet" and isinstance(ls.params.x_set, np.ndarray)
    with pytest.raises(AttributeError):
        TopK(None)  # params must be a dict, as in the reference


def _model_from(gp):
    return SingleTaskGP(gp.train_X, gp.train_Y, kernel=gp.kernel, lengthscale=gp.lengthscale,
                        outputscale=gp.outputscale, noise=gp.noise, mean_const=gp.mean_const)


def test_gp_standin_posterior_matches_oracle():
    gp = make_gp(14, 3, noise=1e-2)
    m = _model_from(gp)
    np.testing.assert_allclose(m.train_targets.numpy(), gp.train_targets.numpy(), atol=1e-12)
    X = torch.rand(6, 3, dtype=torch.float64)
    post = m.posterior(X)
    mean, cov = O.gp_posterior(gp, X)
    np.testing.assert_allclose(post.mean[:, 0].numpy(), mean.numpy(), atol=1e-9)
    np.testing.assert_allclose(post.covariance_matrix.numpy(), cov.numpy(), atol=1e-9)
    assert post.variance.shape == (6, 1)


def test_entropy_pick_matches_oracle():
    gp = make_gp(10, 2, noise=1e-2)
    m = _model_from(gp)
    choices = torch.rand(23, 2, dtype=torch.float64)
    acq = EntropyAcquisitionFunction(m)
    np.testing.assert_allclose(acq(choices[:5].unsqueeze(1)).numpy(),
                               O.entropy_acqf(gp, choices[:5].unsqueeze(1)).numpy(), atol=1e-9)
    x_next, _ = optimize_acqf_discrete(acq, 3, choices, max_batch_size=7)
    ref, _ = O.optimize_acqf_discrete(gp, choices, 3)
    np.testing.assert_allclose(x_next.numpy(), ref.numpy())
    assert acq.X_pending is None


def _as_oracle(params):
    t = lambda v: None if v is None else v.double().cpu()  # noqa: E731
    return O.PathParams(omega=t(params["omega"]), bias=t(params["bias"]), W=t(params["W"]),
                        kernel=params["kernel"], Xtrain=t(params["Xtrain"]), inv_ls=t(params["inv_ls"]),
                        V=t(params["V"]), mean_const=params["mean_const"], y_mean=params["y_mean"],
                        y_std=params["y_std"])


@pytest.mark.parametrize("draw", [matheron_params, reference_params])
def test_param_draws_interpolate_and_average_to_posterior_mean(draw):
    gp = make_gp(9, 2, noise=1e-6, ls=torch.tensor([0.4, 0.5]))
    m = _model_from(gp)
    gen = torch.Generator().manual_seed(11)
    p = _as_oracle(draw(m, 4, 400, gen, "cpu"))
    f = O.eval_paths(p, gp.train_X)
    assert (f - gp.train_Y[:, None]).abs().max() < 2e-2 * gp.y_std
    assert p.G == (1 if draw is matheron_params else 4)


def test_posterior_mean_params_match_oracle():
    gp = make_gp(12, 3, noise=1e-3)
    m = _model_from(gp)
    X = torch.rand(20, 3, dtype=torch.float64)
    p = _as_oracle(posterior_mean_params(m, "cpu"))
    mean, _ = O.gp_posterior(gp, X)
    np.testing.assert_allclose(O.eval_paths(p, X)[:, 0].numpy(), mean.numpy(), atol=1e-8)


def test_gp_fit_runs_and_improves_mll():
    gp = make_gp(20, 2, noise=1e-2)
    m = SingleTaskGP(gp.train_X, gp.train_Y)
    m.fit(max_iter=30)
    assert torch.isfinite(m.covar_module.base_kernel.lengthscale).all()
    assert float(m.likelihood.noise) >= 1e-4


def test_product_fails_loudly_without_cuda():
    from psbax_b200 import PsbaxError, SamplePathBatch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(PsbaxError):
        SamplePathBatch(torch.zeros(1, 4, 2), torch.zeros(1, 4), torch.zeros(1, 4))


def test_product_does_not_import_oracle():
    import os
    import re

    from helpers import ROOT

    pkg = os.path.join(ROOT, "psbax_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_dkgp_linear_weight_draw_matches_oracle_moments():
    """get_linear_weight_dist (src/utils.py:153-182): the product's draw (run on the CPU) has the
    mean and covariance of the oracle's restatement."""
    from psbax_b200.models import DKGP
    from psbax_b200.sample_paths import dkgp_linear_params

    torch.manual_seed(1)
    g = torch.Generator().manual_seed(0)
    X = torch.rand(30, 6, dtype=torch.float64, generator=g)
    y = torch.cos(X.sum(-1) * 2)
    y = (y - y.mean()) / y.std()
    model = DKGP(X, y, architecture=[6, 10, 4, 1])
    model.likelihood.noise = torch.tensor([0.05], dtype=torch.float64)
    model.mean_module.constant = torch.tensor(0.1, dtype=torch.float64)
    prm = dkgp_linear_params(model, 4000, torch.Generator().manual_seed(2), "cpu")
    with torch.no_grad():
        E = model.embedding(X).double()
    ref = O.draw_dkgp_linear_paths(E, y, 1.0, 0.05, 0.1, 4000, torch.Generator().manual_seed(3))
    Wp, Wo = prm["V"], ref.V
    assert prm["kernel"] == "linear" and Wp.shape == (4000, 4)
    se = Wo.std(0) / math.sqrt(4000)
    assert ((Wp.mean(0) - Wo.mean(0)).abs() < 6 * se + 1e-9).all()
    assert (torch.cov(Wp.T) - torch.cov(Wo.T)).abs().max() < 0.08 * float(torch.cov(Wo.T).abs().max())
    with pytest.raises(AttributeError):
        DKGP(X, y, architecture=[6, 1])  # reference quirk: no hidden layers -> AttributeError


def test_fit_model_boundary():
    from psbax_b200.fit_model import fit_model
    from psbax_b200.models import DKGP, SingleTaskGP

    g = torch.Generator().manual_seed(0)
    X = torch.rand(16, 3, dtype=torch.float64, generator=g)
    y = torch.sin(3 * X.sum(-1))
    m = fit_model(X, y, "gp")
    assert isinstance(m, SingleTaskGP) and m.kernel == "matern52" and m.covar_module.base_kernel.lengthscale.shape == (1, 3)
    m2 = fit_model(X, y, "gp", kernel_type="rbf", state_dict={"lengthscale": 0.4, "outputscale": 1.0, "noise": 1e-3})
    assert m2.kernel == "rbf" and float(m2.likelihood.noise) == 1e-3
    m3 = fit_model(X, y, "dkgp", architecture=[8, 4], epochs=5)
    assert isinstance(m3, DKGP) and m3.architecture == [3, 8, 4, 1]
    assert fit_model(X, y, "nope") is None  # every failure returns None, as in the reference


def test_stream_chunk_bounds():
    """Chunking of the streamed level set: full cover, aligned boundaries, short chunks at both ends."""
    from psbax_b200.sample_paths import _stream_bounds

    for N, c in [(16_777_216, 1 << 21), (50_000, 8192), (4 * 8192, 8192), (4 * 8192 + 1, 8192), (1000, 8192),
                 (40_000, 8192), (123_457, 4096), (16_777_293, 1 << 21)]:
        b = _stream_bounds(N, c)
        sizes = [b[i + 1] - b[i] for i in range(len(b) - 1)]
        assert b[0] == 0 and b[-1] == N and all(s > 0 for s in sizes)
        assert all(x % 256 == 0 for x in b[:-1])  # mask words and row tiles stay aligned
        assert max(sizes) <= c
        if N >= 4 * c:
            assert sizes[0] == c // 8 and sizes[-1] < c // 8 + 256


def test_posterior_variance_pseudo_paths_on_cpu():
    """The alpha L^-1 pseudo-paths of the fused max-entropy pick: their squared norm is the variance
    explained by the data (checked in float64 on the CPU through the oracle's evaluator)."""
    from psbax_b200.sample_paths import posterior_variance_params

    gp = make_gp(20, 3, noise=1e-3, seed=3)
    model = SingleTaskGP(gp.train_X, gp.train_Y, kernel=gp.kernel, lengthscale=gp.lengthscale,
                         outputscale=gp.outputscale, noise=gp.noise, mean_const=gp.mean_const)
    Xq = torch.rand(50, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    pend = Xq[:2]
    for pending in (None, pend):
        kw = posterior_variance_params(model, pending, device="cpu")
        p = O.PathParams(omega=kw["omega"], bias=kw["bias"], W=kw["W"], kernel=kw["kernel"], Xtrain=kw["Xtrain"],
                         inv_ls=kw["inv_ls"], V=kw["V"], mean_const=0.0, y_mean=0.0, y_std=1.0)
        vals = O.eval_paths(p, Xq[2:])  # [N, n (+ m)]
        y_std = float(model.y_mean_std()[1])
        var = (float(model.covar_module.outputscale) - (vals ** 2).sum(1)) * y_std ** 2
        if pending is None:
            ref = torch.stack([O.gp_posterior(gp, x[None])[1].reshape(()) for x in Xq[2:]])
        else:  # var(x | P) = Cov([x, P]) Schur complement
            ref = []
            for x in Xq[2:]:
                cov = O.gp_posterior(gp, torch.cat([x[None], pend]))[1]
                ref.append(cov[0, 0] - cov[0, 1:] @ torch.linalg.solve(cov[1:, 1:], cov[1:, 0]))
            ref = torch.stack(ref)
        assert (var - ref).abs().max() <= 1e-6 * float(ref.max()) + 1e-9


def test_fantasy_pseudo_paths_give_the_oracle_eig_on_cpu():
    """Row f3: the block-Cholesky pseudo-paths of ``fantasy_variance_params`` (current model + one fantasy
    model per execution path, one candidate stream) reproduce the oracle's restatement of
    ``BAXAcquisitionFunction.forward`` (``bax_acquisition.py:87-122, 156-174``) in float64, with and without
    pending rows, and the product's plain-torch ``forward`` equals the oracle too."""
    from psbax_b200.acquisition_functions import BAXAcquisitionFunction
    from psbax_b200.algorithms import TopK
    from psbax_b200.sample_paths import fantasy_variance_params

    gp = make_gp(14, 3, noise=1e-3, seed=6)
    model = SingleTaskGP(gp.train_X, gp.train_Y, kernel=gp.kernel, lengthscale=gp.lengthscale,
                         outputscale=gp.outputscale, noise=gp.noise, mean_const=gp.mean_const)
    g = torch.Generator().manual_seed(2)
    Xq = torch.rand(40, 3, dtype=torch.float64, generator=g)
    paths = [(torch.rand(k, 3, dtype=torch.float64, generator=g), torch.randn(k, dtype=torch.float64, generator=g))
             for k in (4, 4, 7, 1)]
    paths[1] = (torch.cat([paths[1][0][:3], gp.train_X[:1]]), paths[1][1])  # a path row that duplicates a training row
    alpha = float(model.covar_module.outputscale)
    group = 4
    for pending in (None, Xq[:2]):
        ref = O.bax_eig(gp, paths, Xq[2:, None, :], pending)
        kw, g0, gs = fantasy_variance_params(model, [p for p, _ in paths], pending=pending, group=group, device="cpu")
        p = O.PathParams(omega=kw["omega"], bias=kw["bias"], W=kw["W"], kernel=kw["kernel"], Xtrain=kw["Xtrain"],
                         inv_ls=kw["inv_ls"], V=kw["V"], mean_const=0.0, y_mean=0.0, y_std=1.0)
        sq = (O.eval_paths(p, Xq[2:]) ** 2).T.reshape(g0 + sum(gs), group, -1).sum(1)  # [groups, N]
        var0 = alpha - sq[:g0].sum(0)
        eig, a = 0.5 * torch.log(var0), g0
        for gk in gs:
            eig = eig - 0.5 * torch.log(var0 - sq[a:a + gk].sum(0)) / len(gs)
            a += gk
        if pending is not None:  # x-independent terms: H_0(P) - mean_s H_s(P)
            h = lambda m: 0.5 * torch.logdet(O.gp_posterior(m, pending)[1])  # noqa: E731
            eig = eig + h(gp) - torch.stack([h(O.condition_on_observations(gp, px, py)) for px, py in paths]).mean()
        assert (eig - ref).abs().max() <= 1e-7
        # the host class's float64 forward (fantasy models by fit_with_old_model)
        acq = BAXAcquisitionFunction(model, TopK({"x_path": Xq.numpy(), "k": 4}), exe_paths=len(paths))
        acq.comb_model_list = [acq.fit_with_old_model(px, py) for px, py in paths]
        acq.set_X_pending(pending)
        assert (acq(Xq[2:, None, :]) - ref).abs().max() <= 1e-8


def test_results_layout_round_trip(tmp_path):
    """Row f4: the savetxt layout of src/one_trial.py:344-372 and its restart read-back (:69-113)."""
    from psbax_b200.results import load_trial, save_trial

    folder = str(tmp_path) + "/results/rosenbrock_3/ps_1/"
    inputs = torch.rand(11, 3, dtype=torch.float64)
    obj = torch.rand(11, dtype=torch.float64)
    runtimes = [0.5, 0.25, 0.125]
    metrics = [np.array([0.1]), np.array([0.2]), np.array([0.3]), np.array([0.4])]
    save_trial(folder, 7, inputs, obj, runtimes, metrics)
    for rel in ("inputs/inputs_7.txt", "obj_vals/obj_vals_7.txt", "runtimes/runtimes_7.txt", "performance_metrics_7.txt"):
        assert os.path.exists(folder + rel)
    assert np.loadtxt(folder + "inputs/inputs_7.txt").shape == (11, 3)
    X, Y, rt, pm, it = load_trial(folder, 7)
    assert torch.equal(X, inputs) and torch.equal(Y, obj) and rt == runtimes and it == 3
    assert len(pm) == 4 and all(v.shape == (1,) for v in pm) and float(pm[2][0]) == 0.3
    save_trial(folder, 8, inputs, obj, runtimes, [np.array([0.1, 1.0]), np.array([0.2, 2.0])])
    assert load_trial(folder, 8)[3][1].tolist() == [0.2, 2.0]
    with pytest.raises(OSError):
        load_trial(folder, 9)


def test_subset_select_class_multiplicative_golden(golden):
    """Host greedy on the multiplicative noise model == the reference's class on the same values."""
    for i in range(int(golden["n_sm"])):
        pre = f"sm{i}_"
        f = _fn(golden, pre)
        x, mask, nonneg = golden[pre + "x"], golden[pre + "mask"], bool(golden[pre + "nonneg"])

        class Obj(_Obj):
            noise_type = "multiplicative"

            def get_noisy_f_lst(self, fx):
                v = np.multiply(np.asarray(fx).squeeze(), mask)
                return np.maximum(0, v) if self.nonneg else v

        algo = SubsetSelect({"k": int(golden[pre + "k"])})
        algo.set_obj_func(Obj(x, mask, nonneg))
        exe_path, labels = algo.run_algorithm_on_f(f(x).numpy())
        assert exe_path.idxes == golden[pre + "idxes"].tolist()
        assert list(labels) == golden[pre + "labels"].tolist()
        np.testing.assert_allclose(exe_path.values, golden[pre + "values"], atol=1e-12)


def test_bench_roofline_line_is_consistent():
    """bench.py's roofline object: `frac` = roofline time of the binding pipe / kernel time against the BURST bf16
    peak; `frac_vs_sustained_peak` = the same against MEASURED_PEAKS.json's sustained figure; every workload names
    its binding pipe among tensor / MUFU / FFMA / HBM with the per-candidate counts of SURVEY.md section 8d."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    pk = bench.load_peaks()
    for wl, eng in (("c2", "tensor_bf16x3"), ("c3", "tensor_bf16x3_proj"), ("c4", "tensor_bf16x3_proj"),
                    ("c5", "tensor_bf16x3_proj")):
        w = bench.WORKLOADS[wl]
        r = bench.roofline(wl, eng, 8.0e-3, w["N"], wl == "c2")
        assert r["bound"] in ("tensor", "mufu", "ffma", "hbm")
        assert math.isclose(r["frac"], r["roofline_ms"] / r["kernel_ms"], rel_tol=1e-12)
        assert math.isclose(r["roofline_ms"], max(r["pipes_ms"].values()), rel_tol=1e-12)
        assert math.isclose(r["achieved"], r["frac"] * r["peak"], rel_tol=1e-12)
        if r["bound"] == "tensor" and pk.get("bf16_tflops_sustained"):
            assert math.isclose(r["frac_vs_sustained_peak"],
                                r["frac"] * pk["bf16_tflops"] / pk["bf16_tflops_sustained"], rel_tol=1e-9)
        else:
            assert r["frac_vs_sustained_peak"] is None or r["bound"] == "tensor"
    c2 = bench.roofline("c2", "tensor_bf16x3", 8.0e-3, bench.WORKLOADS["c2"]["N"], True)
    assert c2["algorithmic"]["contraction_flop_per_candidate"] == 2 * (1000 + 64) * 64
    assert c2["bound"] == "tensor"
