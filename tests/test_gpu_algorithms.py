"""GPU tests of the reference-facing API: algorithm classes and the PS-BAX acquisition driving
the fused kernels, against the oracle's restatement of the reference flow."""
import math
import types

import numpy as np
import pytest
import torch

from helpers import O, make_gp, random_params, to_batch

pytestmark = pytest.mark.gpu


def _model_from(gp):
    from psbax_b200.models import SingleTaskGP

    return SingleTaskGP(gp.train_X, gp.train_Y, kernel=gp.kernel, lengthscale=gp.lengthscale,
                        outputscale=gp.outputscale, noise=gp.noise, mean_const=gp.mean_const)


def _oracle_params(batch):
    t = lambda v: None if v is None else v.double().cpu()  # noqa: E731
    return O.PathParams(omega=t(batch.omega), bias=t(batch.bias), W=t(batch.W), kernel=batch.kernel,
                        Xtrain=t(batch.Xtrain), inv_ls=t(batch.inv_ls), V=t(batch.V),
                        mean_const=float(np.float32(batch.mean_const)), y_mean=float(np.float32(batch.y_mean)),
                        y_std=float(np.float32(batch.y_std)))


@pytest.mark.parametrize("sampler", ["matheron", "reference"])
def test_topk_algorithm_on_drawn_paths(sampler):
    """demos/topk.py shape: 100 random points in 3-D, k = 4."""
    from psbax_b200.algorithms import TopK
    from psbax_b200.utils import generate_random_points, get_function_sample_batch

    gp = make_gp(8, 3, noise=1e-3)
    model = _model_from(gp)
    x_path = generate_random_points(100, 3, seed=1234).double().numpy()
    algo = TopK({"x_path": x_path, "k": 4})
    gen = torch.Generator(device="cuda").manual_seed(0)
    paths = get_function_sample_batch(model, 6, sampler=sampler, generator=gen, num_rff_features=256)
    p = _oracle_params(paths)
    ref = O.eval_paths(p, torch.from_numpy(x_path))
    tol = 1e-4 * O.path_scale(p)
    results = algo.run_algorithm_on_batch(paths)
    for s, (exe_path, out) in enumerate(results):
        oidx, oval = O.topk_reduce(ref[:, s], 4)
        order = torch.sort(ref[:, s], descending=True).values
        if order[3] - order[4] > 2 * tol[s]:
            assert set(exe_path.idx) == set(oidx.tolist())
        assert out.shape == (4, 3) and exe_path.y.shape == (4, 1)
        np.testing.assert_array_equal(out.numpy(), x_path[exe_path.idx])
        assert (exe_path.y[:, 0] - oval).abs().max() <= tol[s]
    # single-path protocol gives the same answer as the batch
    e1, o1 = algo.get_copy().run_algorithm_on_f(paths.path(2))
    assert e1.idx == results[2][0].idx


def test_levelset_algorithm_on_drawn_paths():
    from psbax_b200.algorithms import LevelSetEstimator
    from psbax_b200.utils import get_function_sample_batch, get_mesh, reshape_mesh

    gp = make_gp(12, 2, noise=1e-3)
    model = _model_from(gp)
    x_set = reshape_mesh(get_mesh(2, 50)).double().numpy()  # demos/level-set.py: 50 x 50 grid
    thr = float(np.quantile(gp.train_Y.numpy(), 0.55))
    algo = LevelSetEstimator({"name": "SimpleLevelSet", "threshold": thr, "x_set": x_set})
    gen = torch.Generator(device="cuda").manual_seed(1)
    paths = get_function_sample_batch(model, 5, generator=gen, num_rff_features=512)
    p = _oracle_params(paths)
    ref = O.eval_paths(p, torch.from_numpy(x_set)).numpy()
    tol = (1e-4 * O.path_scale(p)).numpy()
    for s, (exe_path, out) in enumerate(algo.run_algorithm_on_batch(paths)):
        want = set(O.levelset_reduce(ref[:, s], thr).tolist())
        got = set(int(i) for i in exe_path.idx)
        unsure = set(np.nonzero(np.abs(ref[:, s] - thr) <= tol[s])[0].tolist())
        assert (want ^ got) <= unsure
        assert out.shape == (len(got), 2)
        assert np.abs(exe_path.y - ref[sorted(got), s]).max() <= tol[s]
    # threshold above everything -> [argmax]
    algo2 = LevelSetEstimator({"threshold": 1e9, "x_set": x_set})
    e2, o2 = algo2.run_algorithm_on_f(paths.path(0))
    assert len(e2.idx) == 1 and o2.shape == (1, 2)
    assert abs(ref[int(e2.idx[0]), 0] - ref[:, 0].max()) <= 2 * tol[0]


def test_subset_select_algorithm_on_paths():
    from psbax_b200.algorithms import SubsetSelect

    N, B, k = 400, 12, 5
    p = random_params(6, 96, 5, 10, seed=8)
    paths = to_batch(p)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 5, generator=g, dtype=torch.float64)
    etas = 0.5 * torch.randn(B, N, generator=g, dtype=torch.float64).numpy()

    class Obj:
        noise_type, nonneg = "additive", False
        df = types.SimpleNamespace(index=[f"gene{i}" for i in range(N)])

        def __init__(self):
            self.etas = etas

        def get_x(self, idx=None):
            return x

        def get_noisy_f_lst(self, fx):
            return fx.squeeze() + self.etas

    algo = SubsetSelect({"k": k})
    algo.set_obj_func(Obj())
    results = algo.run_algorithm_on_batch(paths)
    fx = paths.values(x).double().cpu().numpy()
    for s, (exe_path, labels) in enumerate(results):
        oidx, values = O.subset_select_reduce(fx[:, s], etas, k)
        assert labels == [f"gene{i}" for i in exe_path.idxes]
        score = lambda sel: np.max(values[:, sel], axis=1).mean()  # noqa: E731
        assert exe_path.idxes == oidx or abs(score(exe_path.idxes) - score(oidx)) < 1e-5


def test_ps_bax_subset_select_branch():
    """``gen_posterior_sampling_batch`` with SubsetSelect (``posterior_sampling.py:49-64``): ceil(batch / k)
    paths, their selected labels pooled, max-entropy pick over the pooled rows, labels returned."""
    from psbax_b200.acquisition_functions import gen_posterior_sampling_batch
    from psbax_b200.algorithms import SubsetSelect

    N, B, k = 300, 10, 4
    gp = make_gp(10, 5, noise=1e-3, seed=6)
    model = _model_from(gp)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, 5, generator=g, dtype=torch.float64)
    etas = 0.5 * torch.randn(B, N, generator=g, dtype=torch.float64).numpy()
    labels = [f"gene{i}" for i in range(N)]

    class Obj:
        noise_type, nonneg = "additive", False
        df = types.SimpleNamespace(index=labels)

        def __init__(self):
            self.etas = etas
            self.x_to_idx = {tuple(r.tolist()): labels[i] for i, r in enumerate(x)}

        def get_x(self, idx=None):
            return x if idx is None else x[[labels.index(i) for i in idx]]

        def get_noisy_f_lst(self, fx):
            return fx.squeeze() + self.etas

        def get_idx_from_x(self, X):
            return [self.x_to_idx[tuple(r.tolist())] for r in X]

    algo = SubsetSelect({"k": k})
    algo.set_obj_func(Obj())
    picked = gen_posterior_sampling_batch(model, algo, 6, generator=torch.Generator(device="cuda").manual_seed(2),
                                          num_rff_features=128)
    assert len(picked) == 6 and len(set(picked)) == 6 and all(lbl in labels for lbl in picked)


def test_ps_bax_acquisition_end_to_end():
    from psbax_b200.acquisition_functions import gen_posterior_sampling_batch
    from psbax_b200.algorithms import TopK
    from psbax_b200.utils import generate_random_points

    gp = make_gp(8, 3, noise=1e-3)
    model = _model_from(gp)
    x_path = generate_random_points(100, 3, seed=1234).double()
    algo = TopK({"x_path": x_path.numpy(), "k": 4})
    gen = torch.Generator(device="cuda").manual_seed(3)
    x_next = gen_posterior_sampling_batch(model, algo.get_copy(), 2, generator=gen, num_rff_features=128)
    assert x_next.shape == (2, 3)
    rows = {tuple(r) for r in x_path.numpy().tolist()}
    assert all(tuple(r) in rows for r in x_next.numpy().tolist())


@pytest.mark.parametrize("n,d,kernel,noise", [(8, 3, "matern52", 1e-3), (40, 2, "matern52", 1e-4),
                                              (64, 5, "rbf", 1e-3), (130, 10, "matern32", 1e-2)])
def test_fused_posterior_variance_and_entropy_pick(n, d, kernel, noise):
    """Row f1: the max-entropy pick as one fused launch (psbax_eval_sumsq over alpha L^-1 pseudo-paths)
    against the oracle's exact posterior (float64): variances, q = 1 entropy values, and the greedy
    q = 3 picks of optimize_acqf_discrete (conditioning on the pending rows)."""
    from psbax_b200.acquisition_functions import EntropyAcquisitionFunction
    from psbax_b200.acquisition_functions.entropy import optimize_acqf_discrete
    from psbax_b200.sample_paths import posterior_variance

    gp = make_gp(n, d, kernel=kernel, noise=noise, seed=n)
    model = _model_from(gp)
    g = torch.Generator().manual_seed(5)
    choices = torch.rand(700, d, dtype=torch.float64, generator=g)
    ref_var = torch.stack([O.gp_posterior(gp, c[None])[1].reshape(()) for c in choices])
    scale = float(ref_var.max())
    var = posterior_variance(model, choices).double().cpu()
    assert (var - ref_var).abs().max() <= 2e-4 * scale  # 3xTF32 engine where it applies, else BF16x3
    acq = EntropyAcquisitionFunction(model=model)
    vals = acq.q1_values(choices)
    ref_vals = O.entropy_acqf(gp, choices[:, None, :])
    big = ref_var > 0.05 * scale  # the entropy is only meaningful (and only ever picked) away from the data
    assert (vals.cpu()[big] - ref_vals[big]).abs().max() <= 5e-3
    assert int(vals.argmax()) == int(ref_vals.argmax())
    # greedy q = 3: fused GPU steps == the reference's chunked CPU evaluation == the oracle
    x_f, v_f = optimize_acqf_discrete(acq, 3, choices, fused=True)
    x_c, v_c = optimize_acqf_discrete(acq, 3, choices, fused=False)
    ox, _ = O.optimize_acqf_discrete(gp, choices, 3)
    assert torch.equal(x_c, ox)
    assert torch.equal(x_f.cpu(), x_c)
    assert (v_f.cpu() - v_c).abs().max() <= 5e-3
    assert acq.X_pending is None


def test_ps_bax_levelset_union_pick_equals_reference_pool(monkeypatch):
    """PS-BAX on a level-set problem: the fused pick over the de-duplicated union of the bit masks ==
    the reference flow (pool every path's coordinate rows with duplicates, chunks of 100 through the
    exact posterior on the CPU)."""
    import psbax_b200.acquisition_functions.posterior_sampling as ps
    from psbax_b200.algorithms import LevelSetEstimator
    from psbax_b200.utils import get_mesh, reshape_mesh

    gp = make_gp(12, 2, noise=1e-3, seed=4)
    model = _model_from(gp)
    x_set = reshape_mesh(get_mesh(2, 24)).numpy()
    for thr, q in [(float(np.quantile(gp.train_Y.numpy(), 0.5)), 2), (1e9, 1)]:  # 1e9: empty sets -> arg max rows
        algo = LevelSetEstimator({"name": "SimpleLevelSet", "threshold": thr, "x_set": x_set})
        fused = ps.gen_posterior_sampling_batch(model, algo.get_copy(), q, num_rff_features=128,
                                                generator=torch.Generator(device="cuda").manual_seed(11))
        with monkeypatch.context() as m:
            m.setattr(ps, "_fusable", lambda acq: False)
            pooled = ps.gen_posterior_sampling_batch(model, algo.get_copy(), q, num_rff_features=128,
                                                     generator=torch.Generator(device="cuda").manual_seed(11))
        assert fused.shape == (q, 2)
        assert torch.equal(torch.as_tensor(fused).double().cpu(), torch.as_tensor(pooled).double().cpu())


def test_info_bax_eig_fused_equals_torch_forward():
    """Row f3: the INFO-BAX expected information gain for every candidate through S + 1 fused variance
    launches == the plain float64 forward (current model and S fantasy models conditioned on the
    execution paths), including the greedy q = 2 pick with a pending row."""
    from psbax_b200.acquisition_functions import BAXAcquisitionFunction
    from psbax_b200.acquisition_functions.entropy import optimize_acqf_discrete
    from psbax_b200.algorithms import TopK
    from psbax_b200.utils import generate_random_points

    gp = make_gp(10, 3, noise=1e-3, seed=2)
    model = _model_from(gp)
    x_path = generate_random_points(150, 3, seed=1234).double()
    acq = BAXAcquisitionFunction(model, TopK({"x_path": x_path.numpy(), "k": 4}), exe_paths=5)
    acq.initialize(generator=torch.Generator(device="cuda").manual_seed(1), num_rff_features=128)
    ref = torch.cat([acq(c.unsqueeze(1)) for c in x_path.split(50)])
    # the float64 oracle restatement of forward / acq_exe_normal on the same execution paths
    exe = [(torch.as_tensor(np.asarray(ep.x)).double(), torch.as_tensor(np.asarray(ep.y)).double().reshape(-1))
           for ep in acq.exe_path_list]
    oref = O.bax_eig(gp, exe, x_path[:, None, :])
    assert (ref - oref).abs().max() <= 1e-8
    got = acq.q1_values(x_path).cpu()  # ONE launch: current model + all fantasy models
    per_model = acq._q1_values_per_model(x_path).cpu()  # S + 1 launches
    informative = oref > oref.max() - 3.0  # away from rows whose conditional variance is ~ noise
    assert (got[informative] - oref[informative]).abs().max() <= 5e-3
    assert (per_model[informative] - oref[informative]).abs().max() <= 5e-3
    assert int(got.argmax()) == int(oref.argmax())
    acq.set_X_pending(x_path[int(oref.argmax())][None])
    opend = O.bax_eig(gp, exe, x_path[:, None, :], pending=x_path[int(oref.argmax())][None])
    gpend = acq.q1_values(x_path).cpu()
    finite = torch.isfinite(opend)  # the pending row itself has a singular joint covariance
    inf2 = finite & (opend > opend[finite].max() - 3.0)
    gpend = torch.where(finite, gpend, torch.full_like(gpend, -1e30))
    opend = torch.where(finite, opend, torch.full_like(opend, -1e30))
    assert (gpend[inf2] - opend[inf2]).abs().max() <= 5e-3 and int(gpend.argmax()) == int(opend.argmax())
    acq.set_X_pending(None)
    x_f, _ = optimize_acqf_discrete(acq, 2, x_path, fused=True)
    x_c, _ = optimize_acqf_discrete(acq, 2, x_path, fused=False)
    assert torch.equal(x_f.cpu(), x_c)


def test_lbfgsb_improves_on_raw_samples():
    from psbax_b200.algorithms import LBFGSB

    p = random_params(1, 200, 3, 10, seed=12)
    paths = to_batch(p)
    torch.manual_seed(0)
    algo = LBFGSB({"n_dim": 3, "raw_samples": 64, "num_restarts": 4})
    exe_path, x = algo.run_algorithm_on_f(paths.path(0))
    assert x.shape == (1, 3) and (x >= 0).all() and (x <= 1).all()
    grid = torch.rand(2000, 3, dtype=torch.float64)
    assert exe_path.y >= float(O.eval_paths(p, grid).max()) - 0.05 * float(O.path_scale(p)[0])
    assert abs(float(O.eval_paths(p, torch.from_numpy(x))[0, 0]) - exe_path.y) < 1e-3


def test_posterior_mean_function():
    from psbax_b200.utils import posterior_mean_function

    gp = make_gp(20, 3, noise=1e-2)
    f = posterior_mean_function(_model_from(gp))
    X = torch.rand(64, 3, dtype=torch.float64)
    mean, _ = O.gp_posterior(gp, X)
    assert (f(X.unsqueeze(1)) - mean).abs().max() < 1e-3 * gp.y_std


def test_dkgp_linear_paths_topk():
    """DKGP / LinearKernel branch (src/utils.py:153-206): exact weight-space draw on a network
    embedding; the GPU path (embedding by the torch network, evaluation + top-k fused) against the
    oracle's restatement evaluated on the same embedding and the same drawn weights."""
    from psbax_b200.algorithms import TopK
    from psbax_b200.models import DKGP
    from psbax_b200.utils import get_function_sample_batch

    torch.manual_seed(0)
    g = torch.Generator().manual_seed(3)
    X = torch.rand(40, 12, dtype=torch.float64, generator=g)
    y = torch.sin(X.sum(-1))
    y = (y - y.mean()) / y.std()  # fit_model.py:76 standardizes before building the DKGP
    model = DKGP(X, y, architecture=[12, 16, 8, 1])
    model.train_model(X, y, lr=1e-2, num_iter=20)
    paths = get_function_sample_batch(model, 24, generator=torch.Generator(device="cuda").manual_seed(1))
    assert paths.kernel == "linear" and paths.L == 0 and paths.n == 8 and paths.d == 8
    cand = torch.rand(500, 12, dtype=torch.float64, generator=g)
    with torch.no_grad():
        emb = model.embedding(cand).double().cpu()
    p = O.PathParams(omega=torch.zeros(1, 0, 8, dtype=torch.float64), bias=torch.zeros(1, 0, dtype=torch.float64),
                     W=torch.zeros(24, 0, dtype=torch.float64), kernel="linear", Xtrain=paths.Xtrain.double().cpu(),
                     inv_ls=paths.inv_ls.double().cpu(), V=paths.V.double().cpu(),
                     mean_const=float(np.float32(paths.mean_const)), y_mean=0.0, y_std=1.0)
    ref = O.eval_paths(p, emb)  # [500, 24]
    tol = 1e-4 * O.path_scale(p, emb) + 1e-5 * emb.abs().max() * p.V.abs().sum(1)  # embedding is fp32 on the GPU
    vals = paths.values(cand).double().cpu()
    assert ((vals - ref).abs().max(0).values <= tol).all()
    algo = TopK({"x_path": cand.numpy(), "k": 5})
    for s, (exe_path, out) in enumerate(algo.run_algorithm_on_batch(paths)):
        order = torch.sort(ref[:, s], descending=True).values
        if order[4] - order[5] > 2 * tol[s]:
            assert set(exe_path.idx) == set(O.topk_reduce(ref[:, s], 5)[0].tolist())
        assert out.shape == (5, 12)
    f = paths.path(3)
    assert (f(cand[:7].unsqueeze(1)) - ref[:7, 3]).abs().max() <= tol[3]


def test_dkgp_ps_bax_pick_fused_equals_torch(monkeypatch):
    """C3 flow (DKGP + TopK + posterior sampling): the max-entropy pick over the pooled top-k rows with the
    deep-kernel variance e(x)^T Sn e(x) as one fused launch == the float64 torch posterior, greedy q = 3."""
    import psbax_b200.acquisition_functions.posterior_sampling as ps
    from psbax_b200.acquisition_functions import EntropyAcquisitionFunction
    from psbax_b200.algorithms import TopK
    from psbax_b200.models import DKGP

    torch.manual_seed(0)
    g = torch.Generator().manual_seed(3)
    X = torch.rand(40, 12, dtype=torch.float64, generator=g)
    y = torch.sin(X.sum(-1))
    y = (y - y.mean()) / y.std()
    model = DKGP(X, y, architecture=[12, 16, 8, 1])
    model.train_model(X, y, lr=1e-2, num_iter=20)
    cand = torch.rand(400, 12, dtype=torch.float64, generator=g)
    # weight-space posterior == exact GP with kernel v <e, e'> on the embedding (float64)
    with torch.no_grad():
        E, Eq = model.embedding(X).double().cpu(), model.embedding(cand[:5]).double().cpu()
    v, s2 = float(model.covar_module.base_kernel.variance), float(model.likelihood.noise[0])
    Kinv = torch.linalg.inv(v * E @ E.T + s2 * torch.eye(40, dtype=torch.float64))
    cov_ref = v * Eq @ Eq.T - v * Eq @ E.T @ Kinv @ E @ Eq.T * v
    assert (model.posterior(cand[:5]).covariance_matrix.cpu() - cov_ref).abs().max() <= 1e-8
    acq = EntropyAcquisitionFunction(model=model)
    ref = torch.cat([acq(c.unsqueeze(1)) for c in cand.split(100)]).cpu()
    got = acq.q1_values(cand).cpu()
    assert (got - ref).abs().max() <= 2e-3 and int(got.argmax()) == int(ref.argmax())
    algo = TopK({"x_path": cand.numpy(), "k": 5})
    fused = ps.gen_posterior_sampling_batch(model, algo.get_copy(), 3,
                                            generator=torch.Generator(device="cuda").manual_seed(11))
    with monkeypatch.context() as m:
        import psbax_b200.acquisition_functions.entropy as ent

        m.setattr(ent, "_fusable", lambda acq_function: False)
        plain = ps.gen_posterior_sampling_batch(model, algo.get_copy(), 3,
                                                generator=torch.Generator(device="cuda").manual_seed(11))
    assert fused.shape == (3, 12) and torch.equal(fused.double().cpu(), plain.double().cpu())


def test_info_bax_initialize_runs_the_batch_path():
    """BAXAcquisitionFunction.initialize / run_algorithm_on_f_list (bax_acquisition.py:221-277):
    exe_paths sample paths in one fused launch; the EIG of a q = 1 candidate is finite and the
    conditioned models hold old data + path."""
    from psbax_b200.acquisition_functions import BAXAcquisitionFunction
    from psbax_b200.acquisition_functions.entropy import optimize_acqf_discrete
    from psbax_b200.algorithms import TopK
    from psbax_b200.utils import generate_random_points

    gp = make_gp(8, 3, noise=1e-3)
    model = _model_from(gp)
    x_path = generate_random_points(100, 3, seed=1234).double()
    algo = TopK({"x_path": x_path.numpy(), "k": 4})
    acq = BAXAcquisitionFunction(model, algo, exe_paths=7)
    acq.initialize(generator=torch.Generator(device="cuda").manual_seed(0), num_rff_features=128)
    assert len(acq.exe_path_list) == 7 and len(acq.output_list) == 7 and len(acq.comb_model_list) == 7
    assert all(o.shape == (4, 3) for o in acq.output_list)
    assert acq.comb_model_list[0].train_inputs[0].shape == (12, 3)
    vals = acq(x_path[:20].unsqueeze(1))
    assert vals.shape == (20,) and torch.isfinite(vals).all() and (vals > -1e-6).all()
    x_next, _ = optimize_acqf_discrete(acq, 2, x_path, max_batch_size=50)
    assert x_next.shape == (2, 3)


def test_batched_ascent_against_scipy_on_the_oracle():
    """Row f4: psbax_ascend runs restarts x samples box-constrained ascents in one launch.  Checked against the
    float64 oracle: the reported value is the path value at the returned point, every ascent improves on its
    start, ends at a (projected-)stationary point, and the best restart per sample is as good as scipy's
    L-BFGS-B driven by the oracle's own value / gradient from the same starts."""
    from scipy.optimize import minimize

    S, d, R = 5, 3, 6
    p = random_params(S, 200, d, 10, seed=21)
    p.omega = (p.omega / 3.0).float().double()  # smoother paths: fewer local maxima between the two optimisers
    paths = to_batch(p)
    g = torch.Generator().manual_seed(4)
    X0 = torch.rand(S * R, d, generator=g)
    sidx = torch.arange(S).repeat_interleave(R)
    x, f, it = paths.ascend(X0, sidx, maxiter=100)
    x, f = x.double().cpu(), f.double().cpu()
    assert (x >= 0).all() and (x <= 1).all() and int(it.max()) <= 100
    scale = O.path_scale(p)
    best_dev = torch.full((S,), -1e30, dtype=torch.float64)
    best_ref = torch.full((S,), -1e30, dtype=torch.float64)
    for i in range(S * R):
        s = int(sidx[i])
        fo = float(O.eval_paths(p, x[i][None])[0, s])
        assert abs(fo - float(f[i])) <= 1e-4 * float(scale[s])  # f_out is f_s(x_out)
        assert fo >= float(O.eval_paths(p, X0[i][None].double())[0, s]) - 1e-4 * float(scale[s])
        gr = O.eval_path_grad(p, x[i][None], s)[0]
        free = ~(((x[i] <= 0) & (gr < 0)) | ((x[i] >= 1) & (gr > 0)))
        assert float((gr * free).abs().max()) <= 2e-2 * float(scale[s])  # stationary on the free variables

        def neg(xv, s=s):
            xt = torch.from_numpy(xv)[None]
            return -float(O.eval_paths(p, xt)[0, s]), -O.eval_path_grad(p, xt, s)[0].numpy()

        res = minimize(neg, X0[i].double().numpy(), jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * d,
                       options={"maxiter": 100})
        best_ref[s] = max(best_ref[s], -res.fun)
        best_dev[s] = max(best_dev[s], fo)
    assert (best_dev >= best_ref - 2e-3 * scale).all()


def test_lbfgsb_batch_equals_per_path_protocol_and_scipy():
    from psbax_b200.algorithms import LBFGSB

    p = random_params(4, 200, 3, 10, seed=12)
    p.omega = (p.omega / 3.0).float().double()
    paths = to_batch(p)
    torch.manual_seed(0)
    dev = LBFGSB({"n_dim": 3, "raw_samples": 128, "num_restarts": 8}).run_algorithm_on_batch(paths)
    torch.manual_seed(0)
    ref = LBFGSB({"n_dim": 3, "raw_samples": 128, "num_restarts": 8, "optimizer": "scipy"}).run_algorithm_on_batch(paths)
    scale = O.path_scale(p)
    for s, ((ep, x), (ep_ref, _)) in enumerate(zip(dev, ref)):
        assert x.shape == (1, 3) and (x >= 0).all() and (x <= 1).all()
        assert abs(float(O.eval_paths(p, torch.from_numpy(x))[0, s]) - ep.y) <= 1e-4 * float(scale[s])
        assert ep.y >= ep_ref.y - 2e-3 * float(scale[s])


def test_evolution_strategies_scores_populations_on_all_paths():
    """C5 / row f4: every generation is one fused top-mu launch over the shared population for all S paths;
    the best row per path must beat a dense random search scored by the oracle and be a row whose oracle
    value is the reported one."""
    from psbax_b200.algorithms import EvolutionStrategies

    S, d = 6, 4
    p = random_params(S, 128, d, 12, seed=31)
    p.omega = (p.omega / 3.0).float().double()
    paths = to_batch(p)
    algo = EvolutionStrategies({"n_dim": d, "n_generation": 12, "n_population": 6000, "n_elite": 16, "seed": 3})
    res = algo.run_algorithm_on_batch(paths)
    scale = O.path_scale(p)
    rand = O.eval_paths(p, torch.rand(20000, d, dtype=torch.float64, generator=torch.Generator().manual_seed(9)))
    for s, (ep, x) in enumerate(res):
        assert ep.x.shape == (12 * 16, d) and ep.y.shape == (12 * 16,)
        assert (ep.x >= 0).all() and (ep.x <= 1).all()
        fo = O.eval_paths(p, torch.from_numpy(ep.x))[:, s].numpy()
        assert np.abs(fo - ep.y).max() <= 1e-4 * float(scale[s])  # the kept values are the path values of the kept rows
        fbest = float(O.eval_paths(p, torch.from_numpy(x))[0, s])
        assert fbest >= float(rand[:, s].max()) - 1e-3 * float(scale[s])
        assert abs(fbest - ep.y.max()) <= 1e-4 * float(scale[s])
    one, _ = EvolutionStrategies({"n_dim": d, "n_generation": 3, "n_population": 512, "seed": 1}).run_algorithm_on_f(paths.path(2))
    assert one.x.shape[1] == d


def test_mlp_embedding_kernel_matches_torch():
    """Row a2: the FFNN embedding (deep_kernel_gp.py:18-84) as psbax_mlp_forward, float32, against the
    float64 torch network; ragged N, every activation, the GB1 architecture 80 -> 128 -> 64 -> 32."""
    from psbax_b200.models import FFNN
    from psbax_b200.sample_paths import DeviceFFNN

    torch.manual_seed(0)
    for arch, act, N in [([80, 128, 64, 32], "leaky_relu", 1000), ([12, 32, 16], "tanh", 77), ([5, 7], "relu", 3),
                         ([33, 200, 9, 250, 31], "swish", 513), ([8, 16, 4], "sigmoid", 64)]:
        net = FFNN(arch, act)
        X = torch.randn(N, arch[0], dtype=torch.float64)
        with torch.no_grad():
            ref = net(X)
        got = DeviceFFNN(net, "cuda")(X).double().cpu()
        assert next(net.parameters()).device.type == "cpu"  # the caller's module is not moved
        assert (got - ref).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max())), (arch, act)


def test_model_list_gp_paths_are_per_output_batches():
    """Row a3: ModelListGP branch (src/utils.py:231-248): one batch of paths per output model, values
    concatenated on the last dim; each output equals the oracle's evaluation of that output's parameters."""
    from psbax_b200.models import ModelListGP
    from psbax_b200.utils import get_function_sample_batch, get_function_samples

    gps = [make_gp(9, 3, noise=1e-3, seed=1), make_gp(9, 3, kernel="rbf", noise=1e-2, seed=1)]
    gps[1].train_Y = (gps[1].train_Y * 2.0 - 1.0)
    gps[1].__post_init__()
    model = ModelListGP(*[_model_from(g) for g in gps])
    mb = get_function_sample_batch(model, 5, generator=torch.Generator(device="cuda").manual_seed(0), num_rff_features=128)
    X = torch.rand(300, 3, dtype=torch.float64)
    vals = mb.values(X)
    assert vals.shape == (300, 5, 2)
    for j in range(2):
        p = _oracle_params(mb.output(j))
        ref = O.eval_paths(p, X)
        assert ((vals[..., j].double().cpu() - ref).abs().max(0).values <= 1e-4 * O.path_scale(p)).all()
    f = mb.path(3)
    out = f(X[:50].unsqueeze(1))
    assert out.shape == (50, 1, 2)
    assert (out[:, 0, :].cpu() - vals[:50, 3, :].double().cpu()).abs().max() <= 1e-4 * 4
    assert f.posterior(X[:7].unsqueeze(1)).mean.shape == (7, 1, 2)
    single = get_function_samples(model, generator=torch.Generator(device="cuda").manual_seed(0), num_rff_features=64)
    assert single(X[:4].unsqueeze(1)).shape == (4, 1, 2)


def test_dkgp_nonlinear_kernel_paths():
    """Row a3: DKGP with a Matern kernel on the embedding (src/utils.py:207-221): RFF / Matheron paths of
    the GP over the embedded training set, evaluated on embedding(X) — against the oracle fed the float64
    torch embedding."""
    from psbax_b200.models import DKGP
    from psbax_b200.utils import get_function_sample_batch

    torch.manual_seed(1)
    X = torch.rand(30, 6, dtype=torch.float64)
    y = torch.sin(3 * X.sum(-1))
    y = (y - y.mean()) / y.std()
    model = DKGP(X, y, [6, 16, 4, 1], kernel="matern52")
    model.train_model(X, y, lr=1e-2, num_iter=20)
    assert model.covar_module.base_kernel.lengthscale.shape == (1, 4)
    paths = get_function_sample_batch(model, 4, generator=torch.Generator(device="cuda").manual_seed(2), num_rff_features=128)
    Xq = torch.rand(200, 6, dtype=torch.float64)
    with torch.no_grad():
        Eq = model.embedding(Xq)
        Et = model.embedding(X)
    p = _oracle_params(paths)
    ref = O.eval_paths(p, Eq)
    got = paths.values(Xq).double().cpu()
    assert ((got - ref).abs().max(0).values <= 2e-4 * O.path_scale(p)).all()  # fp32 embedding + fp32 evaluator
    # the paths interpolate the (standardized) data up to the noise level
    at_train = O.eval_paths(p, Et)
    assert (at_train - y[:, None]).abs().max() <= 6 * math.sqrt(float(model.likelihood.noise))
