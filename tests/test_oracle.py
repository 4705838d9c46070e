"""CPU tests of the oracle itself: golden vectors produced by the reference's own algorithm
classes (tests/golden/make_golden.py) and self-consistency properties of the restated
botorch sampling arithmetic (which the reference cannot pin: see oracle/psbax_oracle.py)."""
import math

import numpy as np
import pytest
import torch

from helpers import GOLDEN, O, make_gp


def _fn(g, pre):
    A, c, w = (torch.from_numpy(g[pre + k]) for k in ("A", "c", "w"))
    return lambda X: torch.cos(torch.as_tensor(X, dtype=torch.float64) @ A + c) @ w


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN, allow_pickle=False)


def test_golden_topk(golden):
    for i in range(int(golden["n_topk"])):
        pre = f"topk{i}_"
        x = torch.from_numpy(golden[pre + "x"])
        f = _fn(golden, pre)
        fx = O.eval_in_batch(f, x)  # chunks of 100, like the reference
        idx, val = O.topk_reduce(fx, int(golden[pre + "k"]))
        assert idx.tolist() == golden[pre + "idx"].tolist()
        np.testing.assert_array_equal(x[idx].numpy(), golden[pre + "out_x"])
        np.testing.assert_allclose(val.numpy().reshape(-1, 1), golden[pre + "out_y"], rtol=0, atol=1e-12)


def test_golden_levelset(golden):
    for i in range(int(golden["n_ls"])):
        pre = f"ls{i}_"
        x = golden[pre + "x"]
        fx = _fn(golden, pre)(x).numpy()
        idx = O.levelset_reduce(fx, float(golden[pre + "thr"]))
        assert idx.tolist() == golden[pre + "idx"].tolist()
        np.testing.assert_array_equal(np.atleast_2d(x[idx]), golden[pre + "out_x"])
        np.testing.assert_allclose(fx[idx], golden[pre + "out_y"], rtol=0, atol=1e-12)


def test_golden_subset_select(golden):
    for i in range(int(golden["n_ss"])):
        pre = f"ss{i}_"
        x = golden[pre + "x"]
        fx = _fn(golden, pre)(x).numpy()
        idxes, values = O.subset_select_reduce(fx, golden[pre + "etas"], int(golden[pre + "k"]),
                                               bool(golden[pre + "nonneg"]))
        assert idxes == golden[pre + "idxes"].tolist()
        np.testing.assert_allclose(values, golden[pre + "values"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(fx[idxes], golden[pre + "out_y"], rtol=0, atol=1e-12)


def test_eval_in_batch_equals_single_call():
    gp = make_gp(12, 3)
    p = O.draw_reference_paths(gp, 2, L=64, generator=torch.Generator().manual_seed(3))
    X = torch.rand(257, 3, dtype=torch.float64)
    f = O.make_callable(p, 1)
    np.testing.assert_allclose(O.eval_in_batch(f, X.unsqueeze(1)).numpy(), O.eval_paths(p, X)[:, 1].numpy(),
                               rtol=0, atol=1e-12)


@pytest.mark.parametrize("kernel", ["rbf", "matern52", "matern32", "matern12"])
def test_rff_gram_matches_exact_kernel(kernel):
    """E[phi(x) phi(x')^T] = k(x, x'): the restated basis (Gamma(nu, nu) scale mixture for
    Matern) has the right spectral density."""
    gen = torch.Generator().manual_seed(0)
    d, L, reps = 3, 1000, 100
    X = torch.rand(24, d, dtype=torch.float64, generator=gen)
    ls = torch.tensor([0.3, 0.5, 0.4], dtype=torch.float64)
    acc = torch.zeros(24, 24, dtype=torch.float64)
    for _ in range(reps):
        w, b = O.rff_basis_draw(d, L, kernel, gen)
        Phi = O.rff_features(X, w, b, ls, 1.3)
        acc += Phi @ Phi.T
    K = O.kernel_matrix(X, X, ls, 1.3, kernel)
    assert (acc / reps - K).abs().max() < 0.03


def test_weight_posterior_moments():
    """w ~ N(m, sigma^2 A^-1): the draw reproduces the Bayesian-linear-regression mean."""
    gp = make_gp(10, 2, noise=1e-2)
    gen = torch.Generator().manual_seed(1)
    w, b = O.rff_basis_draw(2, 50, gp.kernel, gen)
    Phi = O.rff_features(gp.train_X, w, b, gp.lengthscale, gp.outputscale)
    ws = torch.stack([O.weights_posterior_draw(Phi, gp.train_targets, gp.noise, gen) for _ in range(400)])
    _, m, cov = O.weights_posterior_draw(Phi, gp.train_targets, gp.noise, gen, return_moments=True)
    assert (ws.mean(0) - m).abs().max() < 5 * cov.diagonal().max().sqrt() / math.sqrt(400) + 1e-3


@pytest.mark.parametrize("scheme", ["reference", "matheron"])
def test_samples_interpolate_data_when_noise_vanishes(scheme):
    gp = make_gp(8, 2, noise=1e-6, ls=torch.tensor([0.4, 0.5]))
    gen = torch.Generator().manual_seed(5)
    draw = O.draw_reference_paths if scheme == "reference" else O.draw_matheron_paths
    p = draw(gp, 3, L=400, generator=gen)
    f = O.eval_paths(p, gp.train_X)
    assert (f - gp.train_Y[:, None]).abs().max() < 2e-2 * gp.y_std


def test_matheron_mean_is_posterior_mean():
    gp = make_gp(16, 2, noise=1e-2)
    gen = torch.Generator().manual_seed(7)
    X = torch.rand(40, 2, dtype=torch.float64, generator=gen)
    p = O.draw_matheron_paths(gp, 600, L=256, generator=gen)
    mean_paths = O.eval_paths(p, X).mean(1)
    mu = O.eval_paths(O.posterior_mean_params(gp), X)[:, 0]
    post_mean, post_cov = O.gp_posterior(gp, X)
    np.testing.assert_allclose(mu.numpy(), post_mean.numpy(), atol=1e-9)
    sd = post_cov.diagonal().clamp_min(0).sqrt()
    assert ((mean_paths - mu).abs() < 6 * sd / math.sqrt(600) + 0.05 * gp.y_std).all()


def test_grad_matches_finite_difference():
    from helpers import random_params

    p = random_params(3, 40, 4, 9, kernel="matern52", seed=2)
    X = torch.rand(5, 4, dtype=torch.float64)
    g = O.eval_path_grad(p, X, 1)
    h = 1e-6
    for i in range(4):
        e = torch.zeros(4, dtype=torch.float64)
        e[i] = h
        fd = (O.eval_paths(p, X + e)[:, 1] - O.eval_paths(p, X - e)[:, 1]) / (2 * h)
        np.testing.assert_allclose(g[:, i].numpy(), fd.numpy(), atol=1e-5)


def test_mesh_helpers_match_reference_layout():
    xx = O.get_mesh(2, 4)
    X = O.reshape_mesh(xx)
    assert X.shape == (16, 2)
    # 'ij' order: the first coordinate varies slowest (src/utils.py:273-286)
    assert torch.allclose(X[:4, 0], torch.zeros(4, dtype=torch.float64))
    assert torch.allclose(X[:4, 1], torch.linspace(0, 1, 4, dtype=torch.float64))


def test_golden_subset_select_multiplicative(golden):
    """The multiplicative noise model (src/problems.py:150-160) through the reference's own greedy."""
    assert int(golden["n_sm"]) >= 2
    for i in range(int(golden["n_sm"])):
        pre = f"sm{i}_"
        x = torch.from_numpy(golden[pre + "x"])
        fx = (torch.cos(x @ torch.from_numpy(golden[pre + "A"]) + torch.from_numpy(golden[pre + "c"]))
              @ torch.from_numpy(golden[pre + "w"])).numpy()
        idxes, values = O.subset_select_reduce(fx, golden[pre + "mask"], int(golden[pre + "k"]),
                                               bool(golden[pre + "nonneg"]), multiplicative=True)
        assert idxes == golden[pre + "idxes"].tolist()
        np.testing.assert_allclose(values, golden[pre + "values"], atol=1e-12)
