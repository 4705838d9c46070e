"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle on the same
seeded float32-representable inputs.

Tolerance (BASELINE.json north_star: "sample values within 1e-4 relative in fp32, index sets
identical wherever the score gap exceeds that tolerance"): per sample
    tol_s = 1e-4 * y_std * max(1, ||W_s||_2)                      (oracle.path_scale)
i.e. relative to the magnitude of the path — no ||V_s|| term.  The exact
bench problem (Matheron draw, noise 1e-4) has its own test in tests/test_gpu_bench_parity.py.
"""
import math

import numpy as np
import pytest
import torch

from helpers import O, candidates, make_gp, random_params, to_batch, unpack_mask

pytestmark = pytest.mark.gpu

REL = 1e-4

# (S, L, d, n, G, kernel, N)
CASES = [
    (64, 1000, 2, 64, 1, "matern52", 4096),     # C2 shape (level-set grid)
    (64, 1000, 2, 64, 1, "rbf", 1000),          # ragged N
    (32, 256, 10, 122, 1, "matern52", 3001),    # C5 shape (Ackley-10D), generic-d path
    (256, 200, 5, 128, 1, "matern32", 1500),    # several sample chunks
    (16, 128, 3, 0, 1, "matern52", 777),        # no kernel term
    (20, 96, 4, 17, 1, "matern12", 513),        # S not a multiple of anything
    (8, 1000, 3, 0, 8, "matern52", 2048),       # reference form: one basis per sample
    (64, 120, 2, 0, 64, "rbf", 300),            # reference form, full chunk
    (5, 64, 6, 9, 5, "matern52", 129),          # per-sample basis + kernel term, d=6
    (1, 1000, 3, 0, 1, "matern52", 100),        # C1 demo shape: one path, 100 points
    (1, 0, 3, 38, 1, "matern52", 100),          # posterior mean: L = 0
    (3, 40, 80, 12, 1, "rbf", 260),             # d = 80 (GB1 one-hot width), small-S route
    (24, 72, 80, 16, 1, "matern52", 200),       # d = 80 on the shared-basis route
    (70, 1000, 32, 512, 1, "matern52", 1500),   # C3 shape (32-d embedding), generic-d tensor kernel
    (64, 300, 80, 130, 1, "rbf", 700),          # d = 80, generic-d tensor kernel
    (2, 8, 1, 3, 2, "matern32", 33),            # tiny everything
    (64, 0, 32, 32, 1, "linear", 1000),         # DKGP / LinearKernel path (a2): h = 32 embedding, L = 0
    (256, 0, 8, 8, 1, "linear", 700),           # linear rows only, x in registers
    (20, 40, 5, 7, 1, "linear", 300),           # cos rows + linear rows in one table
    (3, 16, 5, 7, 1, "linear", 200),            # small-S route
    (4, 16, 5, 7, 4, "linear", 200),            # per-sample bases + linear rows
    (64, 500, 31, 40, 1, "matern52", 1300),     # tensor-projection variant, bias column in the last K slot
    (64, 200, 24, 30, 1, "rbf", 900),           # projection variant, d % 8 == 0 (bias added by the producers)
    (128, 333, 16, 0, 1, "rbf", 2049),          # projection variant, no kernel term, 2 chunks, ragged
    (40, 64, 7, 200, 1, "matern32", 515),       # projection variant, kernel rows dominate
    (64, 0, 2, 40, 1, "matern52", 1000),        # exact-kernel rows only (posterior-variance pseudo-paths), shared basis
    (130, 0, 10, 100, 1, "rbf", 700),           # the same, d = 10, three sample chunks
]


def _tol(p, X=None):
    return (REL * O.path_scale(p, X)).numpy()


def _tensor_ok(c):
    S, L, d, n, G, kernel, N = c
    return G == 1 and (L > 0 or n > 0) and d <= 128


def _proj_ok(c):
    return _tensor_ok(c) and 5 <= c[2] <= 32


ENGINE_CASES = ([(c, "simt") for c in CASES] + [(c, "tensor") for c in CASES if _tensor_ok(c)] +
                [(c, "tensor_bf16") for c in CASES if _tensor_ok(c)] +
                [(c, "tensor_proj") for c in CASES if _proj_ok(c)])
_REF_CACHE = {}


@pytest.fixture(scope="module", params=ENGINE_CASES,
                ids=lambda ce: "S%d_L%d_d%d_n%d_G%d_%s_N%d" % ce[0] + "_" + ce[1])
def case(request):
    (S, L, d, n, G, kernel, N), engine = request.param
    p = random_params(S, L, d, n, G=G, kernel=kernel, seed=S + L + d)
    X = candidates(N, d)
    key = request.param[0]
    if key not in _REF_CACHE:
        _REF_CACHE[key] = O.eval_paths(p, X.double()).numpy()  # [N, S] float64
    try:
        batch = to_batch(p, engine=engine)
    except Exception as ex:  # e.g. 3xTF32 operands of a d = 80 problem do not fit shared memory
        if "engine needs" in str(ex):
            pytest.skip(f"{engine}: shape not supported by this engine")
        raise
    return p, X, _REF_CACHE[key], batch


def test_values(case):
    p, X, ref, batch = case
    out = batch.values(X.cuda()).double().cpu().numpy()
    assert out.shape == ref.shape
    err = np.abs(out - ref).max(axis=0)
    assert (err <= _tol(p, X)).all(), (err / _tol(p, X)).max()


def test_sumsq(case):
    """Row-wise sum of squared path values (psbax_eval_sumsq) against the oracle's value matrix."""
    p, X, ref, batch = case
    got = batch.sumsq(X.cuda()).double().cpu().numpy()
    want = (ref ** 2).sum(axis=1)
    # |sum (y + e)^2 - sum y^2| <= sum (2 |y| e + e^2) with |e_s| <= tol_s
    tol = _tol(p, X)
    bound = (2 * np.abs(ref) * tol[None, :] + tol[None, :] ** 2).sum(axis=1)
    assert got.shape == want.shape
    assert (np.abs(got - want) <= bound + 1e-6 * np.abs(want)).all()


def test_levelset(case):
    p, X, ref, batch = case
    tol = _tol(p, X)
    for tau in (float(np.quantile(ref, 0.55)), float(ref.max() + 10.0), float(ref.min() - 10.0)):
        res = batch.levelset(X.cuda(), tau)
        mask = unpack_mask(res.mask, X.shape[0])  # [S, N]
        want = (ref > tau).T
        decided = (np.abs(ref - tau) > tol[None, :]).T
        assert (mask[decided] == want[decided]).all()
        count = res.count.cpu().numpy()
        assert (count == mask.sum(1)).all()
        lo = (want & decided).sum(1)
        hi = lo + (~decided).sum(1)
        assert ((count >= lo) & (count <= hi)).all()
        # arg max (the empty-set fallback of levelset.py:35-37)
        am = res.argmax.cpu().numpy()
        mv = res.maxval.double().cpu().numpy()
        best = ref.max(axis=0)
        assert (np.abs(mv - best) <= tol).all()
        assert (np.abs(ref[am, np.arange(ref.shape[1])] - best) <= 2 * tol).all()
        for s in range(ref.shape[1]):
            idx = res.indices(s).cpu().numpy()
            if count[s] == 0:
                assert idx.tolist() == [am[s]]
            else:
                assert (idx == np.nonzero(mask[s])[0]).all()


def test_argmax_only(case):
    p, X, ref, batch = case
    mv, am = batch.argmax(X.cuda(), row_offset=1000)
    tol = _tol(p, X)
    best = ref.max(axis=0)
    assert (np.abs(mv.double().cpu().numpy() - best) <= tol).all()
    am = am.cpu().numpy() - 1000
    assert (np.abs(ref[am, np.arange(ref.shape[1])] - best) <= 2 * tol).all()


@pytest.mark.parametrize("k", [1, 4, 10, 64])
def test_topk(case, k):
    p, X, ref, batch = case
    N, S = ref.shape
    if k > N:
        pytest.skip("k > N")
    tol = _tol(p, X)
    val, idx = batch.topk(X.cuda(), k)
    val, idx = val.double().cpu().numpy(), idx.cpu().numpy()
    for s in range(S):
        oidx, oval = O.topk_reduce(torch.from_numpy(ref[:, s]), k)
        oidx, oval = oidx.numpy(), oval.numpy()
        assert len(set(idx[s].tolist())) == k
        # values agree position by position, and are the oracle's values at the chosen rows
        assert (np.abs(val[s] - oval) <= tol[s]).all()
        assert (np.abs(ref[idx[s], s] - val[s]) <= tol[s]).all()
        assert (np.diff(val[s]) <= 0).all()
        # index sets identical wherever the gap to the first excluded value exceeds tolerance
        order = np.sort(ref[:, s])[::-1]
        kth, nxt = order[k - 1], (order[k] if k < N else -np.inf)
        if kth - nxt > 2 * tol[s]:
            assert set(idx[s].tolist()) == set(oidx.tolist())
        sure = oidx[oval - nxt > 2 * tol[s]]
        assert set(sure.tolist()) <= set(idx[s].tolist())


@pytest.mark.parametrize("engine", ["simt", "tensor", "tensor_bf16"])
def test_ties_break_towards_smaller_index(engine):
    p = random_params(16, 64, 2, 8, seed=3)
    base = candidates(50, 2)
    X = torch.cat([base, base, base])  # every value appears three times, bit-identically
    batch = to_batch(p, engine=engine)
    val, idx = batch.topk(X.cuda(), 6)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for s in range(16):
        assert (val[s, 0] == val[s, 1]) and (val[s, 1] == val[s, 2])
        assert idx[s, 0] + 50 == idx[s, 1] and idx[s, 1] + 50 == idx[s, 2] and idx[s, 0] < 50
        assert idx[s, 3] + 50 == idx[s, 4] and idx[s, 4] + 50 == idx[s, 5]
    _, am = batch.argmax(X.cuda())
    assert (am.cpu().numpy() == idx[:, 0]).all()


@pytest.mark.parametrize("engine", ["simt", "tensor", "tensor_bf16"])
def test_row_offset_and_sharded_merge_equal_whole(engine):
    """Shards of the candidate set + psbax_topk_merge == one launch over the whole set, bit for bit."""
    import ctypes as C

    from psbax_b200 import _lib

    p = random_params(64, 200, 3, 20, seed=9)
    X = candidates(5000, 3).cuda()
    batch = to_batch(p, engine=engine)
    k = 10
    val, idx = batch.topk(X, k)
    cuts = [0, 1300, 1301, 4000, 5000]
    vs, is_ = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        kk = min(k, b - a)
        v, i = batch.topk(X[a:b], kk, row_offset=a)
        pad_v = torch.full((64, k), -float("inf"), device="cuda")
        pad_i = torch.full((64, k), -1, dtype=torch.int64, device="cuda")
        pad_v[:, :kk], pad_i[:, :kk] = v, i
        vs.append(pad_v)
        is_.append(pad_i)
    lv, li = torch.stack(vs).contiguous(), torch.stack(is_).contiguous()
    ov = torch.empty(64, k, device="cuda")
    oi = torch.empty(64, k, dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().psbax_topk_merge(lv.data_ptr(), li.data_ptr(), len(vs), 64, k, ov.data_ptr(),
                                            oi.data_ptr(), torch.cuda.current_stream().cuda_stream), "merge")
    assert torch.equal(oi, idx) and torch.equal(ov, val)
    # level set: shard masks concatenate to the whole mask, counts add up
    tau = 0.0
    whole = batch.levelset(X, tau)
    a = batch.levelset(X[:2048], tau)
    b = batch.levelset(X[2048:], tau, row_offset=2048)
    assert torch.equal(a.count + b.count, whole.count)
    assert torch.equal(torch.cat([a.mask, b.mask], 1), whole.mask)
    gm = torch.where(a.maxval >= b.maxval, a.argmax, b.argmax)
    assert torch.equal(gm, whole.argmax)


def test_grad_matches_oracle():
    for (S, L, d, n, G, kernel) in [(4, 300, 3, 10, 1, "matern52"), (3, 100, 10, 20, 3, "rbf"),
                                    (20, 64, 2, 8, 1, "matern32"), (20, 16, 6, 6, 1, "linear"),
                                    (2, 0, 6, 6, 1, "linear")]:
        p = random_params(S, L, d, n, G=G, kernel=kernel, seed=4)
        batch = to_batch(p)
        X = candidates(40, d)
        sidx = torch.arange(40) % S
        val, g = batch.grad(X.cuda(), sidx)
        ref = O.eval_paths(p, X.double())
        for s in range(S):
            rows = (sidx == s).nonzero().reshape(-1)
            og = O.eval_path_grad(p, X[rows].double(), s).numpy()
            gnorm = max(1.0, float(np.abs(og).max()))
            assert np.abs(g[rows].double().cpu().numpy() - og).max() <= 2e-4 * gnorm * float(O.path_scale(p, X)[s])
            assert np.abs(val[rows].double().cpu().numpy() - ref[rows, s].numpy()).max() <= _tol(p, X)[s]


def test_subset_select_matches_oracle():
    # (samples, candidates, noise draws, k, nonneg): every samples-per-CTA variant (1, 2, 4, 8), draw counts
    # that are not multiples of the register block (4), fewer candidates than threads
    for (S, N, B, k, nonneg) in [(3, 300, 20, 5, False), (40, 500, 16, 10, True), (700, 120, 8, 4, False),
                                 (3, 300, 13, 5, False), (9, 5, 3, 3, True), (1300, 90, 6, 3, True),
                                 (300, 64, 5, 2, False), (200, 1100, 7, 3, True)]:
        p = random_params(S, 64, 5, 12, seed=S)
        X = candidates(N, 5)
        g = torch.Generator().manual_seed(2)
        etas = 0.7 * torch.randn(B, N, generator=g)
        batch = to_batch(p)
        idx = batch.subset_select(X.cuda(), etas, k, nonneg).cpu().numpy()
        fx = batch.values(X.cuda()).double().cpu().numpy()  # same values the GPU greedy saw
        for s in range(0, S, max(1, S // 7)):
            oidx, values = O.subset_select_reduce(fx[:, s], etas.double().numpy(), k, nonneg)
            if idx[s].tolist() == oidx:
                continue
            # float32 vs float64 means may swap near-ties: the greedy objective must still agree
            def score(sel):
                return np.max(values[:, sel], axis=1).mean()
            assert abs(score(idx[s].tolist()) - score(oidx)) <= 1e-5 * max(1.0, abs(score(oidx)))


def test_single_path_callable_protocol():
    p = random_params(3, 128, 3, 6, seed=5)
    batch = to_batch(p)
    X = candidates(250, 3).double()
    ref = O.eval_paths(p, X).numpy()
    f = batch.path(1)
    y = torch.cat([f(c) for c in X.unsqueeze(1).split(100)])  # eval_in_batch, topk.py:48-49
    assert y.dtype == torch.float64 and y.device.type == "cpu" and y.shape == (250,)
    assert np.abs(y.numpy() - ref[:, 1]).max() <= _tol(p)[1]
    assert f(X[:7]).shape == (7,)
    Xg = X[:5].clone().requires_grad_(True)
    f(Xg).sum().backward()
    og = O.eval_path_grad(p, X[:5], 1).numpy()
    assert np.abs(Xg.grad.numpy() - og).max() <= 2e-4 * max(1.0, np.abs(og).max()) * float(O.path_scale(p)[1])


def test_errors_are_reported():
    from psbax_b200 import PsbaxError

    p = random_params(4, 32, 2, 0, seed=1)
    batch = to_batch(p)
    X = candidates(10, 2).cuda()
    with pytest.raises(PsbaxError):
        batch.topk(X, 11)  # k > N, torch.topk raises too
    with pytest.raises(PsbaxError):
        batch.topk(X, 65)
    with pytest.raises(PsbaxError):
        batch.values(candidates(10, 3).cuda())


@pytest.mark.parametrize("engine", ["tensor", "tensor_bf16", "tensor_proj"])
def test_tensor_engine_many_tiles_matches_simt(engine):
    """Several tiles per persistent CTA (accumulator double-buffering, barrier phase wrap-around)
    and several sample chunks: the tcgen05 engines against the FFMA engine and the oracle."""
    d = 10 if engine == "tensor_proj" else 2
    p = random_params(100, 1000, d, 64, seed=21)
    N = 148 * 128 * 5 + 77
    X = candidates(N, d).cuda()
    bt, bs = to_batch(p, engine=engine), to_batch(p, engine="simt")
    vt, vs = bt.values(X), bs.values(X)
    tol = torch.from_numpy(_tol(p)).cuda().float()
    assert ((vt - vs).abs().max(0).values <= tol).all()
    rows = torch.randperm(N)[:2048]
    ref = O.eval_paths(p, X[rows].double().cpu()).numpy()
    assert (np.abs(vt[rows].double().cpu().numpy() - ref).max(axis=0) <= _tol(p)).all()
    tau = float(np.quantile(ref, 0.55))
    rt, rs = bt.levelset(X, tau), bs.levelset(X, tau)
    mt, ms = unpack_mask(rt.mask, N), unpack_mask(rs.mask, N)
    undecided = ((vs - tau).abs() <= tol[None, :]).T.cpu().numpy()
    assert ((mt != ms) <= undecided).all()
    assert (rt.count.cpu().numpy() == mt.sum(1)).all()
    tv, ti = bt.topk(X, 10)
    sv, si = bs.topk(X, 10)
    assert ((tv - sv).abs() <= tol[:, None]).all()
    # index sets identical wherever the gap between the 10th and 11th value exceeds tolerance
    top11 = torch.topk(vs, 11, dim=0).values  # [11, S]
    clear = (top11[9] - top11[10]) > 2 * tol
    same_set = torch.tensor([set(ti[s].tolist()) == set(si[s].tolist()) for s in range(ti.shape[0])]).cuda()
    assert (same_set | ~clear).all()


@pytest.mark.parametrize("engine", ["simt", "tensor", "tensor_bf16", "tensor_proj"])
def test_large_feature_arguments(engine):
    """Short lengthscales: |omega . x + b| up to a few hundred.  The projection error (fp32 FFMA, or the
    3xTF32 products of the projection variant) grows with the argument; the contract still holds."""
    p = random_params(32, 256, 16, 24, seed=77)
    p.omega = p.omega * 4.0
    X = candidates(1500, 16)
    ref = O.eval_paths(p, X.double()).numpy()
    out = to_batch(p, engine=engine).values(X.cuda()).double().cpu().numpy()
    err = np.abs(out - ref).max(axis=0)
    assert (err <= _tol(p, X)).all(), (engine, float((err / _tol(p, X)).max()))


@pytest.mark.parametrize("engine,d", [("simt", 2), ("tensor", 2), ("tensor_bf16", 2), ("tensor_bf16", 40), ("tensor_proj", 10)])
def test_single_candidate(engine, d):
    """N = 1: one valid row in the first tile of the only pass."""
    p = random_params(64, 100, d, 20, seed=9)
    X = candidates(1, d)
    ref = O.eval_paths(p, X.double()).numpy()  # [1, S]
    b = to_batch(p, engine=engine)
    tol = _tol(p, X)
    assert (np.abs(b.values(X.cuda()).double().cpu().numpy() - ref) <= tol[None, :]).all()
    res = b.levelset(X.cuda(), float(np.median(ref)))
    mask = unpack_mask(res.mask, 1)
    decided = np.abs(ref[0] - np.median(ref)) > tol
    assert (mask[:, 0][decided] == (ref[0] > np.median(ref))[decided]).all()
    assert (res.argmax.cpu().numpy() == 0).all()
    val, idx = b.topk(X.cuda(), 1)
    assert (idx.cpu().numpy() == 0).all() and (np.abs(val[:, 0].double().cpu().numpy() - ref[0]) <= tol).all()
    assert np.abs(b.sumsq(X.cuda()).double().cpu().numpy()[0] - (ref[0] ** 2).sum()) <= (2 * np.abs(ref[0]) * tol + tol ** 2).sum() + 1e-6


def test_levelset_streamed_equals_resident():
    """Host-resident candidates streamed in chunks == one launch over the resident set."""
    p = random_params(64, 200, 2, 16, seed=31)
    batch = to_batch(p)
    X = candidates(50_000, 2)
    tau = 0.3
    whole = batch.levelset(X.cuda(), tau)
    masks, bounds, cnt, bv, bi = batch.levelset_streamed(X.pin_memory(), tau, chunk_rows=8192)
    assert bounds[0] == 0 and bounds[-1] == 50_000 and len(masks) == len(bounds) - 1
    ref_mask = unpack_mask(whole.mask, 50_000)
    got = np.concatenate([unpack_mask(m, bounds[c + 1] - bounds[c]) for c, m in enumerate(masks)], axis=1)
    assert (got == ref_mask).all()
    assert torch.equal(cnt, whole.count.cpu()) and torch.equal(bv, whole.maxval.cpu())
    assert torch.equal(bi, whole.argmax.cpu())


def test_randomised_shapes_all_engines():
    """Seeded sweep over random shapes (dims, feature counts, sample counts, ragged N, both basis
    modes, every kernel): values / level set / top-k of every engine that supports the shape against
    the oracle."""
    rng = np.random.default_rng(2024)
    kernels = ["rbf", "matern12", "matern32", "matern52"]
    for trial in range(24):
        d = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 33]))
        L = int(rng.choice([0, 8, 37, 100, 256, 1000]))
        n = int(rng.choice([0, 1, 5, 64, 130]))
        if L + n == 0:
            n = 3
        S = int(rng.choice([1, 2, 3, 15, 16, 17, 63, 64, 65, 130]))
        G = 1 if (L == 0 or rng.random() < 0.6) else S
        N = int(rng.choice([1, 31, 32, 33, 127, 128, 129, 255, 257, 1000, 4097]))
        kernel = kernels[trial % 4]
        p = random_params(S, L, d, n, G=G, kernel=kernel, seed=100 + trial)
        X = candidates(N, d, seed=trial)
        ref = O.eval_paths(p, X.double()).numpy()
        tol = _tol(p)
        engines = ["simt"]
        if G == 1 and (L > 0 or n > 0):
            engines += ["tensor", "tensor_bf16"]
            if 5 <= d <= 32:
                engines.append("tensor_proj")
        for eng in engines:
            try:
                b = to_batch(p, engine=eng)
            except Exception as ex:
                if "engine needs" in str(ex):
                    continue
                raise
            tag = (trial, eng, S, L, d, n, G, kernel, N)
            out = b.values(X.cuda()).double().cpu().numpy()
            assert (np.abs(out - ref).max(axis=0) <= tol).all(), tag
            tau = float(np.median(ref))
            res = b.levelset(X.cuda(), tau)
            mask = unpack_mask(res.mask, N)
            decided = (np.abs(ref - tau) > tol[None, :]).T
            assert (mask[decided] == (ref > tau).T[decided]).all(), tag
            assert (res.count.cpu().numpy() == mask.sum(1)).all(), tag
            assert (np.abs(res.maxval.double().cpu().numpy() - ref.max(0)) <= tol).all(), tag
            k = min(N, int(rng.choice([1, 2, 10, 33, 64])))
            val, idx = b.topk(X.cuda(), k)
            val, idx = val.double().cpu().numpy(), idx.cpu().numpy()
            srt = -np.sort(-ref, axis=0)[:k].T  # [S, k]
            assert (np.abs(val - srt) <= tol[:, None]).all(), tag
            assert (np.abs(np.take_along_axis(ref.T, idx, 1) - val) <= tol[:, None]).all(), tag


def test_subset_select_multiplicative_and_per_sample_noise():
    """psbax_subset_select_ex: the multiplicative noise model (values = fx * mask, src/problems.py:150-160)
    with one mask per sample path (the reference draws a fresh one per get_noisy_f_lst call), and additive
    noise with per-sample matrices, against the oracle's greedy on the same values."""
    for (S, N, B, k, nonneg, noise) in [(5, 300, 12, 5, False, "multiplicative"), (33, 130, 7, 4, True, "multiplicative"),
                                        (4, 200, 9, 3, False, "additive")]:
        p = random_params(S, 64, 5, 12, seed=S + 3)
        X = candidates(N, 5)
        g = torch.Generator().manual_seed(7)
        if noise == "multiplicative":
            etas = (torch.rand(S, B, N, generator=g) < 0.6).float()
        else:
            etas = 0.7 * torch.randn(S, B, N, generator=g)
        batch = to_batch(p)
        idx = batch.subset_select(X.cuda(), etas, k, nonneg, noise=noise).cpu().numpy()
        shared = batch.subset_select(X.cuda(), etas[0], k, nonneg, noise=noise).cpu().numpy()
        fx = batch.values(X.cuda()).double().cpu().numpy()
        for s in range(S):
            for got, e in ((idx[s], etas[s]), (shared[s], etas[0])):
                oidx, values = O.subset_select_reduce(fx[:, s], e.double().numpy(), k, nonneg,
                                                      multiplicative=noise == "multiplicative")
                if got.tolist() == oidx:
                    continue

                def score(sel):
                    return np.max(values[:, sel], axis=1).mean()
                assert abs(score(got.tolist()) - score(oidx)) <= 1e-5 * max(1.0, abs(score(oidx)))
