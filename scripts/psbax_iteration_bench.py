"""One PS-BAX level-set acquisition at the C2 scale, stage by stage: draw S paths, fused level set over
the grid, union of the masks, fused posterior variance, max-entropy pick (q = 1)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from psbax_b200.acquisition_functions import EntropyAcquisitionFunction, gen_posterior_sampling_batch
from psbax_b200.algorithms import LevelSetEstimator
from psbax_b200.models import SingleTaskGP
from psbax_b200.sample_paths import draw_matheron_paths, posterior_variance

dev = torch.device("cuda:0")
grid = int(os.environ.get("GRID", "4096"))
Xt, y, ls, tau = bench.problem()
model = SingleTaskGP(Xt, y, kernel=bench.KERNEL, lengthscale=ls, outputscale=1.0, noise=bench.NOISE)
a = torch.linspace(0, 1, grid, dtype=torch.float64)
x_set = torch.stack(torch.meshgrid(a, a, indexing="ij"), -1).reshape(-1, 2).numpy()
algo = LevelSetEstimator({"name": "SimpleLevelSet", "threshold": tau, "x_set": x_set})
X = algo.device_candidates(dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


gen = torch.Generator(device=dev).manual_seed(0)
t_draw, paths = timed(lambda: draw_matheron_paths(model, bench.S, bench.L, generator=gen, device=dev))
t_ls, res = timed(lambda: paths.levelset(X, tau))
t_union, pool = timed(lambda: algo.pooled_mask(paths))
t_var, var = timed(lambda: posterior_variance(model, X))
t_pick, pick = timed(lambda: int(torch.where(pool, var, torch.full_like(var, -1.0)).argmax()))
t_all, x_next = timed(lambda: gen_posterior_sampling_batch(model, algo, 1, generator=gen, num_rff_features=bench.L), reps=1)
t_q4, x4 = timed(lambda: gen_posterior_sampling_batch(model, algo, 4, generator=gen, num_rff_features=bench.L), reps=1)
print(f"N = {X.shape[0]} candidates, S = {bench.S} paths, n = {Xt.shape[0]}")
print(f"draw paths {t_draw:.1f} ms | fused level set {t_ls:.1f} ms | pooled mask incl. level set {t_union:.1f} ms ({int(pool.sum())} rows) | "
      f"posterior variance of all N {t_var:.1f} ms | masked arg max {t_pick:.1f} ms")
print(f"gen_posterior_sampling_batch end to end: batch_size 1 (one path, one pick) {t_all:.1f} ms -> {np.asarray(x_next).tolist()}; "
      f"batch_size 4 (four paths, greedy q = 4) {t_q4:.1f} ms")
