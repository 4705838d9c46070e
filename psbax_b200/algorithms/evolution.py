"""Evolution-strategy local optimiser scored on S sample paths at once (BASELINE.json configs[4]:
"local BO Ackley-10D: evolution-strategy populations scored on S sample paths").

The reference ships no ES class (its local-BO algorithm is ``LBFGSB``,
``/root/reference/src/algorithms/lbfgsb.py``); the INFO-BAX code base it derives from has
``EvolutionStrategies`` (params ``n_generation``, ``n_population``, ``samp_str``, ``init_x``,
``domain``, ``normal_scale``, ``keep_frac``, ``crop``), whose interface this class keeps under the
reference's ``Algorithm`` protocol.  One generation = ONE fused launch: the whole population is
scored on all S paths and each path keeps its own top-mu rows (``psbax_eval_topk``: the N x S score
matrix never exists).  Every path then moves its own search distribution (mean / per-dimension spread
of its elite rows); the next population is the union of the paths' offspring, so any path may pick
any other path's offspring.
"""
from argparse import Namespace

import torch

from ..sample_paths import SamplePath, SamplePathBatch
from .algorithm import Algorithm


class EvolutionStrategies(Algorithm):
    def set_params(self, params):
        super().set_params(params)
        self.params.name = params.get("name", "EvolutionStrategies")
        self.params.n_dim = params.get("n_dim", None)
        self.params.n_generation = params.get("n_generation", 15)
        self.params.n_population = params.get("n_population", 4096)  # rows scored per generation (all paths)
        self.params.keep_frac = params.get("keep_frac", None)  # elite fraction of a path's share; or n_elite
        self.params.n_elite = params.get("n_elite", 16)
        self.params.normal_scale = params.get("normal_scale", 0.15)  # initial spread, in domain units
        self.params.min_scale = params.get("min_scale", 1e-4)
        self.params.domain = params.get("domain", None)  # [[lo, hi]] * d; default the unit cube
        self.params.init_x = params.get("init_x", None)  # optional [m, d] rows seeded into generation 0
        self.params.crop = params.get("crop", True)
        self.params.seed = params.get("seed", None)

    def initialize(self):
        self.exe_path = Namespace(x=[], y=[])

    def run_algorithm_on_batch(self, paths: SamplePathBatch):
        """-> list over the S paths of (exe_path{x [G * mu, d], y [G * mu]}, best x [1, d])."""
        p, dev = self.params, paths.device
        S, d = paths.S, p.n_dim or paths.d
        dom = torch.as_tensor(p.domain if p.domain is not None else [[0.0, 1.0]] * d, dtype=torch.float32, device=dev)
        lo, hi = dom[:, 0], dom[:, 1]
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(p.seed) if p.seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,))))
        per = max(1, p.n_population // S)  # offspring per path and generation
        mu = min(p.n_elite if p.keep_frac is None else max(1, int(p.keep_frac * per)), 64, per * S)
        pop = lo + (hi - lo) * torch.rand(per * S, d, device=dev, generator=gen)
        if p.init_x is not None:
            pop = torch.cat([torch.as_tensor(p.init_x).to(dev, torch.float32).reshape(-1, d), pop])
        xs, ys = [], []
        best_v = torch.full((S,), -float("inf"), device=dev)
        best_x = torch.zeros(S, d, device=dev)
        for g in range(p.n_generation):
            val, idx = paths.topk(pop, mu)  # ONE fused launch: every row scored on every path
            elite = pop[idx]  # [S, mu, d]
            xs.append(elite)
            ys.append(val)
            better = val[:, 0] > best_v
            best_v = torch.where(better, val[:, 0], best_v)
            best_x = torch.where(better[:, None], elite[:, 0], best_x)
            if g == p.n_generation - 1:
                break
            mean = elite.mean(1)
            spread = elite.std(1, unbiased=False) if mu > 1 else torch.zeros_like(mean)
            floor = max(p.min_scale, p.normal_scale * 0.7 ** (g + 1))  # keeps exploring while the elites agree
            sigma = torch.maximum(spread, (hi - lo) * floor)
            child = mean[:, None, :] + sigma[:, None, :] * torch.randn(S, per, d, device=dev, generator=gen)
            if p.crop:
                child = torch.minimum(torch.maximum(child, lo), hi)
            pop = torch.cat([elite.reshape(-1, d), child.reshape(-1, d)])  # elitism: the best rows stay candidates
        X, Y = torch.cat(xs, 1).double().cpu(), torch.cat(ys, 1).double().cpu()
        bx = best_x.double().cpu()
        out = []
        for s in range(S):
            ep = Namespace(x=X[s].numpy(), y=Y[s].numpy())
            out.append((ep, bx[s].reshape(1, d).numpy()))
        return out

    def run_algorithm_on_f(self, f):
        self.initialize()
        if isinstance(f, SamplePath):
            f = f._one()
        if not isinstance(f, SamplePathBatch):
            raise TypeError("EvolutionStrategies scores populations on a SamplePath / SamplePathBatch")
        self.exe_path, out = self.run_algorithm_on_batch(f)[0]
        return self.exe_path, out

    def get_output(self):
        return self.exe_path.x[:1]
