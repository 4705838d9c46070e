"""Timing probe: level-set pass over a 2-D grid (config-2 shape) for one engine.
usage: tc_probe.py [grid] [engine] [reps].  Profiling aid, not a benchmark."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from psbax_b200.models import SingleTaskGP
from psbax_b200.sample_paths import draw_matheron_paths

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
engine = sys.argv[2] if len(sys.argv) > 2 else "tensor"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
Xt, y, ls, tau = bench.problem()
model = SingleTaskGP(Xt, y, kernel=bench.KERNEL, lengthscale=ls, outputscale=1.0, noise=bench.NOISE)
paths = draw_matheron_paths(model, bench.S, bench.L, generator=torch.Generator(device=dev).manual_seed(0),
                            device=dev, engine=engine)
a = torch.linspace(0, 1, grid, dtype=torch.float64)
X = torch.stack(torch.meshgrid(a, a, indexing="ij"), -1).reshape(-1, 2).float().to(dev)
N = X.shape[0]
for _ in range(2):
    r = paths.levelset(X, tau)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    r = paths.levelset(X, tau)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
clk_per_row = ms * 1e-3 * 148 * 1.965e9 / N  # at the nominal 1965 MHz
print(f"engine={engine} N={N} {ms:8.3f} ms  "
      f"{N * bench.S / ms / 1e6:8.2f} Gevals/s  {clk_per_row:7.1f} clk/row/SM  count0={int(r.count[0])}")
