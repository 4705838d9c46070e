"""MUFU.COS throughput probe: the per-sample-basis kernel (G = S) does L cos per evaluation and
little else, so its evals/s * L bounds the achievable MUFU.COS rate.  Profiling aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import SamplePathBatch

dev = torch.device("cuda:0")
S, L, d, N = 64, 1000, 2, 1 << 20
g = torch.Generator().manual_seed(0)
omega = 6 * torch.randn(S, L, d, generator=g)
bias = 6.28 * torch.rand(S, L, generator=g)
W = torch.randn(S, L, generator=g) * (2.0 / L) ** 0.5
paths = SamplePathBatch(omega, bias, W, "matern52", device=dev)
X = torch.rand(N, d, generator=g).to(dev)
for _ in range(2):
    paths.levelset(X, 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    paths.levelset(X, 0.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
cos_per_s = N * S * L / (ms * 1e-3)
print(f"per-sample kernel: {ms:.2f} ms, {N * S / ms / 1e6:.2f} Gevals/s, {cos_per_s / 1e12:.3f} Tcos/s "
      f"= {cos_per_s / 148 / 1.965e9:.2f} cos/clk/SM (at 1965 MHz)")
