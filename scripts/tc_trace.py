"""Dump the clock64 pipeline trace of CTA 0 of the tensor-engine kernel (profiling aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from psbax_b200 import _lib
from psbax_b200.models import SingleTaskGP
from psbax_b200.sample_paths import draw_matheron_paths

dev = torch.device("cuda:0")
Xt, y, ls, tau = bench.problem()
model = SingleTaskGP(Xt, y, kernel=bench.KERNEL, lengthscale=ls, outputscale=1.0, noise=bench.NOISE)
paths = draw_matheron_paths(model, bench.S, bench.L, generator=torch.Generator(device=dev).manual_seed(0),
                            device=dev, engine=os.environ.get("PSBAX_ENGINE", "tensor"))
grid = 1024
a = torch.linspace(0, 1, grid, dtype=torch.float64)
X = torch.stack(torch.meshgrid(a, a, indexing="ij"), -1).reshape(-1, 2).float().to(dev)
paths.levelset(X, tau)
torch.cuda.synchronize()
buf = torch.zeros(4 * 128 * 8, dtype=torch.int64, device=dev)
_lib.load().psbax_debug_trace_buffer(buf.data_ptr())
paths.levelset(X, tau)
torch.cuda.synchronize()
_lib.load().psbax_debug_trace_buffer(None)
t = buf.cpu().reshape(4, 128, 8)
t0 = int(t[t > 0].min())
rel = lambda v: int(v) - t0 if v > 0 else -1  # noqa: E731
print("MMA    it:  enter  a_full    done | producer(w4): start computed a_empty stored arrived")
for it in range(40, 60):
    m = [rel(v) for v in t[0, it, :3]]
    p = [rel(v) for v in t[1, it, :5]]
    print(f"{it:4d}: " + " ".join(f"{v:8d}" for v in m) + "   | " + " ".join(f"{v:8d}" for v in p))
print("epilogue passes (enter, d_full, d_empty_arrive, done):")
for tl in range(6):
    print("   ", [rel(v) for v in t[2, tl, :4]])
