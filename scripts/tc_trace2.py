"""Clock trace of the four producer warps of one scheduler (lane quarter 0) of CTA 0 of the register-x tensor
kernel (library built with -DPSBAX_TC_TRACE=2).  Profiling aid."""
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
                            device=dev, engine="auto")
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
print("it | per warp h = 0..3: start  computed(+)  a_empty(+)  stored(+)  arrived(+)")
for it in range(36, 58):
    row = []
    for h in range(4):
        p = [rel(v) for v in t[h, it, :5]]
        row.append(f"{p[0]:7d} +{p[1]-p[0]:5d} +{p[2]-p[1]:4d} +{p[3]-p[2]:4d} +{p[4]-p[3]:4d}" if p[1] > 0 else
                   f"{p[0]:7d}  (kernel blk) .. {p[4]-p[0]:5d}")
    print(f"{it:3d} | " + " | ".join(row))
