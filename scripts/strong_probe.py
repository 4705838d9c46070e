"""Why is the 1-GPU strong-scaling leg (3163 x 3163 mesh, mask + pooled) slow?  Timing probe."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda:0")
paths, tau, _ = bench.draw_paths("c2", dev, "auto")


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for steps in (3163, 3168, 4096):
    ax = torch.linspace(0, 1, steps, dtype=torch.float64).to(torch.float32).to(dev)
    N = steps * steps
    X = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2).contiguous()
    for name, kw in (("mask+pooled", dict(want_mask=True, want_pooled=True)), ("mask", dict(want_mask=True)),
                     ("pooled", dict(want_mask=False, want_pooled=True)), ("none", dict(want_mask=False))):
        a = t(lambda: paths.levelset(None, tau, mesh=[ax, ax], N=N, **kw))
        b = t(lambda: paths.levelset(X, tau, **kw))
        print(f"steps {steps} N {N} {name:12s} mesh {a:7.3f} ms   rows {b:7.3f} ms", flush=True)
