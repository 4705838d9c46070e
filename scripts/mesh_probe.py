"""Device time of the level-set step by input / output form (c2 shape): uploaded rows vs mesh
descriptor, full masks vs pooled OR-mask only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
paths, tau, _ = bench.draw_paths("c2", dev, "auto")
X = bench.grid_slab(0, 1, dev)
a0, a1 = (a.to(dev) for a in bench.grid_axes(1))
N = X.shape[0]


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


variants = {
    "rows, masks": lambda: paths.levelset(X, tau, want_mask=True),
    "rows, pooled only": lambda: paths.levelset(X, tau, want_mask=False, want_pooled=True),
    "rows, nothing": lambda: paths.levelset(X, tau, want_mask=False),
    "mesh, masks": lambda: paths.levelset(None, tau, want_mask=True, mesh=[a0, a1], N=N),
    "mesh, pooled only": lambda: paths.levelset(None, tau, want_mask=False, want_pooled=True, mesh=[a0, a1], N=N),
    "mesh, masks + pooled": lambda: paths.levelset(None, tau, want_mask=True, want_pooled=True, mesh=[a0, a1], N=N),
}
acc = {k: [] for k in variants}
for rnd in range(4):  # interleaved: the board is power-capped and drifts over a run
    for k, fn in variants.items():
        acc[k].append(t(fn, reps=4))
for k, v in acc.items():
    print("%-22s %s  min %.3f ms" % (k, " ".join("%.3f" % x for x in v), min(v)))
