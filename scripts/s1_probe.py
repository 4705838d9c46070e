"""One sample path (the reference's batch_size = 1 call) over large candidate sets: time against the MUFU bound."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev


def timeit(fn, reps=5):
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


for (S, N, d, n) in [(1, 1_000_000, 2, 64), (1, 16_777_216, 2, 64), (1, 10_000_000, 10, 122), (1, 149_361, 32, 512),
                     (4, 1_000_000, 2, 64), (8, 1_000_000, 2, 64), (15, 1_000_000, 2, 64), (16, 1_000_000, 2, 64)]:
    p = make(S, 1000, d, n, "auto")
    X = torch.rand(N, d, generator=torch.Generator().manual_seed(1)).to(dev)
    ms = timeit(lambda: p.levelset(X, 0.0))
    cos = N * (1000 + 2 * n)  # MUFU ops: cos per RFF feature, sqrt + ex2 per exact-kernel feature (computed once per row)
    bound = cos / (15.9 * 148 * 1.965e9) * 1e3
    print(f"S={S:3d} N={N:9d} d={d:2d} n={n:3d}: {ms:8.3f} ms  (MUFU bound for one basis {bound:.3f} ms)", flush=True)
