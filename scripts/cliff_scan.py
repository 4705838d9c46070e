"""Scan shape transitions (engine hand-overs) for performance cliffs: level set over N = 1e6 candidates."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev

N = 1_000_000


def timeit(fn, reps=3):
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


print("S     d    L     n :   level set ms | top-10 ms | values ms")
for S in (1, 2, 64, 65, 256):
    for d in (1, 2, 4, 5, 8, 16, 17, 32, 33, 64, 128):
        for (L, n) in ((1000, 64), (1000, 0), (0, 64)):
            if S in (2, 65) and (L, n) != (1000, 64):
                continue
            if d in (1, 8, 17, 64) and (L, n) != (1000, 64):
                continue
            p = make(S, L, d, n, "auto")
            X = torch.rand(N, d, generator=torch.Generator().manual_seed(1)).to(dev)
            a = timeit(lambda: p.levelset(X, 0.0))
            b = timeit(lambda: p.topk(X, 10))
            c = timeit(lambda: p.values(X)) if S <= 64 else float("nan")
            print(f"{S:4d} {d:4d} {L:5d} {n:4d} : {a:10.3f} {b:10.3f} {c:10.3f}", flush=True)
