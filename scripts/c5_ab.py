"""C5-shape timing only (argmax S = 32 and top-32 S = 64, N = 1e7, d = 10, n = 122): same-box A/B via PSBAX_LIB."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev

X = torch.rand(10_000_000, 10, generator=torch.Generator().manual_seed(1)).to(dev)


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


p32 = make(32, 1000, 10, 122, "auto")
p64 = make(64, 1000, 10, 122, "auto")
print(f"{os.environ.get('PSBAX_LIB', 'default'):45s} argmax {timeit(lambda: p32.argmax(X)):.3f} ms   "
      f"top-32 {timeit(lambda: p64.topk(X, 32)):.3f} ms   level set S=64 {timeit(lambda: p64.levelset(X, 0.0)):.3f} ms", flush=True)
