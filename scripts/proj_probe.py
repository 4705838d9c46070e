"""Projection-variant probe: tensor_proj against the FFMA engine (values / level set / top-k) and timings."""
import sys, os
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


for (S, L, d, n, N) in [(64, 96, 10, 0, 1000), (64, 1000, 10, 122, 5000), (256, 1000, 31, 512, 3001), (256, 1000, 32, 512, 3001), (64, 300, 8, 20, 3001),
                        (70, 200, 5, 9, 100_000)]:
    g = torch.Generator().manual_seed(7)
    X = torch.rand(N, d, generator=g).to(dev)
    a, b = make(S, L, d, n, "tensor_proj"), make(S, L, d, n, "simt")
    va, vb = a.values(X), b.values(X)
    err = (va - vb).abs().max().item()
    scale = vb.abs().max().item()
    ta, ia = a.topk(X, 8)
    tb, ib = b.topk(X, 8)
    ra, rb = a.levelset(X, 0.1), b.levelset(X, 0.1)
    print(f"S={S} L={L} d={d} n={n} N={N}: values max err {err:.3e} (scale {scale:.2f}); topk val err "
          f"{(ta - tb).abs().max().item():.3e} idx equal {(ia == ib).float().mean().item():.3f}; "
          f"count diff {(ra.count - rb.count).abs().max().item()}", flush=True)

for (name, N, d, S, L, n, op, arg) in [("C5 argmax", 10_000_000, 10, 32, 1000, 122, "argmax", None),
                                       ("C5 top-32", 10_000_000, 10, 64, 1000, 122, "topk", 32),
                                       ("C3 d=32", 149_361, 32, 256, 1000, 512, "topk", 10),
                                       ("d=16 level-set", 4_000_000, 16, 64, 1000, 64, "levelset", 0.0),
                                       ("d=8 level-set", 4_000_000, 8, 64, 1000, 64, "levelset", 0.0),
                                       ("d=5 level-set", 4_000_000, 5, 64, 1000, 64, "levelset", 0.0)]:
    X = torch.rand(N, d, generator=torch.Generator().manual_seed(1)).to(dev)
    for eng in ["tensor_proj", "tensor_bf16"]:
        p = make(S, L, d, n, eng)
        fn = {"topk": lambda: p.topk(X, arg), "levelset": lambda: p.levelset(X, arg), "argmax": lambda: p.argmax(X)}[op]
        ms = timeit(fn, reps=3)
        print(f"{name:18s} {eng:12s} {ms:9.3f} ms {N * S / ms / 1e6:9.2f} Gevals/s", flush=True)
