"""Timing table over the BASELINE.json config shapes (synthetic data): fused op per engine.
Profiling aid; the contract benchmark is bench.py."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import SamplePathBatch

dev = torch.device("cuda:0")


def make(S, L, d, n, engine, G=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    omega = 6 * torch.randn(G, L, d, generator=g)
    bias = 6.28 * torch.rand(G, L, generator=g)
    W = torch.randn(S, L, generator=g) * (2.0 / max(L, 1)) ** 0.5
    kw = {}
    if n:
        kw = dict(Xtrain=torch.rand(n, d, generator=g) / 0.3, inv_ls=torch.full((d,), 1 / 0.3),
                  V=torch.randn(S, n, generator=g))
    return SamplePathBatch(omega, bias, W, "matern52", device=dev, engine=engine, **kw)


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


CONFIGS = [
    # name, N, d, S, L, n, op, arg
    ("C1 demo top-k", 100, 3, 1, 1000, 8, "topk", 4),
    ("C1 demo top-k (30 paths)", 100, 3, 30, 1000, 38, "topk", 4),
    ("C2 level-set 1e6", 1_000_000, 2, 64, 1000, 64, "levelset", 0.0),
    ("C2 level-set 1.68e7", 16_777_216, 2, 64, 1000, 64, "levelset", 0.0),
    ("C3 GB1 top-k d=32 emb", 149_361, 32, 256, 1000, 512, "topk", 10),
    ("C3 GB1 top-k d=80 onehot", 149_361, 80, 256, 1000, 512, "topk", 10),
    ("C4 DiscoBAX values d=5", 17_176, 5, 1024, 1000, 128, "values", None),
    ("C4 DiscoBAX subset d=5", 17_176, 5, 1024, 1000, 128, "subset", 10),
    ("C5 Ackley argmax", 10_000_000, 10, 32, 1000, 122, "argmax", None),
    ("C5 Ackley top-32", 10_000_000, 10, 64, 1000, 122, "topk", 32),
    ("ref-form level-set (G=S)", 1_000_000, 2, 64, 1000, 0, "levelset_ref", 0.0),
]
engines = sys.argv[1].split(",") if len(sys.argv) > 1 else ["auto", "tensor", "simt"]
print(f"{'config':32s} {'engine':12s} {'ms':>10s} {'Gevals/s':>10s}")
for (name, N, d, S, L, n, op, arg) in CONFIGS:
    g = torch.Generator().manual_seed(1)
    X = torch.rand(N, d, generator=g).to(dev)
    etas = torch.randn(100, N, generator=g).to(dev) if op == "subset" else None
    for eng in engines:
        if op == "levelset_ref" and eng != "auto":
            continue
        try:
            p = make(S, L, d, n, eng if op != "levelset_ref" else "simt", G=S if op == "levelset_ref" else 1)
        except Exception as ex:  # engine does not support the shape
            print(f"{name:32s} {eng:12s} {'n/a':>10s}   ({str(ex)[:60]})")
            continue
        fn = {"topk": lambda: p.topk(X, arg), "levelset": lambda: p.levelset(X, arg),
              "levelset_ref": lambda: p.levelset(X, arg), "values": lambda: p.values(X),
              "argmax": lambda: p.argmax(X), "subset": lambda: p.subset_select(X, etas, arg)}[op]
        ms = timeit(fn, reps=3 if N * S > 1e8 else 10)
        print(f"{name:32s} {eng:12s} {ms:10.3f} {N * S / ms / 1e6:10.2f}")
