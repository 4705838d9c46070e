"""Small shapes through every kernel family, for compute-sanitizer (memcheck / racecheck).
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from helpers import candidates, random_params, to_batch

for (S, L, d, n, G, N, engines) in [
    (64, 96, 2, 16, 1, 700, ["simt", "tensor", "tensor_bf16"]),
    (20, 64, 10, 9, 1, 300, ["simt", "tensor", "tensor_bf16"]),
    (16, 64, 40, 8, 1, 300, ["simt", "tensor_bf16"]),
    (5, 64, 3, 4, 5, 200, ["simt"]),
    (1, 0, 3, 12, 1, 100, ["simt"]),
]:
    p = random_params(S, L, d, n, G=G, seed=1)
    X = candidates(N, d).cuda()
    for eng in engines:
        b = to_batch(p, engine=eng)
        b.values(X)
        b.levelset(X, 0.0)
        b.topk(X, min(10, N))
        b.argmax(X)
        b.grad(X[:8], torch.arange(8) % S)
        if eng == "simt":
            b.subset_select(X, torch.randn(6, N), 3)
        torch.cuda.synchronize()
        print("ok", eng, S, L, d, n, G, N)
print("sanitize run finished")
