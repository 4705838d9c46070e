"""Launch ONE of the secondary kernels a few times (target of an `ncu -k regex:<kernel> -c 1` capture):
    one_kernel.py persample | simt | generic     Profiling aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.config_bench_lib import make, dev

mode = sys.argv[1]
g = torch.Generator().manual_seed(1)
if mode == "persample":    # reference-parity form: one RFF basis per sample (k_persample)
    p, X, fn = make(64, 1000, 2, 0, "simt", G=64), torch.rand(1_000_000, 2, generator=g).to(dev), "levelset"
elif mode == "simt":       # FFMA engine on the C2 shape (k_shared_simt)
    p, X, fn = make(64, 1000, 2, 64, "simt"), torch.rand(1_000_000, 2, generator=g).to(dev), "levelset"
else:                      # generic-d tensor kernel (k_shared_tensor<0, ...>): C3 one-hot shape, d = 80
    p, X, fn = make(256, 1000, 80, 512, "auto"), torch.rand(149_361, 80, generator=g).to(dev), "topk"
for _ in range(4):
    r = p.levelset(X, 0.0) if fn == "levelset" else p.topk(X, 10)
torch.cuda.synchronize()
print(mode, "ok")
