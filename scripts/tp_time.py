"""Timing of the tensor-projection kernel on one shape (experiment builds via PSBAX_LIB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev

d = int(os.environ.get("TP_D", "8"))
n = int(os.environ.get("TP_N", "0"))
N = int(os.environ.get("TP_ROWS", "4000000"))
eng = os.environ.get("TP_ENGINE", "tensor_proj")
p = make(64, 1000, d, n, eng)
X = torch.rand(N, d, generator=torch.Generator().manual_seed(1)).to(dev)
for _ in range(2):
    p.levelset(X, 0.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    p.levelset(X, 0.0)
e1.record()
torch.cuda.synchronize()
print(f"{os.environ.get('PSBAX_LIB', 'default')} d={d} n={n} {eng}: {e0.elapsed_time(e1) / 3:.3f} ms", flush=True)
