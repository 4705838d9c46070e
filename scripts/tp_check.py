"""Quick parity probe of the projection engine against the FFMA engine on one shape (experiment aid).
   TP_D, TP_N, TP_S, TP_L, TP_ROWS"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev

d = int(os.environ.get("TP_D", "5"))
n = int(os.environ.get("TP_N", "128"))
S = int(os.environ.get("TP_S", "256"))
L = int(os.environ.get("TP_L", "200"))
N = int(os.environ.get("TP_ROWS", "1500"))
X = torch.rand(N, d, generator=torch.Generator().manual_seed(1)).to(dev)
ref = make(S, L, d, n, "simt").values(X)
torch.cuda.synchronize()
out = make(S, L, d, n, "tensor_proj").values(X)
torch.cuda.synchronize()
err = (out - ref).abs()
print(f"d={d} n={n} S={S} L={L} N={N}: max err {err.max().item():.3e} (scale {ref.abs().max().item():.2f}); "
      f"worst row {int(err.max(dim=1).values.argmax())}, worst sample {int(err.max(dim=0).values.argmax())}", flush=True)
