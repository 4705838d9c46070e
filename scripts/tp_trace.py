"""clock64 pipeline trace of CTA 0 of the tensor-projection kernel (build with `python -m psbax_b200.build --trace`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import _lib
from scripts.proj_probe_lib import make

dev = torch.device("cuda:0")
d = int(os.environ.get("TP_D", "10"))
n = int(os.environ.get("TP_N", "0"))
p = make(64, 1000, d, n, "tensor_proj")
X = torch.rand(148 * 256 * 4, d, generator=torch.Generator().manual_seed(1)).to(dev)
p.levelset(X, 0.0)
torch.cuda.synchronize()
buf = torch.zeros(4 * 128 * 8, dtype=torch.int64, device=dev)
_lib.load().psbax_debug_trace_buffer(buf.data_ptr())
p.levelset(X, 0.0)
torch.cuda.synchronize()
_lib.load().psbax_debug_trace_buffer(None)
t = buf.cpu().reshape(4, 128, 8)
t0 = int(t[t > 0].min())
rel = lambda v: int(v) - t0 if v > 0 else -1  # noqa: E731
print("it: P waited, P issued | C enter, a_full, C issued | producer(w0): start z_full z_read computed a_empty stored")
for it in range(36, 48):
    m = [rel(t[0, it, k]) for k in (4, 5, 0, 2, 3)]
    q = [rel(v) for v in t[1, it, :6]]
    print(f"{it:4d}: " + " ".join(f"{v:8d}" for v in m) + "   | " + " ".join(f"{v:8d}" for v in q))
