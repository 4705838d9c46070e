"""clock64 pipeline trace of CTA 0 of the projection kernel in its one-tile / several-chunk shape
(S > 64).  Build the trace library first: python -m psbax_b200.build --trace; run with
PSBAX_LIB=psbax_b200/libpsbax_b200_trace.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import _lib
from scripts.proj_probe_lib import make

dev = torch.device("cuda:0")
d = int(os.environ.get("TP_D", "32"))
n = int(os.environ.get("TP_N", "512"))
S = int(os.environ.get("TP_S", "256"))
op = os.environ.get("TP_OP", "topk")
p = make(S, 1000, d, n, "tensor_proj")
X = torch.rand(148 * 128 * 3, d, generator=torch.Generator().manual_seed(1)).to(dev)
run = (lambda: p.topk(X, 10)) if op == "topk" else (lambda: p.values(X))
run()
torch.cuda.synchronize()
buf = torch.zeros(4 * 128 * 8, dtype=torch.int64, device=dev)
assert _lib.load().psbax_debug_trace_buffer(buf.data_ptr()) == 0, "not the trace build"
run()
torch.cuda.synchronize()
_lib.load().psbax_debug_trace_buffer(None)
t = buf.cpu().reshape(4, 128, 8)
t0 = int(t[t > 0].min())
rel = lambda v: int(v) - t0 if v > 0 else -1  # noqa: E731
print("it: P waited, P issued | C enter, ready, C issued | producer(w0): start z_full z_read computed a_empty stored | loader: wait, got; stored by producer warps 5, 10, 15")
for it in list(range(0, 8)) + list(range(28, 40)) + list(range(60, 72)) + list(range(120, 128)):
    m = [rel(t[0, it, k]) for k in (4, 5, 0, 2, 3)]
    q = [rel(v) for v in t[1, it, :6]]
    l = [rel(v) for v in t[3, it, :5]]
    print(f"{it:4d}: " + " ".join(f"{v:8d}" for v in m) + "   | " + " ".join(f"{v:8d}" for v in q) + "   | " + " ".join(f"{v:8d}" for v in l))
print("epilogue per pass: wait start, d_full, released, done")
for pl in range(3):
    print(pl, [rel(v) for v in t[2, pl, :4]])
