"""clock64 pipeline trace of CTA 0 of the projection kernel, any shape: per-block phase durations of one
steady-state pass.  Build the trace library first (python -m psbax_b200.build --trace) and run with
PSBAX_LIB=psbax_b200/libpsbax_b200_trace.so.   TP_D, TP_N, TP_S, TP_L, TP_OP = levelset | topk | values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200 import _lib
from scripts.proj_probe_lib import make

dev = torch.device("cuda:0")
d = int(os.environ.get("TP_D", "10"))
n = int(os.environ.get("TP_N", "122"))
S = int(os.environ.get("TP_S", "32"))
L = int(os.environ.get("TP_L", "1000"))
op = os.environ.get("TP_OP", "levelset")
rows_per_pass = 256 if S <= 64 else 128
p = make(S, L, d, n, "tensor_proj")
X = torch.rand(148 * rows_per_pass * 4, d, generator=torch.Generator().manual_seed(1)).to(dev)
run = {"levelset": lambda: p.levelset(X, 0.0), "topk": lambda: p.topk(X, 10), "values": lambda: p.values(X)}[op]
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
nkb_r = (L + 31) // 32
nkb = nkb_r + (n + 15) // 16
print(f"d={d} n={n} S={S} L={L} op={op}: {nkb_r} RFF blocks + {nkb - nkb_r} exact-kernel blocks per pass")
print("blk: C.issued(abs)  dC | P: proj wait->issued | C: enter->ready->issued | producer w0: start zwait zread compute awaitA store")
prev = None
for it in range(nkb, min(2 * nkb + 2, 128)):
    c_enter, c_ready, c_done = (rel(t[0, it, k]) for k in (0, 2, 3))
    p_w, p_i = rel(t[0, it, 4]), rel(t[0, it, 5])
    q = [rel(v) for v in t[1, it, :6]]
    dq = [q[i + 1] - q[i] for i in range(5)]
    dc = c_done - prev if prev is not None else 0
    prev = c_done
    kind = "K" if (it % nkb) >= nkb_r else "r"
    print(f"{it:4d}{kind} {c_done:9d} {dc:6d} | {p_i - p_w:6d} | {c_ready - c_enter:6d} {c_done - c_ready:6d} | "
          + " ".join(f"{v:6d}" for v in dq))
print("epilogue per pass: wait start, d_full, released, done")
for pl in range(4):
    print(pl, [rel(v) for v in t[2, pl, :4]])
