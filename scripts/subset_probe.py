"""SubsetSelect timing on the C4 shape (N = 17176, d = 5, S = 1024 paths, 100 noise draws, k = 10)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scripts.proj_probe_lib import make, dev

N, d, S, B, k = 17_176, 5, int(os.environ.get("SS_S", "1024")), 100, 10
p = make(S, 1000, d, 128, "auto")
g = torch.Generator().manual_seed(1)
X = torch.rand(N, d, generator=g).to(dev)
etas = torch.randn(B, N, generator=g).to(dev)
for _ in range(2):
    p.subset_select(X, etas, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    p.subset_select(X, etas, k)
e1.record()
torch.cuda.synchronize()
print(f"subset_select S={S}: {e0.elapsed_time(e1) / 5:.3f} ms (values + greedy selection)")
