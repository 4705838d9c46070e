"""Multi-GPU check (run under torchrun, one rank per GPU): sharded fused top-k + one NCCL
all-gather + device merge == the unsharded launch, bit for bit; level-set summaries reduce to the
unsharded ones.  Times the all-gather + merge step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from psbax_b200 import SamplePathBatch
from psbax_b200 import distributed as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(0)
S, L, d, n, N, k = 256, 1000, 5, 128, 149_361, 10  # GB1-like shape (config 3), 5-d features
omega = 6 * torch.randn(1, L, d, generator=g)
bias = 6.28 * torch.rand(1, L, generator=g)
W = torch.randn(S, L, generator=g) * (2.0 / L) ** 0.5
Xt = torch.rand(n, d, generator=g)
V = torch.randn(S, n, generator=g)
paths = SamplePathBatch(omega, bias, W, "matern52", Xt / 0.3, torch.full((d,), 1 / 0.3), V, device=dev)
X = torch.rand(N, d, generator=g).to(dev)
lo, hi = D.my_shard(N)
wv, wi = paths.topk(X, k)
for _ in range(3):
    sv, si = D.topk_sharded(lambda kk: paths.topk(X[lo:hi], kk, row_offset=lo), k, hi - lo)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    sv, si = D.topk_sharded(lambda kk: paths.topk(X[lo:hi], kk, row_offset=lo), k, hi - lo)
e1.record()
torch.cuda.synchronize()
ok = torch.equal(si, wi) and torch.equal(sv, wv)
whole = paths.levelset(X, 0.0)
part = paths.levelset(X[lo:hi], 0.0, row_offset=lo)
tot, bv, bi = D.reduce_levelset(part.count, part.maxval, part.argmax)
ok = ok and torch.equal(tot, whole.count) and torch.equal(bv, whole.maxval) and torch.equal(bi, whole.argmax)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded==whole: {bool(flag.item())}; sharded top-k step (kernel + all-gather + merge): "
          f"{e0.elapsed_time(e1) / 20:.3f} ms for N={N}, S={S}, k={k}")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
