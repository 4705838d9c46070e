"""World-size-2 gloo/CPU tests of the multi-GPU host logic: row sharding, the single all-gather
of the per-shard top-k lists, the deterministic merge, and the level-set summary reduction.
The per-shard reductions and the merge are played by the oracle (test infrastructure); on the
GPUs they are the fused kernels and ``psbax_topk_merge`` (covered by tests/test_gpu_parity.py)."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, random_params

from psbax_b200 import distributed as D


def _merge_numpy(vals, idxs, k):
    lists, S, kk = vals.shape
    ov = torch.empty(S, k)
    oi = torch.empty(S, k, dtype=torch.int64)
    for s in range(S):
        v = vals[:, s].reshape(-1).numpy()
        i = idxs[:, s].reshape(-1).numpy()
        keep = i >= 0
        v, i = v[keep], i[keep]
        order = np.lexsort((i, -v))[:k]
        ov[s, :len(order)] = torch.from_numpy(v[order])
        oi[s, :len(order)] = torch.from_numpy(i[order])
        ov[s, len(order):] = float("-inf")
        oi[s, len(order):] = -1
    return ov, oi


def _worker(rank, world, init_file, N, k, out_dir):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        p = random_params(5, 48, 3, 6, seed=17)
        X = torch.rand(N, 3, generator=torch.Generator().manual_seed(5)).double()
        ref = O.eval_paths(p, X).float()  # what the fused kernels would produce, [N, S]
        lo, hi = D.my_shard(N)
        assert (lo, hi) == (D.shard_bounds(N, world)[rank], D.shard_bounds(N, world)[rank + 1])

        def local_topk(kk):
            vs, is_ = [], []
            for s in range(ref.shape[1]):
                i, v = O.topk_reduce(ref[lo:hi, s], kk)
                vs.append(v)
                is_.append(i + lo)
            return torch.stack(vs), torch.stack(is_)

        val, idx = D.topk_sharded(local_topk, k, hi - lo, merge=_merge_numpy)
        cnt = (ref[lo:hi] > 0.1).sum(0)
        mv = ref[lo:hi].max(0).values
        am = ref[lo:hi].argmax(0) + lo
        total, best_v, best_i = D.reduce_levelset(cnt, mv, am)
        torch.save({"val": val, "idx": idx, "total": total, "best_v": best_v, "best_i": best_i, "ref": ref},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,k", [(1001, 10), (7, 5)])
def test_sharded_topk_and_levelset_world2(N, k):
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, init_file, N, k, tmp), nprocs=world, join=True)
        outs = [torch.load(os.path.join(tmp, f"r{r}.pt")) for r in range(world)]
    ref = outs[0]["ref"]
    for s in range(ref.shape[1]):
        oi, ov = O.topk_reduce(ref[:, s], k)
        for o in outs:  # identical on every rank, equal to the unsharded reduction
            assert o["idx"][s].tolist() == oi.tolist()
            assert torch.equal(o["val"][s], ov)
    for o in outs:
        assert torch.equal(o["total"], (ref > 0.1).sum(0))
        assert torch.equal(o["best_v"], ref.max(0).values)
        assert torch.equal(o["best_i"], ref.argmax(0))


def test_shard_bounds_cover_everything():
    for N in (1, 5, 1000, 16_777_216):
        for w in (1, 2, 3, 8):
            b = D.shard_bounds(N, w)
            assert b[0] == 0 and b[-1] == N and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1


def test_pack_roundtrip():
    v = torch.tensor([[1.5, -float("inf")], [3.25e-7, 2.0]])
    i = torch.tensor([[7, -1], [2**40, 3]])
    vv, ii = D._unpack(D._pack(v, i))
    assert torch.equal(v, vv) and torch.equal(i, ii)
