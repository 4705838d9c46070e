"""Candidate-set sharding over the GPUs of one node.

Candidates are independent given the (replicated, a few MB) path parameters, so rank r owns the
contiguous row block [bounds[r], bounds[r+1]) of X*, resident on its GPU.

* level set / arg-max / ES scoring: no data-path communication.  The per-shard masks stay
  sharded; ``reduce_levelset`` optionally combines the S counts and the S (max, arg max) pairs
  (the empty-set fallback of ``src/algorithms/levelset.py:35-37``) with one small all-gather.
* top-k: each rank runs the fused kernel on its shard (global indices via ``row_offset``), then
  ONE all-gather of the [S, k] (value, index) lists and an identical deterministic k-of-(G k)
  merge (``psbax_topk_merge``, ties to the smaller global index) on every rank.

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPUs; the same code runs on
gloo/CPU tensors in the unit tests, where the local reduction and the merge are injected).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(N: int, world: int) -> List[int]:
    """Row boundaries of the `world` contiguous shards of N candidates (sizes differ by <= 1)."""
    return [(r * N) // world for r in range(world + 1)]


def my_shard(N: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    b = shard_bounds(N, world)
    return b[rank], b[rank + 1]


def _pack(val: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(float32 [..], int64 [..]) -> one int64 [.., 2] buffer, so that a single collective moves both."""
    bits = val.contiguous().view(torch.int32).to(torch.int64)
    return torch.stack([idx.to(torch.int64), bits], dim=-1).contiguous()


def _unpack(buf: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    idx = buf[..., 0].contiguous()
    val = buf[..., 1].to(torch.int32).contiguous().view(torch.float32)
    return val, idx


def all_gather_lists(val: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One all-gather of the per-rank [S, k] lists -> ([world, S, k] values, [world, S, k] indices)."""
    world = dist.get_world_size(group)
    mine = _pack(val, idx)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return _unpack(torch.stack(out))


def merge_topk_device(vals: torch.Tensor, idxs: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[lists, S, k] sorted lists on one GPU -> merged [S, k] (CUDA kernel ``psbax_topk_merge``)."""
    from . import _lib

    lib = _lib.load()
    lists, S, kk = vals.shape
    assert kk == k and vals.is_cuda
    vals, idxs = vals.contiguous(), idxs.contiguous()
    ov = torch.empty(S, k, dtype=torch.float32, device=vals.device)
    oi = torch.empty(S, k, dtype=torch.int64, device=vals.device)
    with torch.cuda.device(vals.device):
        _lib.check(lib.psbax_topk_merge(vals.data_ptr(), idxs.data_ptr(), lists, S, k, ov.data_ptr(),
                                        oi.data_ptr(), torch.cuda.current_stream(vals.device).cuda_stream),
                   "psbax_topk_merge")
    return ov, oi


def topk_sharded(local_topk: Callable[[int], Tuple[torch.Tensor, torch.Tensor]], k: int, shard_rows: int,
                 merge: Callable = merge_topk_device, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global per-sample top-k from per-shard top-k.

    ``local_topk(kk)`` returns this rank's ([S, kk] values, [S, kk] GLOBAL indices) for
    kk = min(k, shard_rows) (e.g. ``lambda kk: paths.topk(X_local, kk, row_offset=lo)``).  Lists
    shorter than k are padded with (-inf, -1), which the merge ignores."""
    kk = min(k, shard_rows)
    if kk > 0:
        val, idx = local_topk(kk)
    else:
        raise ValueError("empty shard: use fewer ranks than candidates")
    S = val.shape[0]
    if kk < k:
        pv = torch.full((S, k), float("-inf"), dtype=val.dtype, device=val.device)
        pi = torch.full((S, k), -1, dtype=torch.int64, device=val.device)
        pv[:, :kk], pi[:, :kk] = val, idx
        val, idx = pv, pi
    gv, gi = all_gather_lists(val, idx, group)
    return merge(gv, gi, k)


def reduce_levelset(count: torch.Tensor, maxval: torch.Tensor, argmax: torch.Tensor, group=None):
    """Combine per-shard level-set summaries: total counts and the global (max, arg max) per sample
    (ties to the smaller global index).  One all-gather of [S, 3] int64."""
    world = dist.get_world_size(group)
    mine = torch.stack([count.to(torch.int64), argmax.to(torch.int64),
                        maxval.contiguous().view(torch.int32).to(torch.int64)], dim=-1).contiguous()
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    allb = torch.stack(out)  # [world, S, 3]
    total = allb[..., 0].sum(0)
    vals = allb[..., 2].to(torch.int32).contiguous().view(torch.float32)  # [world, S]
    idxs = allb[..., 1]
    best_v = vals.max(0).values
    cand = torch.where(vals == best_v[None], idxs, torch.full_like(idxs, torch.iinfo(torch.int64).max))
    cand = torch.where(idxs < 0, torch.full_like(idxs, torch.iinfo(torch.int64).max), cand)
    best_i = cand.min(0).values
    return total, best_v, best_i
