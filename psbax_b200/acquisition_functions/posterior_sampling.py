"""PS-BAX acquisition (``/root/reference/src/acquisition_functions/posterior_sampling.py:46-83``).

The reference draws ``batch_size`` sample paths one after another and runs the algorithm on
each in Python chunks of 100 candidates.  Here all paths are drawn together and the
algorithm runs on the whole batch in one fused GPU launch; pooling and the max-entropy pick
are unchanged.
"""
import math

import torch

from ..utils import get_function_sample_batch
from .entropy import EntropyAcquisitionFunction, _fusable, optimize_acqf_discrete


def _fused_levelset_pick(model, algorithm, paths, q):
    """Max-entropy pick for a level-set problem without leaving the GPU (row f1 of SURVEY.md section 8).

    The reference pools the coordinate rows of every path's super-level set (S x ~0.45 N rows with
    duplicates) and scores them in chunks of 100.  A duplicate of a picked row can never be picked
    again (singular joint covariance), so the greedy pick over the de-duplicated pool is the same
    pick.  Here: one fused level-set launch, an OR over the S bit masks, and per greedy step one
    fused posterior-variance launch over ALL candidates (conditioned on the rows picked so far)
    followed by a masked arg max.  Only the q picked rows are gathered from ``x_set``."""
    from ..sample_paths import posterior_variance

    x_set = torch.as_tensor(algorithm.params.x_set)
    X = algorithm.device_candidates(paths.device)
    pool = algorithm.pooled_mask(paths)
    picked = []
    for _ in range(min(q, int(pool.sum()))):
        pend = x_set[picked] if picked else None
        var = posterior_variance(model, X, pending=pend)
        j = int(torch.where(pool, var, torch.full_like(var, -float("inf"))).argmax())
        picked.append(j)
        pool[j] = False
    return x_set[picked]


def gen_posterior_sampling_batch(model, algorithm, batch_size, **kwargs):
    if algorithm.params.name == "SubsetSelect":
        # at least batch_size labels: ceil(batch_size / k) paths (posterior_sampling.py:51-54)
        k = algorithm.params.k
        num_paths = max(1, math.ceil(batch_size / k))
        paths = get_function_sample_batch(model, num_paths, **kwargs)
        batch = []
        for labels in algorithm.execute_batch(paths):
            batch.extend(labels)
        x_batch = algorithm.index_to_x(batch)
        acq_func = EntropyAcquisitionFunction(model=model)
        x_next, _ = optimize_acqf_discrete(acq_function=acq_func, q=batch_size, choices=x_batch,
                                           max_batch_size=100)
        return algorithm.obj_func.get_idx_from_x(x_next)

    paths = get_function_sample_batch(model, batch_size, **kwargs)
    acq_func = EntropyAcquisitionFunction(model=model)
    if algorithm.params.name == "SimpleLevelSet" and hasattr(algorithm, "pooled_mask") and _fusable(acq_func):
        return _fused_levelset_pick(model, algorithm, paths, batch_size)
    rows = []
    for x_output in algorithm.execute_batch(paths):
        x_output = torch.as_tensor(x_output)
        if x_output.dim() == 1:
            x_output = x_output.reshape(1, -1)
        rows.append(x_output)
    x_batch = torch.cat(rows)  # pooled outputs of all paths, duplicates kept
    x_next, _ = optimize_acqf_discrete(acq_function=acq_func, q=batch_size, choices=x_batch,
                                       max_batch_size=100)
    return x_next
