"""Sharding of a SeqPairList over the GPUs of one box, and the one exchange step of the path.

Every sequence pair is an independent unit: Forward, Viterbi and alignment
need no communication.  The E-step of Baum-Welch (MachineCounts over the list, src/counts.cpp:37-43)
sums per-pair counts, so after each rank has processed its shard the count vector and the total
log-likelihood are combined with ONE all-reduce of nTransitions + 1 doubles -- NCCL over NVLink on
the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def lpt_assign(costs, world: int):
    """Longest-processing-time-first assignment of pairs to ranks; returns a list of index arrays.

    costs[k] ~ (Li+1)*(Lo+1).  Ties and order are deterministic so every rank computes the same plan.
    """
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    bins = [[] for _ in range(world)]
    for k in order:
        r = int(np.argmin(load))
        bins[r].append(int(k))
        load[r] += costs[k]
    return [np.array(sorted(b), dtype=np.int64) for b in bins]


def allreduce_counts(counts: np.ndarray, loglike: float, device=None):
    """Sum the per-rank count vectors and log-likelihoods over the process group (no-op if not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts, loglike
    t = torch.from_numpy(np.concatenate([np.asarray(counts, dtype=np.float64), [float(loglike)]]))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t = t.cpu().numpy()
    return t[:-1].copy(), float(t[-1])


def gather_by_pair(values: np.ndarray, mine: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Place this rank's per-pair results at their pair indices and combine over ranks."""
    import torch
    import torch.distributed as dist
    out = np.zeros(n_total, dtype=np.float64)
    out[mine] = values
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out
    t = torch.from_numpy(out)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
