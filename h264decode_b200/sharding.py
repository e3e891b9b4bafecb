"""Host-side partitioning of independent streams / slices over the GPUs of one box.

The path shards with no exchange step: slices share no CABAC state (cabac.go:439-446, :148-174 re-initialise per
slice) and NAL boundaries are local byte predicates (server.go:28-39), so every rank works on its own streams with its
own context and the only cross-rank traffic is the timing / counter reduction of the benchmark.  This module is that
host logic: a deterministic LPT (longest processing time first) bin packing by bytes, identical on every rank, so no
rank has to be told its share."""
import heapq

import numpy as np


def lpt_assign(sizes, n_ranks):
    """Greedy LPT: items by decreasing size, each to the currently lightest rank (ties: lowest rank, lowest item).
    Returns a list of n_ranks sorted index arrays.  Deterministic, so every rank computes the same partition."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.lexsort((np.arange(len(sizes)), -sizes))  # by -size, then by index
    heap = [(0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    out = [[] for _ in range(n_ranks)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(int(i))
        heapq.heappush(heap, (load + int(sizes[i]), r))
    return [np.array(sorted(x), dtype=np.int64) for x in out]


def my_share(sizes, rank, world):
    return lpt_assign(sizes, world)[rank]


def imbalance(sizes, parts):
    """max rank load / mean rank load (1.0 = perfect)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    loads = np.array([int(sizes[p].sum()) for p in parts], dtype=np.float64)
    return float(loads.max() / loads.mean()) if loads.mean() > 0 else 1.0


def reduce_job(dist, device, seconds, counters):
    """Whole-job numbers from per-rank ones: MAX of the time, SUM of the counters (None dist: single rank)."""
    import torch
    if dist is None:
        return float(seconds), [int(c) for c in counters]
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([int(x) for x in counters], dtype=torch.int64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t[0]), [int(x) for x in c]
