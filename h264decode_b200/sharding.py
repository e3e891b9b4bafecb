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


START_CODE = b"\x00\x00\x00\x01"


def cut_byte_ranges(stream, n_shards):
    """One long Annex-B stream cut into n_shards byte ranges that can be scanned independently ("a
    single long stream may instead be cut into byte ranges").  Every cut is moved forward from its nominal position
    k * len / n_shards to the next start code 00 00 00 01 (server.go:19, :28-39: the only boundary the reference
    knows), and a range keeps the 4 bytes of the start code that opens the next one, because the reference's NAL is
    "payload plus the following start code" (server.go:64-111) and K start codes yield K - 1 NAL units.  The pattern
    cannot overlap itself, so the NAL units of range r are exactly those of the whole stream whose start code lies in
    [cut_r, cut_{r+1}); positions in a range's results are relative to its begin.  Ranges can be empty (fewer NAL
    units than shards).  Only the bytes between a nominal cut and the next start code are looked at: O(n_shards x NAL
    size) host work, no pass over the stream.

    -> list of (begin, end) with stream[begin:end] the input of shard r; deterministic, identical on every rank."""
    buf = stream if isinstance(stream, (bytes, bytearray, memoryview)) else memoryview(np.ascontiguousarray(stream, np.uint8))
    buf = bytes(buf) if isinstance(buf, bytearray) else buf
    n = len(buf)
    find = buf.find if isinstance(buf, bytes) else None
    cuts = [0]
    for k in range(1, n_shards):
        nominal = max(cuts[-1], (k * n) // n_shards)
        lo = max(nominal - 3, cuts[-1])  # a start code that straddles the nominal cut belongs to this shard
        if find is not None:
            p = find(START_CODE, lo)
        else:  # memoryview over a numpy array: search window by window without copying the stream
            p, w = -1, lo
            while w < n and p < 0:
                q = bytes(buf[w:w + (1 << 20) + 3]).find(START_CODE)
                p = w + q if q >= 0 else -1
                w += 1 << 20
        cuts.append(n if p < 0 else p)
    cuts.append(n)
    return [(cuts[r], min(n, cuts[r + 1] + 4) if r + 1 < n_shards else n) for r in range(n_shards)]


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
