#!/usr/bin/env python3
"""Model of cabac_decode_kernel's run time on BASELINE configs[3] as a function of how the 2 500 bundles (32 slices of
similar length each, longest first) are placed on the 592 warp schedulers.  CPU only (numpy).

A scheduler that runs k warps of the fast loop issues for each of them one op per  max(c1, cs * k)  cycles: c1 = a lone
warp's cycles per op (latency bound), cs = the issue port's cycles per warp-op (tools/cabac_exp2.py measures both:
equal-length slices at 0.5 and at 4 warps per scheduler).  A scheduler's time follows from its bundles' lengths; the
launch ends with the slowest scheduler.  With c1 = 115, cs = 52 (kLoop 2) the model gives 79.6 ms for the round-robin
placement (80.7 ms measured), 62 ms for "longest first onto the least loaded scheduler, near-critical bundles weighted
1.5" (bundle_assign_kernel), 59 ms for a greedy on the simulated time itself."""
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness as hz  # noqa: E402


def t_sched(L, c1, cs):
    L = sorted(L)
    t, done = 0.0, 0.0
    for i, l in enumerate(L):
        t += (l - done) * max(c1, cs * (len(L) - i))
        done = l
    return t


def main():
    c1, cs = float(os.environ.get("C1", 115)), float(os.environ.get("CS", 52))
    n, S, G = 80000, 592, 148
    nb = hz.slice_bins_normal(n, 455000, 4, id_base=0)
    bund = np.sort(nb)[::-1].reshape(-1, 32).max(axis=1).astype(float)
    B = len(bund)

    def report(assign, name):
        ts = [t_sched(a, c1, cs) for a in assign]
        print("%-44s makespan %6.1f ms   mean scheduler %5.1f ms   most warps %d" % (
            name, max(ts) / 1.965e6, np.mean(ts) / 1.965e6, max(len(a) for a in assign)))

    W = 17
    cur = [[] for _ in range(S)]
    for g in range(B):
        rank, b = divmod(g, G)
        cur[b * 4 + (W - 1 - rank) % 4].append(bund[g])
    report(cur, "round robin over the SMs (map 1)")
    h = [(0.0, s) for s in range(S)]
    heapq.heapify(h)
    a = [[] for _ in range(S)]
    thr = bund.sum() / S
    for g in range(B):
        tot, s = heapq.heappop(h)
        a[s].append(bund[g])
        heapq.heappush(h, (tot + bund[g] * (1.5 if bund[g] * 2.2 > 0.8 * thr else 1.0), s))
    report(a, "longest first, weighted (map 3)")
    gr = [[] for _ in range(S)]
    for g in range(B):
        s = int(np.argmin([t_sched(gr[k] + [bund[g]], c1, cs) for k in range(S)]))
        gr[s].append(bund[g])
    report(gr, "greedy on the simulated time")
    print("lower bounds: longest slice alone %.1f ms, all ops at the issue rate %.1f ms" % (
        bund[0] * c1 / 1.965e6, bund.sum() * cs / S / 1.965e6))


if __name__ == "__main__":
    main()
