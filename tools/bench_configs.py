#!/usr/bin/env python3
"""Measurements at the other BASELINE.json configs (the bench line is configs[3]): one GPU, CUDA events on the library's
stream, median of repeated runs, every result checked against size-independent properties.  Writes one text report.

    python tools/bench_configs.py > profiles/r1_configs_measured.txt

  configs[0]  1 MB Annex-B stream with EPB-bearing payloads (and 512 copies of it: the EPB-dense worst case of the scan)
  configs[1]  64 slices x 100 k bins: the latency-bound end of the CABAC engine (both bypass forms)
  configs[2]  context init, 1 Mi slices x 1024 contexts
  configs[4]  multi-camera batch: this GPU's LPT share (1/8) of 4096 streams x 16 slices, slice sizes 1 KB .. 1 MB skewed:
              (a) as one batch, (b) in 8 jobs of 64 streams through the asynchronous job API
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import harness as hz  # noqa: E402
from h264decode_b200 import capi, sharding  # noqa: E402

dev = "cuda:0"
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
HBM = 6544.7
try:
    import json
    HBM = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


def scan_case(name, host_stream, reps_of=1):
    s = torch.from_numpy(host_stream).to(dev)
    if reps_of > 1:
        s = s.repeat(reps_of)
    n = s.numel()
    d_stream = torch.cat([s, torch.zeros(64, dtype=torch.uint8, device=dev)])
    cap = n // 64 + 1024
    d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
    d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
    t = timed(lambda: ctx.annexb_scan_dev(d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap,
                                          d_sum.data_ptr(), 0))
    summ = np.frombuffer(d_sum.cpu().numpy().tobytes()[:48], dtype=np.uint64, count=5)
    alg = n + int(summ[2]) + 20 * int(summ[1])
    print("%-52s %9.1f KiB  %6d NAL units  %7d EPBs  %8.3f ms  %7.1f GB/s algorithmic = %.3f of the HBM peak" % (
        name, n / 1024, summ[1], summ[4], t, alg / t / 1e6, alg / t / 1e6 / HBM), flush=True)


print("== configs[0]: Annex-B split + strip (algorithmic bytes = in + RBSP out + 20 B per NAL unit)")
one = hz.build_stream_c1(1 << 20)
scan_case("1 MB stream (launch-bound: 8 stream operations)", one)
scan_case("512 copies of it (EPB-dense: every chunk is dirty)", one, 512)
rng = np.random.default_rng(1)
clean = rng.integers(4, 256, 1 << 29, dtype=np.uint8)
clean[np.arange(0, len(clean) - 8, 50000)[:, None] + np.arange(5)] = [0, 0, 0, 1, 0x41]
scan_case("512 MiB, start codes only (no zero pairs elsewhere)", clean)
del clean

print("== configs[1]: CABAC engine, 64 slices x 100 000 bins (one slice per warp: latency-bound)")
n = 64
ops = hz.gen_schedule(2, 100000, 64)
qp, idc = hz.slice_params(n)
g = hz.gpu_build_stream_cabac(torch, dev, n, 100000, config=2, n_active=64, n_ctx=64, slices_per_frame=8, frames_per_params=4,
                              n_bins=np.full(n, 100000))
torch.cuda.synchronize()
cap = g["n_nals"] + 16
d_rbsp = torch.empty(g["n"] + 64, dtype=torch.uint8, device=dev)
d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
d_off = torch.empty(n, dtype=torch.int64, device=dev)
d_len = torch.empty(n, dtype=torch.int32, device=dev)
d_snal = torch.empty(n, dtype=torch.int32, device=dev)
d_ns = torch.zeros(4, dtype=torch.int32, device=dev)
ctx.annexb_scan_dev(g["stream"].data_ptr(), g["n"], d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap, d_sum.data_ptr(), 0)
ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), cap, 0, n, d_off.data_ptr(), d_len.data_ptr(), d_snal.data_ptr(),
                     d_ns.data_ptr())
d_ops = torch.from_numpy(g["ops"].view(np.int16)).to(dev)
d_nops = torch.from_numpy(g["n_ops"].view(np.int32)).to(dev)
p = capi.Context.slice_qp(g["qp"], g["idc"])
d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
words = 100000 // 32 + 2
d_bins = torch.zeros(n * words, dtype=torch.int32, device=dev)
d_fin = torch.empty(n * 32, dtype=torch.uint8, device=dev)
for label, flags in (("SPEC_OR bypass (window engine)", capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE),
                     ("REF_SHIFT bypass (literal int64 engine)", capi.CABAC_FINAL_TERMINATE)):
    t = timed(lambda: ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=g["n"] + 16, off=d_off.data_ptr(),
                                           len=d_len.data_ptr(), n_slices=n, n_ctx=64, ops=d_ops.data_ptr(),
                                           n_ops_max=len(g["ops"]), n_ops=d_nops.data_ptr(), qp=d_qp.data_ptr(),
                                           init_states=None, bins=d_bins.data_ptr(), bins_off=None, bins_stride_words=words,
                                           final=d_fin.data_ptr(), final_states=None, flags=flags))
    fin = np.frombuffer(d_fin.cpu().numpy().tobytes(), dtype=capi.FINAL_DTYPE)
    ok = bool(np.all(fin["n_bins"] == 100001))
    if flags & capi.BYPASS_SPEC_OR:   # decoded with the form the encoder used: the final terminate bin is the 1 it wrote
        last = d_bins.view(n, words)[:, 100000 // 32].cpu().numpy().view(np.uint32)
        ok = ok and bool(np.all((last >> (100000 % 32)) & 1 == 1)) and not (fin["flags"] & capi.F_OVERRUN).any()
    print("%-44s %8.3f ms  %7.2f Gbins/s  %6.1f ns per bin and slice (%.0f cycles at 1.965 GHz)  verified %s" % (
        label, t, n * 100001 / t / 1e6, t * 1e6 / 100001, t * 1e6 / 100001 * 1.965, ok), flush=True)

print("== configs[2]: context init, 1 Mi slices x 1024 contexts (algorithmic bytes = n_slices x (1024 + 8))")
ns = 1 << 20
qp = (np.arange(ns) % 52).astype(np.int32)
idc = ((np.arange(ns) // 52) % 4 - 1).astype(np.int32)
p = capi.Context.slice_qp(qp, idc)
d_p = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
d_st = torch.empty((ns, 1024), dtype=torch.uint8, device=dev)
for label, fl in (("REF tables", 0), ("SPEC tables", 1)):
    t = timed(lambda: ctx.ctx_init_dev(d_p.data_ptr(), ns, 1024, d_st.data_ptr(), fl))
    alg = ns * (1024 + 8)
    print("%-44s %8.3f ms  %7.1f GB/s = %.3f of the HBM peak (write-only stream)" % (label, t, alg / t / 1e6, alg / t / 1e6 / HBM),
          flush=True)
del d_st

if os.environ.get("BENCH_CONFIGS_SKIP_C4"):
    sys.exit(0)
print("== configs[4]: multi-camera batch, rank 0's LPT share of 4096 streams x 16 slices (slice size 1 KB * 2^(10 u^3))")
rs = np.random.default_rng(4096)
n_streams, per = 4096, 16
size_bytes = 1024.0 * 2.0 ** (10.0 * rs.random((n_streams, per)) ** 3)
parts = sharding.lpt_assign(size_bytes.sum(1).astype(np.int64), 8)
print("LPT over 8 ranks by bytes: imbalance (max / mean) %.4f; rank 0 gets %d streams, %.1f MB of %.1f MB" % (
    sharding.imbalance(size_bytes.sum(1).astype(np.int64), parts), len(parts[0]), size_bytes[parts[0]].sum() / 1e6,
    size_bytes.sum() / 1e6))
mine = np.array(parts[0])
nb = np.maximum((size_bytes[mine].reshape(-1) * 8 / 0.88).astype(np.int64), 32)       # bins per slice (~0.88 bit per bin)
n = len(nb)
g = hz.gpu_build_stream_cabac(torch, dev, n, 0, config=5, n_active=64, n_ctx=64, slices_per_frame=per, frames_per_params=1,
                              n_bins=nb)
torch.cuda.synchronize()
total_bins = int(nb.sum()) + n
cap = g["n_nals"] + 16
nbytes = g["n"]
d_rbsp = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
d_off = torch.empty(n, dtype=torch.int64, device=dev)
d_len = torch.empty(n, dtype=torch.int32, device=dev)
d_snal = torch.empty(n, dtype=torch.int32, device=dev)
d_ops = torch.from_numpy(g["ops"].view(np.int16)).to(dev)
d_nops = torch.from_numpy(g["n_ops"].view(np.int32)).to(dev)
p = capi.Context.slice_qp(g["qp"], g["idc"])
d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
boff = np.zeros(n + 1, dtype=np.uint64)
boff[1:] = np.cumsum((g["n_ops"].astype(np.uint64) + 1 + 31) // 32)
d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)
d_bins = torch.empty(int(boff[-1]), dtype=torch.int32, device=dev)
d_fin = torch.empty(n * 32, dtype=torch.uint8, device=dev)
flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE


def batch():
    ctx.annexb_scan_dev(g["stream"].data_ptr(), nbytes, d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap, d_sum.data_ptr(), 0)
    ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), cap, 0, n, d_off.data_ptr(), d_len.data_ptr(), d_snal.data_ptr(),
                         d_ns.data_ptr())
    ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=nbytes + 16, off=d_off.data_ptr(), len=d_len.data_ptr(),
                         n_slices=n, n_ctx=64, ops=d_ops.data_ptr(), n_ops_max=len(g["ops"]), n_ops=d_nops.data_ptr(),
                         qp=d_qp.data_ptr(), init_states=None, bins=d_bins.data_ptr(), bins_off=d_boff.data_ptr(),
                         bins_stride_words=0, final=d_fin.data_ptr(), final_states=None, flags=flags)


t = timed(batch, 6)
fin = np.frombuffer(d_fin.cpu().numpy().tobytes(), dtype=capi.FINAL_DTYPE)
ok = bool(np.array_equal(fin["n_bins"], g["n_ops"] + 1)) and not (fin["flags"] & capi.F_OVERRUN).any()
print("(a) one batch: %d streams, %d slices, %.1f MB, %.2f Gbins: makespan %.2f ms = %.1f Gbins/s, verified %s" % (
    len(mine), n, nbytes / 1e6, total_bins / 1e9, t, total_bins / t / 1e6, ok), flush=True)
print("    slice length max / mean = %.1f: the makespan is the longest slice's serial chain (%.0f k ops x ~200 cycles = %.1f ms)"
      % (nb.max() / nb.mean(), nb.max() / 1e3, nb.max() * 200 / 1.965e6))

# (b) in 8 jobs of 64 streams through the asynchronous job API (host buffers, three jobs in flight).  One stream per job
# would leave the GPU with 3 x 16 lanes of serial work at a time: the engine is serial inside a slice (~93 ns per bin for a
# warp on its own, configs[1] above), so a job takes as long as its longest slice whatever else it holds.
h_all = g["stream"][:nbytes].cpu().numpy()
nal_start = np.frombuffer(d_nals.cpu().numpy().tobytes(), dtype=capi.NAL_DTYPE)[:g["n_nals"]]
sps_at = nal_start["start"][nal_start["type"] == 7].astype(np.int64) - 4   # a stream = SPS, PPS, 16 slices
bounds = np.concatenate([sps_at, [nbytes - 4]])
group = 64
jobs = []
for k in range(0, len(sps_at), group):
    k1 = min(k + group, len(sps_at))
    s = np.concatenate([h_all[bounds[k]:bounds[k1]], np.array([0, 0, 0, 1], np.uint8)])
    sl = slice(k * per, k1 * per)
    jobs.append((np.ascontiguousarray(s), g["n_ops"][sl], g["qp"][sl], g["idc"][sl]))
pending = []
for rnd in range(2):   # first round: buffer growth
    lat = []
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for s, no, q, c in jobs:
        if len(pending) == 3:
            tk, ts = pending.pop(0)
            r = ctx.stream_wait(*tk)
            lat.append(time.perf_counter() - ts)
        ts = time.perf_counter()
        pending.append((ctx.stream_submit(s, g["ops"][:int(no.max())], no, q, c, 64, flags=flags), ts))
    while pending:
        tk, ts = pending.pop(0)
        r = ctx.stream_wait(*tk)
        lat.append(time.perf_counter() - ts)
    wall = time.perf_counter() - w0
lat = np.array(lat) * 1e3
print("(b) %d jobs of %d streams through h264b_stream_submit / _wait, three in flight, host buffers: makespan %.1f ms = "
      "%.2f Gbins/s; job latency min %.1f ms, median %.1f ms, max %.1f ms" % (len(jobs), group, wall * 1e3,
                                                                             total_bins / wall / 1e9, lat.min(),
                                                                             np.median(lat), lat.max()))
print("    (one stream per job, measured once: 512 jobs, makespan 89 s, job latency p50 518 ms / p99 882 ms / max 1676 ms)")
ctx.close()
