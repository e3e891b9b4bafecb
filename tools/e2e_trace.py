import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
import bench
dev = "cuda:0"
n_slices = 80000
g = hz.gpu_build_stream_cabac(torch, dev, n_slices, bench.MEAN_BINS, config=4, n_active=64, n_ctx=64, slices_per_frame=8,
                              frames_per_params=250, want_bins=False)
ctx = capi.Context(0)
n = g["n"]
h_stream = ctx.host_alloc(n)
ctx.d2h(h_stream, g["stream"].data_ptr()); ctx.sync()
ops, n_ops = g["ops"], g["n_ops"]
p = capi.Context.slice_qp(g["qp"], g["idc"])
flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
t = [bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags) for _ in range(3)]
for x in t: bench._stream_wait_raw(ctx, capi, x)
print("---- timed", file=sys.stderr, flush=True)
t0 = time.perf_counter()
def stamp(what): print("%8.1f ms  %s" % ((time.perf_counter() - t0) * 1e3, what), flush=True)
pend = []
for k in range(9):
    if len(pend) == 3:
        bench._stream_wait_raw(ctx, capi, pend.pop(0)); stamp("wait returned")
    pend.append(bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags)); stamp("submit %d returned" % k)
while pend:
    bench._stream_wait_raw(ctx, capi, pend.pop(0)); stamp("wait returned")
