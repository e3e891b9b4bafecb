"""Time the phases of pipelined h264b_stream_submit / _wait calls (measurement aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
import bench
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
dev = "cuda:0"
n_slices = frames * 8
g = hz.gpu_build_stream_cabac(torch, dev, n_slices, bench.MEAN_BINS, config=4, n_active=64, n_ctx=64, slices_per_frame=8,
                              frames_per_params=250, want_bins=False)
ctx = capi.Context(0)
n = g["n"]
h_stream = ctx.host_alloc(n)
ctx.d2h(h_stream, g["stream"].data_ptr()); ctx.sync()
ops, n_ops = g["ops"], g["n_ops"]
p = capi.Context.slice_qp(g["qp"], g["idc"])
flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
for _ in range(2):
    t = [bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags) for _ in range(2)]
    for x in t: bench._stream_wait_raw(ctx, capi, x)
t0 = time.perf_counter()
def stamp(what): print("%8.1f ms  %s" % ((time.perf_counter() - t0) * 1e3, what), flush=True)
pend = [bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags)]; stamp("submit 0 returned")
for k in range(1, 5):
    pend.append(bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags)); stamp("submit %d returned" % k)
    bench._stream_wait_raw(ctx, capi, pend.pop(0)); stamp("wait %d returned" % (k - 1))
bench._stream_wait_raw(ctx, capi, pend.pop(0)); stamp("wait 4 returned")
# sequential for comparison
t0 = time.perf_counter()
for k in range(3):
    x = bench._stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags); bench._stream_wait_raw(ctx, capi, x); stamp("sequential job %d done" % k)
# ---- phases alone
import ctypes as C
d_stream = g["stream"]
torch.cuda.synchronize()
def timeit(f, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); ctx.sync(); a = time.perf_counter(); f(); ctx.sync(); torch.cuda.synchronize(); ts.append((time.perf_counter() - a) * 1e3)
    return min(ts)
d_tmp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
print("H2D stream alone        %.1f ms" % timeit(lambda: ctx.h2d(d_tmp.data_ptr(), h_stream)))
nal_cap = g["n_nals"] + 16
d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_nals = torch.empty(nal_cap * 32, dtype=torch.uint8, device=dev)
d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
print("scan alone              %.1f ms" % timeit(lambda: ctx.annexb_scan_dev(d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, nal_cap, d_sum.data_ptr(), 0)))
d_off = torch.empty(n_slices, dtype=torch.int64, device=dev); d_len = torch.empty(n_slices, dtype=torch.int32, device=dev)
d_snal = torch.empty(n_slices, dtype=torch.int32, device=dev); d_ns = torch.zeros(4, dtype=torch.int32, device=dev)
ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), nal_cap, 0, n_slices, d_off.data_ptr(), d_len.data_ptr(), d_snal.data_ptr(), d_ns.data_ptr())
d_ops = torch.from_numpy(ops.view(np.int16)).to(dev); d_nops = torch.from_numpy(n_ops.view(np.int32)).to(dev)
d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
boff = np.zeros(n_slices + 1, dtype=np.uint64); boff[1:] = np.cumsum((n_ops.astype(np.uint64) + 1 + 31) // 32)
d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)
d_bins = torch.empty(int(boff[-1]), dtype=torch.int32, device=dev); d_fin = torch.empty(n_slices * 32, dtype=torch.uint8, device=dev)
def cab():
    ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=n + 16, off=d_off.data_ptr(), len=d_len.data_ptr(),
                         n_slices=n_slices, n_ctx=64, ops=d_ops.data_ptr(), n_ops_max=len(ops), n_ops=d_nops.data_ptr(),
                         qp=d_qp.data_ptr(), init_states=None, bins=d_bins.data_ptr(), bins_off=d_boff.data_ptr(),
                         bins_stride_words=0, final=d_fin.data_ptr(), final_states=None, flags=flags)
print("cabac alone             %.1f ms" % timeit(cab))
h_bins = ctx.host_alloc(int(boff[-1]) * 4)
print("D2H bins alone          %.1f ms" % timeit(lambda: ctx.d2h(h_bins, d_bins.data_ptr())))
