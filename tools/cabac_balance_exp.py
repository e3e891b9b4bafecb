"""Measurement aid: CABAC kernel throughput with equal-length slices (no tail) as a function of the number of warps per
scheduler, next to the real length distribution.  Separates the bulk (pipe / latency) rate from the tail effect."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
dev = "cuda:0"
N = 80000
g = hz.gpu_build_stream_cabac(torch, dev, N, 455000, config=4, n_active=64, n_ctx=64, slices_per_frame=8,
                              frames_per_params=250, id_base=0, want_bins=False)
torch.cuda.synchronize()
n, d_stream, n_nals = g["n"], g["stream"], g["n_nals"]
ops, n_ops, qp, idc = g["ops"], g["n_ops"], g["qp"], g["idc"]
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
flags = capi.BYPASS_SPEC_OR
cap = n_nals + 16
d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
d_off = torch.empty(N, dtype=torch.int64, device=dev); d_len = torch.empty(N, dtype=torch.int32, device=dev)
d_snal = torch.empty(N, dtype=torch.int32, device=dev); d_ns = torch.zeros(4, dtype=torch.int32, device=dev)
ctx.annexb_scan_dev(d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap, d_sum.data_ptr(), 0)
ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), cap, 0, N, d_off.data_ptr(), d_len.data_ptr(), d_snal.data_ptr(), d_ns.data_ptr())
d_ops = torch.from_numpy(ops.view(np.int16)).to(dev)
p = capi.Context.slice_qp(qp, idc)
d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
d_fin = torch.empty(N * 32, dtype=torch.uint8, device=dev)
words = int(n_ops.max()) // 32 + 2

def run(ns, nops_arr, label):
    d_nops = torch.from_numpy(nops_arr.astype(np.uint32).view(np.int32)).to(dev)
    boff = np.zeros(ns + 1, dtype=np.uint64); boff[1:] = np.cumsum((nops_arr.astype(np.uint64) + 1 + 31) // 32)
    d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)
    d_bins = torch.empty(int(boff[-1]), dtype=torch.int32, device=dev)
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=n + 16, off=d_off.data_ptr(), len=d_len.data_ptr(),
                             n_slices=ns, n_ctx=64, ops=d_ops.data_ptr(), n_ops_max=len(ops), n_ops=d_nops.data_ptr(),
                             qp=d_qp.data_ptr(), init_states=None, bins=d_bins.data_ptr(), bins_off=d_boff.data_ptr(),
                             bins_stride_words=0, final=d_fin.data_ptr(), final_states=None, flags=flags)
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[1:])); bins = float(nops_arr.astype(np.int64).sum())
    print("%-44s %6d slices %8.2f ms  %7.1f Gbins/s  (%.2f warps/scheduler)" % (label, ns, t, bins / t / 1e6, ns / 32 / 592), flush=True)

run(N, n_ops, "real length distribution")
K = 70000
for ns in (80000, 75776, 56832, 37888, 18944, 9472):   # 592 * 32 * {4, 3, 2, 1, 0.5} and the full count
    run(ns, np.full(ns, K), "equal length %d ops" % K)
srt = np.sort(n_ops)[::-1]
print("length quantiles (max, p99, p90, p50, min) / mean:", [round(float(x) / n_ops.mean(), 2) for x in
      (srt[0], srt[N // 100], srt[N // 10], srt[N // 2], srt[-1])])
