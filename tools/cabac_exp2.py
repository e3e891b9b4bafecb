"""Measurement aid (round 2): CABAC kernel variants (fast loop, warps per CTA, bundle -> warp mapping) on the real
configs[3] length distribution and on equal-length slices at 4 / 1 / 0.5 warps per scheduler.  The knobs are the
environment variables launch_cabac reads per call (H264B_CABAC_LOOP / _W / _MAP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
dev = "cuda:0"
N = int(os.environ.get("EXP_SLICES", "80000"))
MEAN = int(os.environ.get("EXP_MEAN_BINS", "455000"))
N_ACTIVE = int(os.environ.get("EXP_ACTIVE", "64"))      # active contexts of the schedule (64, repeat with 460)
N_CTX = int(os.environ.get("EXP_NCTX", str(N_ACTIVE)))   # context rows per slice
g = hz.gpu_build_stream_cabac(torch, dev, N, MEAN, config=4, n_active=N_ACTIVE, n_ctx=N_CTX, slices_per_frame=8,
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
ref_fin = {}

def run(ns, nops_arr, label, reps=3, check_key=None):
    d_nops = torch.from_numpy(nops_arr.astype(np.uint32).view(np.int32)).to(dev)
    boff = np.zeros(ns + 1, dtype=np.uint64); boff[1:] = np.cumsum((nops_arr.astype(np.uint64) + 1 + 31) // 32)
    d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)
    d_bins = torch.zeros(int(boff[-1]), dtype=torch.int32, device=dev)
    ts = []
    for it in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=n + 16, off=d_off.data_ptr(), len=d_len.data_ptr(),
                             n_slices=ns, n_ctx=N_CTX, ops=d_ops.data_ptr(), n_ops_max=len(ops), n_ops=d_nops.data_ptr(),
                             qp=d_qp.data_ptr(), init_states=None, bins=d_bins.data_ptr(), bins_off=d_boff.data_ptr(),
                             bins_stride_words=0, final=d_fin.data_ptr(), final_states=None, flags=flags,
                             n_ctx_used=N_ACTIVE if N_ACTIVE < N_CTX else 0)
        e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[1:])); bins = float(nops_arr.astype(np.int64).sum())
    ok = ""
    if check_key is not None:   # every variant must produce the same bins and final records as the first one run
        sig = (int(d_bins.to(torch.int64).sum().item()), int(d_fin[:ns * 32].to(torch.int64).sum().item()))
        if check_key in ref_fin:
            ok = "  same" if ref_fin[check_key] == sig else "  DIFFERENT RESULT"
        else:
            ref_fin[check_key] = sig
    print("%-40s %6d slices %8.2f ms  %7.1f Gbins/s  (%.2f warps/sched, %.0f cycles/op/warp)%s" % (
        label, ns, t, bins / t / 1e6, ns / 32 / 592, t * 1e-3 * 1.965e9 / (bins / ns) if len(set(nops_arr.tolist())) == 1 else 0, ok), flush=True)

variants = [v for v in os.environ.get("EXP_VARIANTS", "0:2:0,1:0:1,2:0:1").split(",")]
K = int(os.environ.get("EXP_K", "70000"))
only = os.environ.get("EXP_ONLY")   # "ns": one equal-length configuration only (the ncu capture)
for v in variants:
    loop, w, mp = v.split(":")
    os.environ["H264B_CABAC_LOOP"], os.environ["H264B_CABAC_W"], os.environ["H264B_CABAC_MAP"] = loop, w, mp
    print("--- loop %s, warps/CTA %s (0 = one wave), map %s, %d active contexts, %d context rows" % (loop, w, mp, N_ACTIVE, N_CTX), flush=True)
    if only:
        run(int(only), np.full(int(only), K), "equal length %d ops" % K, reps=1)
        continue
    run(N, n_ops, "real length distribution", check_key="real")
    for ns in (75776, 18944, 9472):
        if ns <= N:
            run(ns, np.full(ns, K), "equal length %d ops" % K, reps=2, check_key="eq%d" % ns)
