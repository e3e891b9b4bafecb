"""Measurement aid: what per-lane, data-dependent op sequences cost.  The same number of slices decodes
(a) mb_type syntax elements through h264b_mb_type_decode_dev (every lane walks the binarisation trie on its own: its
    contexts, its op kinds and the length of its elements depend on its bins), and
(b) about as many bins on one shared op schedule through h264b_cabac_decode_dev (the fast loop).
Prints bins/s of both and the ratio."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
from oracle import oracle as orc
import test_mb_type as tm

dev = "cuda:0"
rng = np.random.default_rng(1)
n_distinct, n_slices, n_ctx, n_mb = 512, int(os.environ.get("EXP_SLICES", "75776")), 32, int(os.environ.get("EXP_MB", "4000"))
qp, idc = hz.slice_params(n_distinct, first=0)
init = orc.ctx_init(qp, idc, n_ctx)
datas, kinds, bins_per = [], [], []
for s in range(n_distinct):
    kind = s & 1
    t = tm.random_types(rng, kind, n_mb, False)
    ops, bins = tm.ops_for(kind, t)
    d, _ = hz.encode_explicit(ops, bins, init[s])
    datas.append(d); kinds.append(kind); bins_per.append(len(ops))
stride = (max(len(d) for d in datas) + 19) // 4 * 4
buf = np.zeros((n_distinct, stride), np.uint8)
for s, d in enumerate(datas):
    buf[s, :len(d)] = d
rep = (n_slices + n_distinct - 1) // n_distinct
sel = np.tile(np.arange(n_distinct), rep)[:n_slices]
off = (sel.astype(np.uint64) * stride)
length = np.array([len(datas[k]) for k in sel], np.uint32)
total_bins = int(sum(bins_per[k] for k in sel))
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
d_buf = torch.from_numpy(buf.reshape(-1)).to(dev)
d_off = torch.from_numpy(off.view(np.int64)).to(dev); d_len = torch.from_numpy(length.view(np.int32)).to(dev)
d_kind = torch.from_numpy(np.array([kinds[k] for k in sel], np.uint8)).to(dev)
d_nmb = torch.full((n_slices,), n_mb, dtype=torch.int32, device=dev)
p = capi.Context.slice_qp(qp[sel], idc[sel])
d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
d_out = torch.zeros((n_slices, n_mb), dtype=torch.uint8, device=dev)
d_fin = torch.zeros(n_slices * 40, dtype=torch.uint8, device=dev)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

t_mb = timed(lambda: ctx.mb_type_decode_dev(bytes=d_buf.data_ptr(), total_bytes=buf.size, off=d_off.data_ptr(), len=d_len.data_ptr(),
                                            n_slices=n_slices, n_ctx=n_ctx, slice_kind=d_kind.data_ptr(), n_mb=d_nmb.data_ptr(),
                                            n_mb_max=n_mb, flags=0, qp=d_qp.data_ptr(), init_states=None, mb_type=d_out.data_ptr(),
                                            final=d_fin.data_ptr(), final_states=None))
fin = np.frombuffer(d_fin.cpu().numpy().tobytes(), dtype=capi.MB_FINAL_DTYPE)
assert (fin["n_mb"] == n_mb).all() and int(fin["n_bins"].astype(np.int64).sum()) == total_bins
print("mb_type walk (per-lane sequences): %d slices x %d elements, %.2f bins per element: %8.2f ms  %7.1f Gbins/s" % (
    n_slices, n_mb, total_bins / n_slices / n_mb, t_mb, total_bins / t_mb / 1e6), flush=True)

# (b) the shared-schedule engine on as many bins per slice
K = total_bins // n_slices
g = hz.gpu_build_stream_cabac(torch, dev, 2048, K, config=4, n_active=16, n_ctx=32, slices_per_frame=8, frames_per_params=250,
                              id_base=0, want_bins=False, n_bins=np.full(2048, K, np.uint32))
torch.cuda.synchronize()
n, cap = g["n"], g["n_nals"] + 16
d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev); d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
o2 = torch.empty(2048, dtype=torch.int64, device=dev); l2 = torch.empty(2048, dtype=torch.int32, device=dev)
sn = torch.empty(2048, dtype=torch.int32, device=dev); ns = torch.zeros(4, dtype=torch.int32, device=dev)
ctx.annexb_scan_dev(g["stream"].data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap, d_sum.data_ptr(), 0)
ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), cap, 0, 2048, o2.data_ptr(), l2.data_ptr(), sn.data_ptr(), ns.data_ptr())
idx = torch.from_numpy(np.tile(np.arange(2048), (n_slices + 2047) // 2048)[:n_slices]).to(dev)
d_off2, d_len2 = o2[idx].contiguous(), l2[idx].contiguous()
pq = capi.Context.slice_qp(g["qp"], g["idc"])
d_qp2 = torch.from_numpy(pq.view(np.int32).reshape(-1, 2).copy()).to(dev)[idx].contiguous()
d_ops = torch.from_numpy(g["ops"].view(np.int16)).to(dev)
d_nops = torch.full((n_slices,), K, dtype=torch.int32, device=dev)
words = K // 32 + 2
d_bins = torch.zeros(n_slices * words, dtype=torch.int32, device=dev)
d_fin2 = torch.zeros(n_slices * 32, dtype=torch.uint8, device=dev)
t_sh = timed(lambda: ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=n + 16, off=d_off2.data_ptr(), len=d_len2.data_ptr(),
                                          n_slices=n_slices, n_ctx=32, ops=d_ops.data_ptr(), n_ops_max=len(g["ops"]),
                                          n_ops=d_nops.data_ptr(), qp=d_qp2.data_ptr(), init_states=None, bins=d_bins.data_ptr(),
                                          bins_off=None, bins_stride_words=words, final=d_fin2.data_ptr(), final_states=None,
                                          flags=capi.BYPASS_SPEC_OR))
print("shared op schedule (fast loop):    %d slices x %d bins:                           %8.2f ms  %7.1f Gbins/s" % (
    n_slices, K, t_sh, n_slices * K / t_sh / 1e6))
print("per-lane sequences cost a factor %.1f in bins/s" % ((n_slices * K / t_sh) / (total_bins / t_mb)))
ctx.close()
