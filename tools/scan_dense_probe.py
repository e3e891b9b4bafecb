"""Annex-B pass on an EPB-dense stream (BASELINE configs[0] payload statistics, ~0.5 % emulation-prevention bytes):
the worst case for the copy/dirty split -- almost every chunk goes through the general kernel and every NAL needs the
slide post-pass.  Measurement aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import harness as hz
from h264decode_b200 import capi
from oracle import oracle as orc
dev = "cuda:0"
one = hz.build_stream_c1(1 << 20)
reps = 512
s = torch.from_numpy(one).to(dev).repeat(reps)
n = s.numel()
d_stream = torch.cat([s, torch.zeros(64, dtype=torch.uint8, device=dev)])
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
cap = n // 64 + 1024
d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_nals = torch.empty(cap * 32, dtype=torch.uint8, device=dev)
d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.annexb_scan_dev(d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, cap, d_sum.data_ptr(), 0)
    e1.record(stream); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
summ = np.frombuffer(d_sum.cpu().numpy().tobytes()[:48], dtype=np.uint64, count=5)
onal, orbsp = orc.read_nal_units_arrays(one)
t = float(np.median(ts[2:]))
print("EPB-dense stream: %d MiB, %d NAL units, %d EPBs (%.2f %% of bytes): %.3f ms = %.1f GB/s of stream (%.1f GB/s algorithmic)"
      % (n >> 20, summ[1], summ[4], 100.0 * summ[4] / n, t, n / t / 1e6, (n + summ[2]) / t / 1e6))
print("expected per copy: %d NAL units (+1 per seam), %d rbsp bytes" % (len(onal["start"]), len(orbsp)))
assert summ[1] == (len(onal["start"]) + 1) * reps - 1
