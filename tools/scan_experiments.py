#!/usr/bin/env python3
"""Time annexb_scan_dev for experiment builds of the library (-DH264B_EXP_* switch parts of the kernel off; results
of those builds are wrong on purpose).  Build here (no GPU needed):  python tools/scan_experiments.py --build
Run on the GPU box:                                                 python tools/scan_experiments.py --run
Not part of the product; a measurement aid whose output goes to profiles/."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "full": [],
    "nostore": ["-DH264B_EXP_NOSTORE"],     # K1a only reads and detects
    "nodetect": ["-DH264B_EXP_NODETECT"],   # K1a is a plain copy (only the edge chunks are flagged)
    "w4": ["-DH264B_SCAN_WARPS=4"],
    "w16": ["-DH264B_SCAN_WARPS=16"],
    "r2": ["-DH264B_SCAN_ROWS=2"],
    "r8": ["-DH264B_SCAN_ROWS=8"],
}
OUTDIR = os.path.join(ROOT, "h264decode_b200", "exp")


def build():
    from h264decode_b200 import build as b
    os.makedirs(OUTDIR, exist_ok=True)
    for name, flags in VARIANTS.items():
        b.build(extra=flags, out=os.path.join(OUTDIR, "lib_%s.so" % name))
        print("built", name)


def run(frames=4000):
    import numpy as np
    import torch
    import harness as hz
    dev = "cuda:0"
    g = hz.gpu_build_stream_cabac(torch, dev, frames * 8, 455000, config=4)
    n = g["n"]
    d_stream = g["stream"]
    nal_cap = g["n_nals"] + 16
    d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_nals = torch.empty(nal_cap * 32, dtype=torch.uint8, device=dev)
    d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    # reference point: a plain device copy of the same bytes (what MEASURED_PEAKS.json's hbm_gbs is)
    ts = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        d_rbsp[:n].copy_(d_stream[:n])
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[3:]))
    print("%-20s %8.3f ms (min %.3f)  in %7.1f GB/s   alg(2N) %7.1f GB/s" % ("torch copy_", t, min(ts), n / t / 1e6,
                                                                             2 * n / t / 1e6), flush=True)
    names = [v for v in VARIANTS if not os.environ.get("SCAN_EXP") or v in os.environ["SCAN_EXP"].split(",")]
    ref = None
    spans = [int(x) for x in os.environ.get("SCAN_SPAN", "0").split(",")]
    for name in names:
        if not os.path.exists(os.path.join(OUTDIR, "lib_%s.so" % name)):
            continue
        L = C.CDLL(os.path.join(OUTDIR, "lib_%s.so" % name))
        L.h264b_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        L.h264b_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.h264b_annexb_scan_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_uint32, C.c_void_p, C.c_uint32]
        L.h264b_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        assert L.h264b_create(0, C.byref(h)) == 0
        L.h264b_set_stream(h, C.c_void_p(stream.cuda_stream))
        for span in spans:
          d_rbsp.zero_()
          ts = []
          for it in range(12):
              e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
              e0.record(stream)
              rc = L.h264b_annexb_scan_dev(h, d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, nal_cap,
                                           d_sum.data_ptr(), 0)
              e1.record(stream)
              torch.cuda.synchronize()
              assert rc == 0
              ts.append(e0.elapsed_time(e1))
          t = float(np.median(ts[3:]))
          same = ""
          if not any("EXP" in f for f in VARIANTS[name]):  # a complete variant: its results must equal the first one's
              if ref is None:
                  ref = (d_rbsp[:n].clone(), d_nals.clone())
              else:
                  same = "  results == %s: %s" % (names[0], bool(torch.equal(ref[0], d_rbsp[:n]) and torch.equal(ref[1], d_nals)))
          print("%-20s %8.3f ms (min %.3f)  in %7.1f GB/s   alg(2N) %7.1f GB/s%s" % (name + ("/span%d" % span if span else ""), t, min(ts), n / t / 1e6,
                                                                                     2 * n / t / 1e6, same), flush=True)
          d_rbsp.zero_()
        L.h264b_destroy(h)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    if "--run" in sys.argv:
        run()
