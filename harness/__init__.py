"""Synthetic-data harness (test / bench input generation only; not the product, not the oracle).

Generators follow SURVEY.md §8(d): Annex-B streams with 4-byte start codes, emulation-prevention escaped
payloads, High-profile SPS/PPS templates the reference parses without panicking (Appendix B.3), and
spec-conformant CABAC slices produced by the H.264 9.3.4 arithmetic encoder (Appendix C) over a shared op
schedule.  The CPU build lives in harness.c; harness_gpu.cu is the same encoder as a CUDA kernel for the
multi-GB bench inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libharness.so")

TABLES_SPEC = 1
ESCAPE = 2
OP_DECISION, OP_BYPASS, OP_TERMINATE = 0, 1, 2
SC = bytes([0, 0, 0, 1])
SPS_NAL = bytes.fromhex("67640028ACD94078022640")   # Appendix B.3: High 100, level 40, 1920x1088
PPS_NAL = bytes.fromhex("68EE0F2C8B")               # Appendix B.3: CABAC, High-style tail


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("harness.c", "harness_core.h", "harness_tables.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-g", "-std=c11", "-fPIC", "-Wall", "-shared", "-pthread", "-o", _LIB_PATH,
                           srcs[0]])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hz_init_states.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hz_init_states.restype = None
        L.hz_gen_schedule.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.hz_gen_schedule.restype = None
        L.hz_gen_cabac_slices.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.hz_gen_cabac_slices.restype = C.c_int
        L.hz_encode_explicit.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_int64]
        L.hz_encode_explicit.restype = C.c_int64
        L.hz_random_payload.argtypes = [C.c_uint32, C.c_uint64, C.c_int64, C.c_void_p]
        L.hz_random_payload.restype = None
        L.hz_escape.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.hz_escape.restype = C.c_int64
        _lib = L
    return _lib


def make_op(kind, ctx=0):
    return (kind << 14) | (ctx & 0x3FF)


def init_states(qp, idc, n_ctx, flags=0):
    out = np.zeros(n_ctx, dtype=np.uint8)
    lib().hz_init_states(flags, int(qp), int(idc), n_ctx, out.ctypes.data)
    return out


def gen_schedule(config, n_ops, n_active, sched_id=0xFFF):
    ops = np.zeros(n_ops, dtype=np.uint16)
    lib().hz_gen_schedule(config, sched_id, n_ops, n_active, ops.ctypes.data)
    return ops


def slice_params(n_slices, first=0):
    """(qp, idc) per slice as in SURVEY.md §8(d) C3: qp = s % 52, idc = (s // 52) % 4 - 1."""
    s = np.arange(first, first + n_slices, dtype=np.int64)
    return (s % 52).astype(np.int32), ((s // 52) % 4 - 1).astype(np.int32)


def gen_cabac_slices(config, ops, n_ops, n_active, n_ctx, qp, idc, flags=0, id_base=0, stride=None, want_bins=True,
                     want_states=True, threads=None):
    """Encode len(n_ops) slices.  Returns dict(data[n, stride] uint8, lens int64, bins uint32[n, words] | None,
    final_states uint8[n, n_ctx] | None)."""
    ops = np.ascontiguousarray(ops, dtype=np.uint16)
    n_ops = np.ascontiguousarray(n_ops, dtype=np.uint32)
    n = len(n_ops)
    assert n_ops.max(initial=0) <= len(ops)
    qp = np.ascontiguousarray(qp, dtype=np.int32)
    idc = np.ascontiguousarray(idc, dtype=np.int32)
    if stride is None:
        stride = int(n_ops.max(initial=0)) * 2 // 8 * 2 + 64   # generous: <= 2 bits/bin, x2 for escaping
    data = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int64)
    words = int(n_ops.max(initial=0)) // 32 + 1
    bins = np.zeros((n, words), dtype=np.uint32) if want_bins else None
    fst = np.zeros((n, n_ctx), dtype=np.uint8) if want_states else None
    threads = threads or min(os.cpu_count() or 1, 64)
    ov = lib().hz_gen_cabac_slices(flags, config, id_base, n, ops.ctypes.data, n_ops.ctypes.data, n_active, n_ctx,
                                   qp.ctypes.data, idc.ctypes.data, data.ctypes.data, stride, lens.ctypes.data,
                                   bins.ctypes.data if want_bins else None, words,
                                   fst.ctypes.data if want_states else None, threads)
    if ov:
        raise RuntimeError("slice did not fit its stride")
    return dict(data=data, lens=lens, bins=bins, final_states=fst, stride=stride)


def encode_explicit(ops, bins, states, flags=0):
    """Encode an explicit list of (op word, bin) pairs -- data-dependent op sequences such as syntax elements -- plus the
    final terminate(1).  states: initial context states (uint8[n_ctx]).  -> (data uint8[], final states uint8[n_ctx])"""
    ops = np.ascontiguousarray(ops, dtype=np.uint16)
    bins = np.ascontiguousarray(bins, dtype=np.uint8)
    st = np.ascontiguousarray(states, dtype=np.uint8).copy()
    out = np.zeros(len(ops) * 2 + 64, dtype=np.uint8)
    n = lib().hz_encode_explicit(flags, ops.ctypes.data, bins.ctypes.data, len(ops), st.ctypes.data, len(st), out.ctypes.data,
                                 len(out))
    if n <= 0:
        raise RuntimeError("hz_encode_explicit: output too small")
    return out[:n].copy(), st


def random_payload(config, ident, n):
    out = np.zeros(n, dtype=np.uint8)
    lib().hz_random_payload(config, ident, n, out.ctypes.data)
    return out


def escape(payload):
    p = np.ascontiguousarray(payload, dtype=np.uint8)
    out = np.zeros(len(p) + len(p) // 2 + 2, dtype=np.uint8)
    n = lib().hz_escape(p.ctypes.data, len(p), out.ctypes.data)
    return out[:n]


def _rng_u64(config, ident, n):
    """n splitmix64 draws for (config, ident) -- numpy restatement of hz_next for sizes/lengths."""
    s = (0x4832363400000000 + config * 0x1000 + ident) & 0xFFFFFFFFFFFFFFFF
    out = np.zeros(n, dtype=np.uint64)
    M = 0xFFFFFFFFFFFFFFFF
    for i in range(n):
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        out[i] = z ^ (z >> 31)
    return out


def build_stream_c1(total_bytes=1 << 20, config=1, leading=b""):
    """SPS, PPS, then slice NALs (headers 0x65 / 0x41 alternating) with random
    emulation-prevention-bearing payloads of log-uniform length in [64 B, 64 KB] until total_bytes is reached,
    then a terminating start code.  Returns a uint8 array of exactly >= total_bytes (the last NAL is trimmed to fit
    when possible)."""
    parts = [bytes(leading), SC, SPS_NAL, SC, PPS_NAL]
    size = sum(len(p) for p in parts)
    draws = _rng_u64(config, 0xE00, 4096)
    k = 0
    while size + 4 < total_bytes:
        u = float(draws[k % len(draws)] >> np.uint64(11)) / float(1 << 53)
        ln = int(64 * (1024.0 ** u))
        room = total_bytes - 4 - size - 5
        if room < 8:
            break
        ln = min(ln, max(room * 2 // 3, 4))
        pay = escape(random_payload(config, k, ln))
        if len(pay) > room:
            pay = escape(random_payload(config, k, max(room * 2 // 3 - 2, 1)))
        parts += [SC, bytes([0x65 if k % 2 == 0 else 0x41]), pay.tobytes()]
        size += 5 + len(pay)
        k += 1
    parts.append(SC)
    size += 4
    buf = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    return buf


def assemble_annexb(nal_payloads, headers, params_every=None):
    """Concatenate NALs: start code + header byte + (already escaped) payload each; optional SPS+PPS every
    `params_every` NALs (and always first); terminating start code last."""
    parts = []
    for i, (h, p) in enumerate(zip(headers, nal_payloads)):
        if params_every is not None and i % params_every == 0:
            parts += [SC, SPS_NAL, SC, PPS_NAL]
        parts += [SC, bytes([h]), bytes(p)]
    parts.append(SC)
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def build_stream_cabac(n_slices, mean_bins, config=4, n_active=64, n_ctx=64, slices_per_frame=8, frames_per_params=250,
                       flags=0, id_base=0, sigma_frac=0.2, threads=None):
    """A C4-shaped stream at reduced size: slice NALs carrying encoder-generated CABAC data (escaped), SPS+PPS every
    frames_per_params frames.  Returns dict(stream, ops, n_ops, qp, idc, bins, final_states, n_active, n_ctx)."""
    d = _rng_u64(config, 0xD00 + id_base, 2 * n_slices).astype(np.float64) / float(1 << 64)
    # Box-Muller normal draws for slice sizes
    z = np.sqrt(-2.0 * np.log(np.maximum(d[0::2], 1e-12))) * np.cos(2 * np.pi * d[1::2])
    nb = np.clip(mean_bins * (1.0 + sigma_frac * z), mean_bins * 0.16, mean_bins * 2.56).astype(np.uint32)
    nb = np.maximum(nb, 1)
    ops = gen_schedule(config, int(nb.max()), n_active)
    qp, idc = slice_params(n_slices, first=id_base)
    g = gen_cabac_slices(config, ops, nb, n_active, n_ctx, qp, idc, flags=flags | ESCAPE, id_base=id_base,
                         threads=threads)
    pays = [g["data"][i, :g["lens"][i]].tobytes() for i in range(n_slices)]
    hdrs = [0x65 if (i % slices_per_frame == 0 and (i // slices_per_frame) % frames_per_params == 0) else 0x41
            for i in range(n_slices)]
    stream = assemble_annexb(pays, hdrs, params_every=slices_per_frame * frames_per_params)
    return dict(stream=stream, ops=ops, n_ops=nb, qp=qp, idc=idc, bins=g["bins"], final_states=g["final_states"],
                n_active=n_active, n_ctx=n_ctx, payload_lens=g["lens"])


# ------------------------------------------------------------------------------------------------ GPU generator
_GPU_LIB_PATH = os.path.join(_HERE, "_build", "libharness_gpu.so")
_gpu_lib = None


def build_gpu(force=False):
    """nvcc build of harness_gpu.cu (sm_100a); works without a GPU (cross-compile)."""
    srcs = [os.path.join(_HERE, f) for f in ("harness_gpu.cu", "harness_core.h", "harness_tables.h")]
    if (not force and os.path.exists(_GPU_LIB_PATH)
            and all(os.path.getmtime(_GPU_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _GPU_LIB_PATH
    os.makedirs(os.path.dirname(_GPU_LIB_PATH), exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", _GPU_LIB_PATH,
                           srcs[0]])
    return _GPU_LIB_PATH


def gpu_lib():
    global _gpu_lib
    if _gpu_lib is None:
        L = C.CDLL(build_gpu())
        vp, i64 = C.c_void_p, C.c_int64
        L.hzg_gen_cabac_slices.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, i64, vp, vp, C.c_uint32, C.c_uint32, vp,
                                           vp, vp, i64, vp, vp, i64, vp]
        L.hzg_gen_cabac_slices.restype = C.c_int
        L.hzg_random_payloads.argtypes = [C.c_int, C.c_uint32, C.c_uint64, i64, vp, vp, i64, vp]
        L.hzg_random_payloads.restype = C.c_int
        L.hzg_assemble.argtypes = [C.c_int, vp, i64, vp, vp, vp, i64, vp, vp, C.c_int, i64, i64]
        L.hzg_assemble.restype = C.c_int
        _gpu_lib = L
    return _gpu_lib


def slice_bins_normal(n_slices, mean_bins, config, id_base=0, sigma_frac=0.2, lo=0.16, hi=2.56):
    """per-slice bin counts ~ N(mean, sigma) clipped (50 KB, sigma 10 KB, [8 KB, 128 KB])"""
    rs = np.random.RandomState((0x48323634 + config * 0x1000 + id_base) & 0x7FFFFFFF)
    z = rs.standard_normal(n_slices)
    return np.maximum(np.clip(mean_bins * (1.0 + sigma_frac * z), mean_bins * lo, mean_bins * hi), 1).astype(np.uint32)


def gpu_build_stream_cabac(torch, device, n_slices, mean_bins, config=4, n_active=64, n_ctx=64, slices_per_frame=8,
                           frames_per_params=250, flags=0, id_base=0, want_bins=False, n_bins=None):
    """C4-shaped stream generated on the GPU (torch tensors for memory only).  Returns dict with device tensors
    stream (uint8, padded), n (int), ops, n_ops, qp, idc (host numpy + device), bins (device or None), payload lens."""
    L = gpu_lib()
    nb = slice_bins_normal(n_slices, mean_bins, config, id_base) if n_bins is None else np.asarray(n_bins, np.uint32)
    ops = gen_schedule(config, int(nb.max()), n_active)
    qp, idc = slice_params(n_slices, first=id_base)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        d_ops = torch.from_numpy(ops.view(np.int16)).to(dev)
        d_nops = torch.from_numpy(nb.view(np.int32)).to(dev)
        d_qp = torch.from_numpy(qp).to(dev)
        d_idc = torch.from_numpy(idc).to(dev)
        # <= ~1.02 bits per bin on average for this mix, plus escaping and flush; 1.4 bits/bin + 256 is generous
        stride = int(int(nb.max()) * 1.4 / 8) + 256
        stride = (stride + 15) // 16 * 16
        d_data = torch.empty((n_slices, stride), dtype=torch.uint8, device=dev)
        d_lens = torch.zeros(n_slices, dtype=torch.int64, device=dev)
        words = int(nb.max()) // 32 + 1
        d_bins = torch.zeros((n_slices, words), dtype=torch.int32, device=dev) if want_bins else None
        d_states = torch.empty((n_slices, n_ctx), dtype=torch.uint8, device=dev)
        rc = L.hzg_gen_cabac_slices(dev.index or 0, flags | ESCAPE, config, id_base, n_slices, d_ops.data_ptr(), d_nops.data_ptr(),
                                    n_active, n_ctx, d_qp.data_ptr(), d_idc.data_ptr(), d_data.data_ptr(), stride,
                                    d_lens.data_ptr(), d_bins.data_ptr() if want_bins else None, words,
                                    d_states.data_ptr())
        if rc != 0:
            raise RuntimeError("hzg_gen_cabac_slices rc=%d" % rc)
        pe = slices_per_frame * frames_per_params
        pre = SC + SPS_NAL + SC + PPS_NAL
        idx = torch.arange(n_slices, device=dev)
        nal_sizes = d_lens + 5 + torch.where(idx % pe == 0, len(pre), 0)
        ends = torch.cumsum(nal_sizes, 0)
        pos = ends - (d_lens + 5)                      # position of each slice NAL's start code
        total = int(ends[-1].item()) + 4
        d_stream = torch.zeros(((total + 64 + 15) // 16) * 16, dtype=torch.uint8, device=dev)
        hdr = np.where((np.arange(n_slices) % pe) == 0, 0x65, 0x41).astype(np.uint8)
        d_hdr = torch.from_numpy(hdr).to(dev)
        d_pre = torch.from_numpy(np.frombuffer(pre, np.uint8).copy()).to(dev)
        rc = L.hzg_assemble(dev.index or 0, d_data.data_ptr(), stride, d_lens.data_ptr(), pos.data_ptr(), d_hdr.data_ptr(), n_slices,
                            d_stream.data_ptr(), d_pre.data_ptr(), len(pre), pe, total - 4)
        if rc != 0:
            raise RuntimeError("hzg_assemble rc=%d" % rc)
        lens = d_lens.cpu().numpy()
        del d_data
    return dict(stream=d_stream, n=total, ops=ops, n_ops=nb, qp=qp, idc=idc, bins=d_bins, final_states=d_states,
                payload_lens=lens, n_active=n_active, n_ctx=n_ctx, n_nals=n_slices + 2 * ((n_slices + pe - 1) // pe))
