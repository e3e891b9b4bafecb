"""CPU checks of the product's __host__ __device__ logic (annexb_local.cuh, cabac_lane.cuh) against the oracle.

tests/native/hd_emul.cpp wraps the same predicates and per-lane arithmetic the CUDA kernels execute; compiling it
with g++ lets the position-local split/strip rules and the 64-bit-window CABAC lane be compared with the oracle's
sequential restatement here, without a GPU.  (The GPU parity tests proper are in test_gpu_*.py.)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "hd_emul.cpp")
OUT = os.path.join(HERE, "native", "_build", "libhd_emul.so")


class EmulFinal(C.Structure):
    _fields_ = [("R", C.c_int64), ("O", C.c_int64), ("bits_read", C.c_uint64), ("overrun", C.c_uint32),
                ("n_bins", C.c_uint32)]


@pytest.fixture(scope="module")
def emul():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(HERE, "..", "h264decode_b200", "csrc", f)
                    for f in ("annexb_local.cuh", "cabac_lane.cuh", "tables.inc")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT])
    L = C.CDLL(OUT)
    L.emul_stream.restype = C.c_int64
    L.emul_stream.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                              C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.emul_stream_tiles.restype = C.c_int64
    L.emul_stream_tiles.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int64, C.c_uint32, C.c_void_p]
    L.emul_stream_carry.restype = C.c_int64
    L.emul_stream_carry.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p]
    L.emul_frame.restype = C.c_int64
    L.emul_frame.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.emul_cabac.restype = None
    L.emul_cabac.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p,
                             C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(EmulFinal)]
    return L


def run_emul_stream(L, s):
    s = np.ascontiguousarray(s, dtype=np.uint8)
    cap = len(s) // 4 + 2
    st = np.zeros(cap, np.uint64)
    ro = np.zeros(cap, np.uint64)
    hd = np.zeros(cap, np.uint32)
    rb = np.zeros(len(s) + 16, np.uint8)
    tot, e0, mism = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    K = L.emul_stream(s.ctypes.data, len(s), st.ctypes.data, ro.ctypes.data, hd.ctypes.data, cap, rb.ctypes.data,
                      C.byref(tot), C.byref(e0), C.byref(mism))
    return K, st[:K], ro[:K], hd[:K], rb[:tot.value], e0.value, mism.value


def check_stream(L, s):
    K, st, ro, hd, rb, e0, mism = run_emul_stream(L, s)
    nal, rbsp = orc.read_nal_units_arrays(s)
    # the kernel's tile pipeline (position-preserving RBSP layout: each NAL's RBSP at its body's own position)
    s8 = np.ascontiguousarray(s, dtype=np.uint8)
    n_or = len(nal["start"])
    # (destination shift, chunks per piece, piece visiting order): pieces of 1, 2, 3 and 64 chunks, shuffled or not
    for shift, span, seed in ((0, 1, 0), (16, 2, 7), (5, 3, 1234567), (0, 64, 99)):
        out = np.full(len(s8) + 96 + shift, 0xEE, np.uint8)
        cap = len(s8) // 4 + 2
        t_st, t_epb, t_hd = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        stats = np.zeros(4, np.int64)
        Kt = L.emul_stream_tiles(s8.ctypes.data, len(s8), out.ctypes.data, shift, t_st.ctypes.data, t_epb.ctypes.data,
                                 t_hd.ctypes.data, cap, span, seed, stats.ctypes.data)
        assert stats[3] == 0, "zero-pair filter disagrees with its definition"
        assert max(Kt, 1) - 1 == n_or
        written = np.zeros(len(out), bool)
        for k in range(n_or):
            a0 = int(t_st[k])
            assert a0 == nal["start"][k]
            H = int(nal["header_bytes"][k])
            ln = max(int(t_st[k + 1]) - a0 - H - 2, 0) - int(t_epb[k + 1])
            assert ln == nal["rbsp_len"][k]
            got = out[shift + a0 + H: shift + a0 + H + ln]
            exp = rbsp[int(nal["rbsp_off"][k]): int(nal["rbsp_off"][k]) + ln]
            assert np.array_equal(got, exp), "tile pipeline bytes differ (NAL %d, shift %d)" % (k, shift)
            written[shift + a0 + H: shift + a0 + H + ln] = True
        # nothing may be written outside [first NAL body, stream end) and nothing before the destination
        assert np.all(out[:shift] == 0xEE) and np.all(out[shift + len(s8):] == 0xEE), "stray writes"
    check_carry_pipeline(L, s8, nal, rbsp)
    assert mism == 0
    assert max(K, 1) - 1 == len(nal["start"])
    n = len(nal["start"])
    if n:
        assert np.array_equal(st[:n].astype(np.int64), nal["start"])
        assert np.array_equal((st[1:n + 1] - st[:n]).astype(np.int64), nal["num_bytes"])
        assert np.array_equal(ro[:n].astype(np.int64), nal["rbsp_off"])
        assert np.array_equal((ro[1:n + 1] - ro[:n]).astype(np.int64), nal["rbsp_len"])
        assert np.array_equal((hd[:n] & 31).astype(np.int64), nal["type"])
        assert np.array_equal(rb[:int(ro[n])], rbsp)
        assert e0 == nal["start"][0]
    else:
        assert len(rbsp) == 0


def check_carry_pipeline(L, s8, nal, rbsp):
    """the round-2 pipeline (counts carried from chunk to chunk, image built in place, re-copy list): same outputs"""
    n_or = len(nal["start"])
    stats_all = np.zeros(4, np.int64)
    for shift in (0, 16):
        out = np.full(len(s8) + 96 + shift, 0xEE, np.uint8)
        cap = len(s8) // 4 + 2
        t_st, t_epb, t_hd = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        stats = np.zeros(4, np.int64)
        Kt = L.emul_stream_carry(s8.ctypes.data, len(s8), out.ctypes.data, shift, t_st.ctypes.data, t_epb.ctypes.data,
                                 t_hd.ctypes.data, cap, stats.ctypes.data)
        assert stats[3] == 0, "a chunk looked back at one that had not published (ticket order broken)"
        assert max(Kt, 1) - 1 == n_or
        for k in range(n_or):
            a0 = int(t_st[k])
            assert a0 == nal["start"][k]
            H = int(nal["header_bytes"][k])
            ln = max(int(t_st[k + 1]) - a0 - H - 2, 0) - int(t_epb[k + 1])
            assert ln == nal["rbsp_len"][k], (k, ln, int(nal["rbsp_len"][k]))
            got = out[shift + a0 + H: shift + a0 + H + ln]
            exp = rbsp[int(nal["rbsp_off"][k]): int(nal["rbsp_off"][k]) + ln]
            assert np.array_equal(got, exp), "carry pipeline bytes differ (NAL %d, shift %d)" % (k, shift)
        assert np.all(out[:shift] == 0xEE) and np.all(out[shift + len(s8):] == 0xEE), "stray writes"
        stats_all += stats
    return stats_all


def random_stream(rng, n, p_zero, p_sc, ext_types=False):
    """bytes with many zeros / 1s / 3s, random start codes, optional extension NAL types after them"""
    vals = np.array([0, 1, 2, 3, 0x6E, 0x74, 0x75, 0xF5, 0x80, 0x7F, 0xAA], dtype=np.uint8)
    s = rng.integers(0, 256, n).astype(np.uint8)
    m = rng.random(n) < p_zero
    s[m] = 0
    m = rng.random(n) < 0.15
    s[m] = vals[rng.integers(0, len(vals), m.sum())]
    for p in np.flatnonzero(rng.random(n) < p_sc):
        if p + 6 < n:
            s[p:p + 4] = [0, 0, 0, 1]
            if ext_types:
                s[p + 4] = [0x6E, 0x74, 0x75, 0x65, 0x00, 0x41][rng.integers(0, 6)]
                s[p + 5] = [0x80, 0x00, 0xFF, 0x7F][rng.integers(0, 4)]
    return s


@pytest.mark.parametrize("seed", range(6))
def test_local_split_strip_matches_oracle_random(emul, seed):
    rng = np.random.default_rng(seed)
    for n in [0, 1, 3, 4, 5, 15, 16, 17, 31, 33, 100, 1000, 5000, 16383, 16385, 40000]:
        for p_zero, p_sc in [(0.5, 0.02), (0.2, 0.05), (0.9, 0.01), (0.05, 0.002)]:
            check_stream(emul, random_stream(rng, n, p_zero, p_sc, ext_types=(seed % 2 == 0)))


def test_local_split_strip_kats(emul):
    from tests.test_oracle_kat import B2_STREAM
    check_stream(emul, np.frombuffer(B2_STREAM, np.uint8))
    for hexs in ["00000001", "0000000100000001", "000000014100000000010000000165", "00 00 00 00 00 01 00 00 03 00 00 00 01",
                 "00000001 6E 80 00 00 00 03 55 66 77 88 00 00 00 01",
                 "00000001 75 FF 00 00 03 00 00 03 00 00 00 01 75 7F 00 00 03 01 00 00 00 01"]:
        check_stream(emul, np.frombuffer(bytes.fromhex(hexs), np.uint8))
    check_stream(emul, np.zeros(100, np.uint8))
    check_stream(emul, np.tile(np.array([0, 0, 3], np.uint8), 50))
    check_stream(emul, np.tile(np.array([0, 0, 0, 1], np.uint8), 40))
    check_stream(emul, np.concatenate([np.array([0, 0, 0, 1, 0x65], np.uint8), np.tile(np.array([0, 0, 3], np.uint8), 40),
                                       np.array([0, 0, 0, 1], np.uint8)]))
    # one NAL over several pieces that loses an EPB every third byte / only in its first tile / only late
    sc, hdr = np.array([0, 0, 0, 1], np.uint8), np.array([0x65], np.uint8)
    rng = np.random.default_rng(77)
    body = rng.integers(4, 256, 60000).astype(np.uint8)
    check_stream(emul, np.concatenate([sc, hdr, np.tile(np.array([0, 0, 3], np.uint8), 23000), sc, hdr, body, sc]))
    b2 = body.copy(); b2[100:103] = [0, 0, 3]
    check_stream(emul, np.concatenate([sc, hdr, b2, sc]))
    b3 = body.copy(); b3[40000:40003] = [0, 0, 3]; b3[16380:16383] = [0, 0, 3]
    check_stream(emul, np.concatenate([np.full(7, 9, np.uint8), sc, hdr, b3, sc, hdr, b2[:20000], sc]))


def test_fast_path_is_taken_on_sparse_streams(emul):
    """entropy-coded-looking payload: most chunks must take the verbatim-copy path, and the result still matches"""
    rng = np.random.default_rng(5)
    s = rng.integers(0, 256, 600000).astype(np.uint8)
    s[:5] = [0, 0, 0, 1, 0x65]
    s[300000:300005] = [0, 0, 0, 1, 0x41]
    s[-4:] = [0, 0, 0, 1]
    check_stream(emul, s)
    out = np.zeros(len(s) + 96, np.uint8)
    cap = len(s) // 4 + 2
    a, b, c = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    stats = np.zeros(4, np.int64)
    emul.emul_stream_tiles(s.ctypes.data, len(s), out.ctypes.data, 0, a.ctypes.data, b.ctypes.data, c.ctypes.data, cap,
                           64, 3, stats.ctypes.data)
    n_chunks = (len(s) + 2047) // 2048
    assert stats[0] + stats[1] + stats[2] == n_chunks
    assert stats[0] > 0.9 * n_chunks and stats[2] < 0.05 * n_chunks, stats


def test_local_split_strip_harness_streams(emul):
    check_stream(emul, hz.build_stream_c1(1 << 17))
    check_stream(emul, hz.build_stream_c1(40000, leading=b"\x12\x00\x00\x01\x00"))
    check_stream(emul, hz.build_stream_cabac(9, 2000, slices_per_frame=3, frames_per_params=2)["stream"])


@pytest.mark.parametrize("seed", range(4))
def test_frame_mode_matches_new_nal_unit(emul, seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(400):
        n = int(rng.integers(1, 60))
        f = random_stream(rng, n, 0.5, 0.0)
        f[0] = [0x65, 0x6E, 0x74, 0x75, 0x00, 0x67][rng.integers(0, 6)]
        if rng.random() < 0.3 and n >= 3:
            f[-3:] = [0, 0, 3]
        st, nal, rbsp = orc.new_nal_unit(f)
        out = np.zeros(n + 1, np.uint8)
        k = emul.emul_frame(f.ctypes.data, n, out.ctypes.data)
        if st == orc.PANIC:
            assert nal["HeaderBytes"] > n or n == 0 or True   # header runs past the frame: reference panics
            continue
        assert bytes(out[:k]) == rbsp, f.tobytes().hex()


def _cabac_case(emul, data, off, ops, init, flags_o, final_term):
    _cabac_case1(emul, data, off, ops, init, flags_o, final_term, 0)
    if flags_o & orc.BYPASS_SPEC_OR:
        _cabac_case1(emul, data, off, ops, init, flags_o, final_term, 8)


def _cabac_case1(emul, data, off, ops, init, flags_o, final_term, extra):
    """run one slice through the lane emulation and the oracle, compare everything"""
    buf = np.zeros(off + len(data) + 64, np.uint8)
    buf[off:off + len(data)] = data
    buf[off + len(data):] = 0xA5            # bytes after the slice must never matter
    buf[:off] = 0x5A
    n_ops = len(ops)
    sl_ops = np.concatenate([ops, np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)]) if final_term else ops
    rc, obins, ofin, ost = orc.cabac_decode_slice(data, sl_ops, init, flags_o)
    st = init.copy()
    bins = np.zeros(len(sl_ops) // 32 + 2, np.uint32)
    fin = EmulFinal()
    ef = ((1 if flags_o & orc.TABLES_SPEC else 0) | (2 if flags_o & orc.BYPASS_SPEC_OR else 0) |
          (4 if final_term else 0) | extra)
    emul.emul_cabac(buf.ctypes.data, len(buf), off, len(data), np.ascontiguousarray(ops).ctypes.data, n_ops,
                    st.ctypes.data, len(st), ef, bins.ctypes.data, C.byref(fin))
    assert bool(fin.overrun) == (rc == orc.PANIC)
    if rc == orc.PANIC:
        nb = ofin["n_bins"]
        for i in range(nb):
            assert (bins[i >> 5] >> (i & 31)) & 1 == (obins[i >> 5] >> (i & 31)) & 1
        return
    nw = (len(sl_ops) + 31) // 32
    if len(sl_ops) % 32:
        bins[nw - 1] &= np.uint32((1 << (len(sl_ops) % 32)) - 1)
    assert np.array_equal(bins[:nw], obins[:nw])
    assert (fin.R, fin.O, fin.bits_read, fin.n_bins) == (ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"],
                                                         ofin["n_bins"])
    assert np.array_equal(st, ost)


@pytest.mark.parametrize("flags_o", [orc.BYPASS_SPEC_OR, orc.BYPASS_SPEC_OR | orc.TABLES_SPEC, 0, orc.TABLES_SPEC])
def test_cabac_lane_matches_oracle(emul, flags_o):
    n_active, n_ctx = 64, 64
    ops = hz.gen_schedule(2, 3000, n_active)
    qp, idc = hz.slice_params(8, first=7)
    fh = hz.TABLES_SPEC if flags_o & orc.TABLES_SPEC else 0
    n_ops = np.array([3000, 2999, 1, 31, 32, 33, 383, 1500], np.uint32)
    g = hz.gen_cabac_slices(2, ops, n_ops, n_active, n_ctx, qp, idc, flags=fh)
    init = orc.ctx_init(qp, idc, n_ctx, flags_o & orc.TABLES_SPEC)
    for s in range(8):
        data = g["data"][s, :g["lens"][s]].copy()
        for off in (0, 1, 2, 3, 4, 7):
            _cabac_case(emul, data, off, ops[:n_ops[s]], init[s], flags_o, True)
            _cabac_case(emul, data, off, ops[:n_ops[s]], init[s], flags_o, False)


@pytest.mark.parametrize("flags_o", [orc.BYPASS_SPEC_OR, 0])
def test_cabac_lane_random_bits_and_overrun(emul, flags_o):
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(1, 400))
        data = rng.integers(0, 256, n).astype(np.uint8)
        n_ops = int(rng.integers(1, 4000))
        kinds = rng.choice([0, 0, 0, 1, 2], n_ops)
        ctxs = rng.integers(0, 32, n_ops)
        ops = ((kinds << 14) | ctxs).astype(np.uint16)
        init = rng.integers(0, 128, 32).astype(np.uint8)
        _cabac_case(emul, data, int(rng.integers(0, 9)), ops, init, flags_o, bool(trial & 1))


def test_carry_pipeline_long_carries_and_recopied_chunks(emul):
    """CPU twin of tests/test_gpu_parity.py::test_scan_carry_across_many_chunks: a NAL that loses 70 000 bytes (more than
    the 15-bit per-chunk field holds), one EPB in front of 300 KB of clean payload (verbatim chunks on the re-copy
    list), an EPB in every fourth chunk, a NAL whose every chunk is dirty, and that payload cut into odd NAL units."""
    rng = np.random.default_rng(77)
    sc, hdr = np.array([0, 0, 0, 1], np.uint8), np.array([0x65], np.uint8)
    clean = lambda n: rng.integers(4, 256, n).astype(np.uint8)
    cases = []
    triples = np.tile(np.array([0, 0, 3], np.uint8), 70000)
    cases.append(np.concatenate([sc, hdr, clean(100), triples, clean(50000), sc, hdr, clean(3000), sc]))
    b = clean(300000 + 1234)
    b[10:13] = [0, 0, 3]
    cases.append(np.concatenate([sc, hdr, b, sc, hdr, clean(100), sc]))
    b = clean(64 * 2048 + 321)
    for k in range(0, 64, 4):
        at = k * 2048 + int(rng.integers(8, 2000))
        b[at:at + 3] = [0, 0, 3]
    cases.append(np.concatenate([clean(7), sc, hdr, b, sc]))
    b = clean(120000)
    for p_ in np.sort(rng.choice(np.arange(4, 120000 - 4, 97), 1000, replace=False)):
        b[p_:p_ + 3] = [0, 0, 3]
    cases.append(np.concatenate([sc, hdr, b, sc]))
    parts, pos = [], 0
    while pos < len(b):
        ln = int(rng.integers(50, 9000))
        parts += [sc, hdr, b[pos:pos + ln]]
        pos += ln
    cases.append(np.concatenate(parts + [sc]))
    recopied = 0
    for s in cases:
        s8 = np.ascontiguousarray(s, dtype=np.uint8)
        nal, rbsp = orc.read_nal_units_arrays(s8)
        st = check_carry_pipeline(emul, s8, nal, rbsp)
        recopied += int(st[2])
    assert recopied > 100   # the re-copy list was exercised
