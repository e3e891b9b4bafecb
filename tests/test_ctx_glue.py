"""Syntax-element glue (rows I5 / f3: CtxIdx h264/cabac.go:557-758, NewBinarization :340-427, initCabac :148-174,
binIdxMbMap / binIdxSubMbMap :180-303, IsBinStringMatch :429-436): the product's table-driven functions (ctx_glue.cuh)
against the oracle's literal switch restatements -- exhaustively over every offset / name / type the reference knows
plus out-of-range values, on the CPU through the emulation library and on the GPU through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc

OFFSETS = [0, 3, 11, 14, 17, 21, 24, 27, 32, 36, 40, 47, 54, 60, 64, 68, 69, 70, 73, 77, 276, 399, 1, 5, 100, -3, 10000]
BIN_IDX = list(range(-3, 12)) + [100, -1000, 1 << 40]


@pytest.fixture(scope="module")
def emul():
    from tests.test_param_sets import emul as _  # noqa: F401  (same build recipe)
    from tests import test_hd_logic
    import os
    import subprocess
    deps = [test_hd_logic.SRC] + [os.path.join(test_hd_logic.HERE, "..", "h264decode_b200", "csrc", f)
                                  for f in ("ctx_glue.cuh", "param_sets.cuh", "slice_header.cuh", "annexb_local.cuh",
                                            "cabac_lane.cuh", "tables.inc")] + [
        os.path.join(test_hd_logic.HERE, "..", "include", "h264b200.h")]
    out = test_hd_logic.OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-x", "c++", test_hd_logic.SRC,
                               "-o", out])
    L = C.CDLL(out)
    L.emul_ctx_idx.restype = C.c_int64
    L.emul_ctx_idx.argtypes = [C.c_int64] * 3
    L.emul_new_binarization.restype = None
    L.emul_new_binarization.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
    L.emul_mb_bin_string.restype = None
    L.emul_mb_bin_string.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    L.emul_bin_string_match.restype = C.c_int32
    L.emul_bin_string_match.argtypes = [C.c_int32, C.c_uint32, C.c_int32, C.c_uint32]
    return L


def unpack(n, bits):
    return [(int(bits) >> k) & 1 for k in range(int(n))]


def test_ctx_idx_known_answers():
    """hand-read from cabac.go:557-758"""
    assert orc.ctx_idx(1, 6, 3) == 276 and orc.ctx_idx(0, 6, 3) == orc.NA_CTX_ID and orc.ctx_idx(9, 6, 3) == 7
    assert orc.ctx_idx(1, 0, 69) == orc.NA_CTX_ID          # offset 69 never answers (:714-718)
    assert orc.ctx_idx(-2, 0, 21) == -2                    # offset 21 hands a negative binIdx back
    assert orc.ctx_idx(5, 0, 17) == 3 and orc.ctx_idx(4, 0, 17) == orc.NA_CTX_ID
    assert orc.ctx_idx(0, 0, 68) == 0 and orc.ctx_idx(0, 0, 276) == 0
    # every reachable initCabac lands on (62, 0): binIdx is never set, MNVars[ctxIdx] has no key 0 for ctxIdx <= 10
    for se in range(15):
        for st in range(6):
            b = orc.new_binarization(se, st)
            assert orc.init_cabac(0, b["max_prefix"], b["off_prefix"], -3, 5)[:2] == (62, 0)


def test_ctx_idx_matches_oracle_cpu(emul):
    for off in OFFSETS:
        for b in BIN_IDX:
            assert emul.emul_ctx_idx(b, 7, off) == orc.ctx_idx(b, 7, off), (b, off)


def test_binarization_and_bin_strings_match_oracle_cpu(emul):
    from h264decode_b200 import capi
    for se in range(-1, 16):
        for st in range(-1, 7):
            out = np.zeros(1, capi.BINARIZATION_DTYPE)
            emul.emul_new_binarization(se, st, out.ctypes.data)
            exp = orc.new_binarization(se, st)
            assert {k: int(out[0][k]) for k in exp} == exp, (se, st)
    strings = []
    for st in range(-1, 7):
        for sub in (0, 1):
            for t in list(range(-2, 34)) + [1 << 35]:
                ln, bits = C.c_int32(), C.c_uint32()
                emul.emul_mb_bin_string(st, t, sub, C.byref(ln), C.byref(bits))
                exp = orc.mb_bin_string(st, t, sub)
                assert unpack(ln.value, bits.value) == exp, (st, sub, t)
                strings.append(exp)
    assert orc.mb_bin_string(2, 8) == [1, 0, 0, 1, 0, 1, 1] and orc.mb_bin_string(0, 1) == [0, 1, 1]
    rng = np.random.default_rng(0)
    seen = set()
    for s in strings[::3]:
        for _ in range(6):
            n = int(rng.integers(0, 9))
            bits = [int(x) for x in rng.integers(0, 2, n)] if rng.random() < 0.5 else (s + [1, 0, 1])[:n]
            pack = lambda v: sum(b << k for k, b in enumerate(v))  # noqa: E731
            got = emul.emul_bin_string_match(len(s), pack(s), n, pack(bits))
            assert got == orc.bin_string_match(s, bits), (s, bits)
            seen.add(got)
    assert seen == {0, 1, 2}


@pytest.mark.gpu
def test_glue_gpu_matches_oracle():
    from h264decode_b200 import capi
    ctx = capi.Context(0)
    try:
        b, o = np.meshgrid(np.array(BIN_IDX, np.int64), np.array(OFFSETS, np.int64))
        b, o = b.ravel(), o.ravel()
        got = ctx.ctx_idx(b, np.full(len(b), 7), o)
        assert list(got) == [orc.ctx_idx(int(x), 7, int(y)) for x, y in zip(b, o)]
        se, st = np.meshgrid(np.arange(-1, 16, dtype=np.int32), np.arange(-1, 7, dtype=np.int32))
        se, st = se.ravel(), st.ravel()
        bz = ctx.new_binarization(se, st)
        for i in range(len(se)):
            exp = orc.new_binarization(int(se[i]), int(st[i]))
            assert {k: int(bz[i][k]) for k in exp} == exp, i
        # initCabac over every (binIdx, binarization) and a qp sweep, both table sets
        rng = np.random.default_rng(1)
        n = 4000
        bi = rng.choice(np.array(BIN_IDX[:15], np.int64), n)
        offp = rng.choice(np.array(OFFSETS + [21, 21, 21], np.int64), n)
        piq, sqd = rng.integers(-40, 40, n), rng.integers(-40, 40, n)
        for flags in (0, capi.TABLES_SPEC):
            p, v, c = ctx.init_cabac(bi, np.zeros(n, np.int64), offp, piq, sqd, flags=flags)
            for i in range(n):
                assert (int(p[i]), int(v[i]), int(c[i])) == orc.init_cabac(int(bi[i]), 0, int(offp[i]), int(piq[i]),
                                                                          int(sqd[i]), flags), i
        # CtxIdx never answers 11..39, the only MNVars rows with a cabac_init_idc 0 entry: (62, 0) always
        assert {(int(a), int(b)) for a, b in zip(p, v)} == {(62, 0)}
        st2, sub, t = np.meshgrid(np.arange(-1, 7, dtype=np.int32), np.arange(2, dtype=np.uint8), np.arange(-2, 34, dtype=np.int64))
        st2, sub, t = st2.ravel(), sub.ravel(), t.ravel()
        ln, bits = ctx.mb_bin_string(st2, t, sub)
        for i in range(len(t)):
            assert unpack(ln[i], bits[i]) == orc.mb_bin_string(int(st2[i]), int(t[i]), int(sub[i])), i
        m = ctx.bin_string_match(ln, bits, np.roll(ln, 7), np.roll(bits, 7))
        for i in range(len(t)):
            assert int(m[i]) == orc.bin_string_match(unpack(ln[i], bits[i]), unpack(np.roll(ln, 7)[i], np.roll(bits, 7)[i])), i
        assert len(ctx.ctx_idx([], [], [])) == 0
    finally:
        ctx.close()
