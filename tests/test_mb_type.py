"""Row f3 of SURVEY.md section 8: mb_type decoded as a syntax element -- every lane on its own, data-dependent sequence of
(context, decision | terminate) ops -- through h264b_mb_type_decode, against the oracle's restatement of the reference's
binarisation walk (h264/slice.go:639-672, h264/cabac.go:180-303, :340-436, :557-758) and against what the test encoder
coded.  The (context, bin) sequences the encoder is fed are derived here, in the test, from the bin-string tables and
the ctxIdx rules, independently of both implementations."""
import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc

T = None  # DecodeTerminate


def i_incs(bits, prev):
    """ctxIdxInc of every bin of an I-slice bin string (Table 9-39 as CtxIdx has it, cabac.go:563-580, + 9.3.3.1.1.3 /
    9.3.3.1.2 where it leaves a comment): offset 3"""
    b3 = bits[3] if len(bits) > 3 else 0
    return [prev, T, 3, 4, 5 if b3 else 6, 6 if b3 else 7, 7][:len(bits)]


def suffix_incs(bits):
    b3 = bits[3] if len(bits) > 3 else 0
    return [0, T, 1, 2, 2 if b3 else 3, 3, 3][:len(bits)]   # offset 17, cabac.go:599-613


def ops_for(kind, types):
    """(op words, bins) that code the mb_type sequence of one slice"""
    ops, bins, prev = [], [], 0
    for t in types:
        if kind == 0:
            bits = orc.mb_bin_string(2, t)
            seq = [(3 + i if i is not T else T) for i in i_incs(bits, prev)]
            prev = 1 if t != 0 else 0
        elif t < 4:
            bits = orc.mb_bin_string(0, t)
            seq = [14, 15, 16 if bits[1] != 1 else 17]
        else:
            sb = orc.mb_bin_string(2, t - 5)
            bits = [1] + sb
            seq = [14] + [(17 + i if i is not T else T) for i in suffix_incs(sb)]
        for c, b in zip(seq, bits):
            ops.append(orc.make_op(orc.OP_TERMINATE) if c is T else orc.make_op(orc.OP_DECISION, c))
            bins.append(b)
    return np.array(ops, np.uint16), np.array(bins, np.uint8)


def random_types(rng, kind, n, allow_pcm):
    if kind == 0:
        t = rng.integers(0, 25, n)
        t[rng.random(n) < 0.3] = 0            # I_NxN often, so that the first bin's context moves
    else:
        t = rng.integers(0, 29, n)
        t[t == 4] = 1                         # (mb_type 4 has the empty bin string in the reference's table)
        t[rng.random(n) < 0.5] = rng.integers(0, 4)
    if allow_pcm and n > 3:
        t[-1] = 25 if kind == 0 else 30       # I_PCM ends the walk
    return t


def build_slices(rng, n_slices, n_ctx, tables_spec=False, pcm_every=0):
    kinds = rng.integers(0, 2, n_slices).astype(np.uint8)
    qp, idc = hz.slice_params(n_slices, first=3)
    init = orc.ctx_init(qp, idc, n_ctx, orc.TABLES_SPEC if tables_spec else 0)
    datas, all_types, enc_states = [], [], []
    for s in range(n_slices):
        n = int(rng.integers(1, 400))
        t = random_types(rng, int(kinds[s]), n, pcm_every and s % pcm_every == 0)
        ops, bins = ops_for(int(kinds[s]), t)
        d, st = hz.encode_explicit(ops, bins, init[s], flags=hz.TABLES_SPEC if tables_spec else 0)
        datas.append(d)
        all_types.append(t)
        enc_states.append(st)
    return kinds, qp, idc, init, datas, all_types, enc_states


def pack(datas, rng):
    off, parts, pos = [], [], 0
    for d in datas:
        gap = int(rng.integers(0, 5))
        parts.append(np.full(gap, 0xA5, np.uint8))
        pos += gap
        off.append(pos)
        parts.append(d)
        pos += len(d)
    parts.append(np.zeros(16, np.uint8))
    return np.concatenate(parts), np.array(off, np.uint64), np.array([len(d) for d in datas], np.uint32)


@pytest.mark.parametrize("tables_spec", [False, True])
def test_oracle_walk_returns_what_the_encoder_coded(tables_spec):
    rng = np.random.default_rng(11 + tables_spec)
    kinds, qp, idc, init, datas, types, enc_states = build_slices(rng, 60, 32, tables_spec, pcm_every=7)
    for s in range(60):
        rc, got, fin, st = orc.decode_mb_types(datas[s], int(kinds[s]), len(types[s]), init[s],
                                               orc.TABLES_SPEC if tables_spec else 0)
        assert rc == orc.OK and np.array_equal(got, types[s]), s
        assert np.array_equal(st, enc_states[s]), s


@pytest.mark.gpu
@pytest.mark.parametrize("tables_spec", [False, True])
def test_gpu_mb_type_walk_matches_oracle_and_encoder(tables_spec):
    from h264decode_b200 import capi
    rng = np.random.default_rng(23 + tables_spec)
    n_slices, n_ctx = 300, 32
    kinds, qp, idc, init, datas, types, _ = build_slices(rng, n_slices, n_ctx, tables_spec, pcm_every=9)
    data, off, length = pack(datas, rng)
    n_mb = np.array([len(t) for t in types], np.uint32)
    n_mb[5] += 40     # asks for more elements than the slice holds: the walk runs on into whatever follows, like the oracle
    ctx = capi.Context(0)
    try:
        fl = capi.TABLES_SPEC if tables_spec else 0
        out, fin, fst = ctx.mb_type_decode(data, off, length, kinds, n_mb, n_ctx, qp=qp, idc=idc, flags=fl)
        out2, fin2, fst2 = ctx.mb_type_decode(data, off, length, kinds, n_mb, n_ctx, init_states=init, flags=fl)
    finally:
        ctx.close()
    assert np.array_equal(out, out2) and np.array_equal(fin, fin2) and np.array_equal(fst, fst2)
    for s in range(n_slices):
        d = data[int(off[s]):int(off[s]) + int(length[s])]
        rc, want, ofin, ost = orc.decode_mb_types(d, int(kinds[s]), int(n_mb[s]), init[s], orc.TABLES_SPEC if tables_spec else 0)
        assert fin["n_mb"][s] == len(want) and np.array_equal(out[s, :len(want)], want), s
        assert bool(fin["flags"][s] & capi.F_OVERRUN) == bool(ofin["flags"]), s
        assert (fin["n_bins"][s], fin["cod_i_range"][s], fin["cod_i_offset"][s], fin["bits_read"][s]) == (
            ofin["n_bins"], ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"]), s
        assert np.array_equal(fst[s], ost), s
        if s != 5:
            assert np.array_equal(want, types[s]), s


@pytest.mark.gpu
def test_gpu_mb_type_walk_on_arbitrary_bytes():
    """not CABAC data at all: both walks must stop at the same bin for the same reason"""
    from h264decode_b200 import capi
    rng = np.random.default_rng(5)
    n_slices, n_ctx = 200, 21
    datas = [rng.integers(0, 256, int(rng.integers(2, 200))).astype(np.uint8) for _ in range(n_slices)]
    data, off, length = pack(datas, rng)
    kinds = rng.integers(0, 2, n_slices).astype(np.uint8)
    init = rng.integers(0, 128, (n_slices, n_ctx)).astype(np.uint8)
    init[:, :] &= ~np.uint8(0)   # any 7-bit state, pStateIdx 63 included
    n_mb = rng.integers(1, 300, n_slices).astype(np.uint32)
    ctx = capi.Context(0)
    try:
        out, fin, fst = ctx.mb_type_decode(data, off, length, kinds, n_mb, n_ctx, init_states=init)
    finally:
        ctx.close()
    for s in range(n_slices):
        d = data[int(off[s]):int(off[s]) + int(length[s])]
        rc, want, ofin, ost = orc.decode_mb_types(d, int(kinds[s]), int(n_mb[s]), init[s])
        assert fin["n_mb"][s] == len(want) and np.array_equal(out[s, :len(want)], want), s
        assert bool(fin["flags"][s] & capi.F_OVERRUN) == bool(ofin["flags"]), s
        assert (fin["n_bins"][s], fin["bits_read"][s]) == (ofin["n_bins"], ofin["bitsRead"]), s
        assert (fin["cod_i_range"][s], fin["cod_i_offset"][s]) == (ofin["codIRange"], ofin["codIOffset"]), s
        assert np.array_equal(fst[s], ost), s
