"""The oracle (oracle/oracle.c) against an independent second model of the reference (tests/ref_model.py, written from
the Go files with its own table reader): SURVEY.md section 8 rows N1-N6, E1-E7, I1-I4.  CPU only.

The reference ships no golden vectors and cannot be built (no Go toolchain, un-vendored dependency), so parity is
pinned by two restatements written separately from the same Go source that must agree on every input tried here,
next to the hand-derived known answers of tests/test_oracle_kat.py."""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as orc
from tests import ref_model as rm


def test_snapshot_matches_the_go_files():
    snap = rm._from_json(json.load(open(rm.SNAPSHOT)))
    assert len(snap["range_tab_lps"]) == 64 and len(snap["trans"]) == 64
    assert len(snap["mn_vars"]) == 40 and len(snap["cbp"]) == 35
    if os.path.isdir(os.path.join(rm.REF_ROOT, "h264")):
        assert rm.read_go_tables() == snap, "tests/golden/go_tables.json is stale: run tools/make_go_tables.py"


# ------------------------------------------------------------------------------------------------ E7 / E2 / E3 tables
@pytest.mark.parametrize("spec", [False, True])
def test_engine_tables_and_primitives_exhaustive(spec):
    flags = orc.TABLES_SPEC if spec else 0
    range_tab, trans = rm.spec_tables() if spec else (rm.tables()["range_tab_lps"], rm.tables()["trans"])
    for p in range(64):
        for v in (0, 1):
            for b in (0, 1):
                assert orc.state_transition(p, v, b, flags) == rm.state_transition(p, v, b, trans)
            # BinaryDecision core over every qCodIRangeIdx and both outcomes, plus offsets far outside [0, R)
            for R in (256, 300, 320, 383, 384, 400, 447, 448, 510, 2, 64, 1000, -5):
                for O in (0, 1, 100, 255, 256, 400, 509, 510, 1 << 40, -(1 << 62), -1):
                    got = orc.binary_decision(p, v, R, O, flags)
                    assert got == rm.binary_decision(p, v, R, O, range_tab), (p, v, R, O)


# ------------------------------------------------------------------------------------------------ I1 - I4
def test_mn_every_ctx_and_idc():
    for ctx in list(range(0, 1024)) + [-1, 1024, 5000]:
        for idc in range(-3, 6):
            want = rm.mn_lookup(ctx, idc) if 0 <= ctx < 1024 else (0, 0)
            assert orc.mn(ctx, idc) == want, (ctx, idc)


def test_pre_ctx_state_and_split_exhaustive():
    ms = sorted({m for row in rm.tables()["mn_vars"].values() for m, _ in row.values()} |
                {m for cols, ret in rm.tables()["cbp"].values() for m, _ in cols + [ret]} | {-128, 127, 0})
    ns = sorted({n for row in rm.tables()["mn_vars"].values() for _, n in row.values()} | {-128, 127, 0, 63, 64})
    for m in ms:
        for n in ns:
            for qp in (-7, 0, 1, 17, 26, 50, 51, 52, 80):
                pre = rm.pre_ctx_state(m, n, qp)
                assert orc.pre_ctx_state(m, n, qp) == pre
    for pre in range(1, 127):
        p, v = rm.state_split(pre)
        assert orc.ctx_state(pre) == (p | (v << 6))


def test_ctx_init_rows_match_the_model():
    qps = np.array([q for q in range(-3, 56) for _ in range(7)], np.int32)
    idcs = np.array([i for _ in range(-3, 56) for i in (-2, -1, 0, 1, 2, 3, 7)], np.int32)
    got = orc.ctx_init(qps, idcs, 1024)
    for row, (qp, idc) in enumerate(zip(qps, idcs)):
        want = [rm.ctx_state_byte(c, int(idc), int(qp)) for c in range(1024)]
        assert got[row].tolist() == want, (qp, idc)


# ------------------------------------------------------------------------------------------------ N1 - N6
def _random_stream(rng, n):
    """Bytes rich in the patterns that matter: start codes, 00 00 03, runs of zeros, extension NAL types."""
    out = bytearray()
    while len(out) < n:
        r = rng.random()
        if r < 0.10:
            out += b"\x00\x00\x00\x01"
            if rng.random() < 0.5:  # a header byte, sometimes of an extension type, sometimes followed by zeros
                out.append(rng.choice([0x65, 0x41, 0x67, 0x68, 14, 20, 21, 0x6E, 0x74, 0x75, 0x00, 0x03]))
        elif r < 0.25:
            out += rng.choice([b"\x00\x00\x03", b"\x00\x00\x03\x00\x00\x03", b"\x00\x00\x00\x03", b"\x00\x00",
                               b"\x00\x00\x01", b"\x00\x00\x03\x01", b"\x00\x03", b"\x00\x00\x00\x00\x01"])
        elif r < 0.45:
            out.append(rng.choice([0, 0, 0, 1, 2, 3]))
        else:
            out += bytes(rng.randrange(256) for _ in range(rng.randrange(1, 12)))
    return bytes(out[:n])


def _compare_stream(stream):
    want = rm.read_nal_units(stream)
    for literal in (False, True):
        nal, rbsp = orc.read_nal_units_arrays(stream, literal=literal)
        assert len(nal["start"]) == len(want), (len(nal["start"]), len(want))
        for k, (start, end, fields, body) in enumerate(want):
            assert nal["start"][k] == start and nal["end"][k] == end, k
            got_fields = dict(zip(orc._NAL_FIELDS, nal["fields"][k].tolist()))
            for name in rm.NAL_FIELDS:
                assert got_fields[name] == fields[name], (k, name, got_fields[name], fields[name])
            got_rbsp = bytes(rbsp[nal["rbsp_off"][k]:nal["rbsp_off"][k] + nal["rbsp_len"][k]])
            assert got_rbsp == body, (k, got_rbsp.hex(), body.hex())


def test_split_and_strip_fuzz():
    rng = random.Random(0x4E414C)
    for it in range(1500):
        _compare_stream(_random_stream(rng, rng.randrange(0, 160)))
    for it in range(40):
        _compare_stream(_random_stream(rng, rng.randrange(1000, 6000)))


def test_split_edge_cases():
    for s in (b"", b"\x00", b"\x00\x00\x00\x01", b"\x00\x00\x00\x01\x00\x00\x00\x01",
              b"\x00\x00\x00\x01\x65\x00\x00\x00\x01", b"\x00\x00\x00\x00\x01\x67\x00\x00\x03\x00\x00\x00\x01",
              b"\xff\x00\x00\x00\x01\x14\x80\x00\x00\x03\x00\x00\x03\x00\x00\x00\x01\x41\x00\x00\x00\x01",
              b"\x00\x00\x00\x01" + b"\x00\x00\x03" * 9 + b"\x00\x00\x00\x01",
              b"\x00\x00\x00\x01\x00\x00\x03\x00\x00\x00\x01\x01\x00\x00\x00\x01"):
        _compare_stream(s)


def test_new_nal_unit_direct_calls_fuzz():
    rng = random.Random(0x4E55)
    for it in range(3000):
        frame = _random_stream(rng, rng.randrange(5, 64)).lstrip(b"") or b"\x65\x00\x00\x00\x00"
        # the header bits must exist: the reference panics on frames shorter than the header it announces
        for num in {len(frame), max(1, len(frame) - rng.randrange(0, 4)), len(frame) + 2}:
            try:
                fields, body = rm.new_nal_unit(frame, num)
            except rm.GoPanic:
                st, _, _ = orc.new_nal_unit(frame, num)
                assert st != orc.OK
                continue
            st, got, rbsp = orc.new_nal_unit(frame, num)
            assert st == orc.OK
            for name in rm.NAL_FIELDS:
                assert got[name] == fields[name], (frame.hex(), num, name)
            assert rbsp == body, (frame.hex(), num)


# ------------------------------------------------------------------------------------------------ E1 - E6
def _random_ops(rng, n, n_ctx):
    ops = []
    for _ in range(n):
        r = rng.random()
        if r < 0.65:
            ops.append((rm.OP_DECISION, rng.randrange(n_ctx + 2)))
        elif r < 0.95:
            ops.append((rm.OP_BYPASS, 0))
        else:
            ops.append((rm.OP_TERMINATE, 0))
    return ops


@pytest.mark.parametrize("spec_or", [False, True])
@pytest.mark.parametrize("spec_tables", [False, True])
def test_composed_engine_fuzz(spec_or, spec_tables):
    rng = random.Random(0xCABAC + 2 * spec_or + spec_tables)
    flags = (orc.BYPASS_SPEC_OR if spec_or else 0) | (orc.TABLES_SPEC if spec_tables else 0)
    for it in range(250):
        n_ctx = rng.choice([1, 3, 16, 64])
        n_ops = rng.randrange(0, 400)
        # random bytes (non-conformant streams: offsets beyond the range, terminate bins of 1 in mid-stream) and, every
        # other time, too few bytes for the ops, so that the run-off-the-end convention is compared as well
        data = bytes(rng.randrange(256) for _ in range(rng.randrange(2, 12) if it % 2 else rng.randrange(40, 400)))
        ops = _random_ops(rng, n_ops, n_ctx)
        init = [rng.randrange(128) for _ in range(n_ctx)]
        states = list(init)
        bins, R, O, bits_read, panicked = rm.decode_slice(data, ops, states, spec_or=spec_or, spec_tables_=spec_tables)
        op_words = np.array([orc.make_op(k, c) for k, c in ops], np.uint16)
        rc, obins, fin, ostates = orc.cabac_decode_slice(data, op_words, np.array(init, np.uint8), flags)
        assert (rc != orc.OK) == panicked
        assert fin["n_bins"] == len(bins)
        got_bins = [(int(obins[i >> 5]) >> (i & 31)) & 1 for i in range(len(bins))]
        assert got_bins == bins
        if len(data) * 8 >= 9:
            assert (fin["codIRange"], fin["codIOffset"], fin["bitsRead"]) == (R, O, bits_read), it
        assert ostates.tolist() == states


def test_primitives_on_their_own():
    rng = random.Random(7)
    for it in range(2000):
        data = bytes(rng.randrange(256) for _ in range(4))
        # (codIRange <= 0 doubles for ever and runs off the data: the composed fuzz covers that through the panic rule)
        R = rng.choice([3, 4, 64, 100, 255, 256, 300, 510, rng.randrange(3, 600)])
        O = rng.choice([0, 1, rng.randrange(0, 600), 1 << 62, (1 << 63) - 1, -3])
        for spec_or in (False, True):
            o, b = rm.decode_bypass(rm.BitReader(data), R, O, spec_or)
            assert orc.decode_bypass(data, R, O, orc.BYPASS_SPEC_OR if spec_or else 0) == (o, b)
        r, o, b = rm.decode_terminate(rm.BitReader(data), R, O)
        assert orc.decode_terminate(data, R, O) == (r, o, b)
        br = rm.BitReader(data)
        r, o = rm.renorm_d(br, R, O)
        assert orc.renorm_d(data, R, O) == (r, o, br.bitsRead)
        assert orc.init_decoding_engine(data) == rm.init_decoding_engine(rm.BitReader(data))
