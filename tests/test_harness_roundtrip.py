"""Harness <-> oracle self-checks on CPU: the 9.3.4 test encoder round-trips through the oracle's decoder
(REF tables + SPEC_OR bypass), and escaped Annex-B streams strip back to their payloads."""
import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc


@pytest.mark.parametrize("flags_h,flags_o", [(0, orc.BYPASS_SPEC_OR), (hz.TABLES_SPEC, orc.BYPASS_SPEC_OR | orc.TABLES_SPEC)])
@pytest.mark.parametrize("n_active,n_ctx", [(64, 64), (460, 512)])
def test_encoder_decoder_roundtrip(flags_h, flags_o, n_active, n_ctx):
    n_slices, n_bins = 6, 5000
    ops = hz.gen_schedule(2, n_bins, n_active)
    n_ops = np.array([n_bins, n_bins - 1, 4000, 383, 384, 1], dtype=np.uint32)
    qp, idc = hz.slice_params(n_slices, first=50)
    g = hz.gen_cabac_slices(2, ops, n_ops, n_active, n_ctx, qp, idc, flags=flags_h)
    init = orc.ctx_init(qp, idc, n_ctx, flags_o & orc.TABLES_SPEC)
    for s in range(n_slices):
        assert np.array_equal(init[s], hz.init_states(qp[s], idc[s], n_ctx, flags_h))
        sl_ops = np.concatenate([ops[:n_ops[s]], np.array([orc.make_op(orc.OP_TERMINATE)], dtype=np.uint16)])
        data = g["data"][s, :g["lens"][s]]
        rc, bins, fin, st = orc.cabac_decode_slice(data, sl_ops, init[s], flags_o)
        assert rc == orc.OK
        nw = (len(sl_ops) + 31) // 32
        exp = g["bins"][s, :nw].copy()
        got = np.zeros(nw, dtype=np.uint32)
        got[:len(bins)] = bins
        assert np.array_equal(got, exp), "slice %d bins differ" % s
        assert np.array_equal(st, g["final_states"][s])
        assert fin["n_bins"] == len(sl_ops) and fin["bitsRead"] <= 8 * len(data)
        assert (bins[n_ops[s] >> 5] >> (n_ops[s] & 31)) & 1 == 1   # the final terminate bin


def test_bits_per_bin_is_plausible():
    ops = hz.gen_schedule(2, 20000, 64)
    qp, idc = hz.slice_params(4)
    g = hz.gen_cabac_slices(2, ops, np.full(4, 20000, np.uint32), 64, 64, qp, idc)
    bpb = g["lens"].mean() * 8 / 20000
    assert 0.6 < bpb < 1.0


def test_escape_then_strip_is_identity():
    for k in range(8):
        pay = hz.random_payload(1, k, 3000 + k)
        esc = hz.escape(pay)
        assert not any(esc[i] == 0 and esc[i + 1] == 0 and esc[i + 2] <= 2 for i in range(len(esc) - 2))
        stream = np.concatenate([np.frombuffer(hz.SC + b"\x65", np.uint8), esc, np.frombuffer(hz.SC, np.uint8)])
        cnt, recs, rbsp = orc.read_nal_units(stream)
        assert cnt == 1
        body = rbsp[recs[0].rbsp_off:recs[0].rbsp_off + recs[0].nal.rbsp_len]
        # reference RBSP = payload followed by the first two bytes of the next start code (A7/A8)
        # (a payload ending in a single 00 gets a trailing 03 that is not an EPB -- only one zero before it)
        keep03 = len(pay) >= 1 and pay[-1] == 0 and not (len(pay) >= 2 and pay[-2] == 0)
        exp = np.concatenate([pay, np.array(([3] if keep03 else []) + [0, 0], dtype=np.uint8)])
        assert np.array_equal(body, exp)


def test_c1_stream_shape():
    s = hz.build_stream_c1(1 << 16)
    assert abs(len(s) - (1 << 16)) < 64
    nal, rbsp = orc.read_nal_units_arrays(s)
    assert nal["type"][0] == 7 and nal["type"][1] == 8 and set(nal["type"][2:]) <= {5, 1}
    assert nal["end"][-1] == len(s)
    epb = (nal["num_bytes"] - nal["header_bytes"] - 2 - nal["rbsp_len"]).sum()
    assert epb > 0.002 * len(s)


def test_cabac_stream_feeds_decoder_after_strip():
    b = hz.build_stream_cabac(12, 3000, n_active=64, n_ctx=64, slices_per_frame=4, frames_per_params=2)
    nal, rbsp = orc.read_nal_units_arrays(b["stream"])
    sl = np.flatnonzero((nal["type"] == 1) | (nal["type"] == 5))
    assert len(sl) == 12
    init = orc.ctx_init(b["qp"], b["idc"], b["n_ctx"])
    for i, k in enumerate(sl):
        data = rbsp[nal["rbsp_off"][k]:nal["rbsp_off"][k] + nal["rbsp_len"][k]]
        sl_ops = np.concatenate([b["ops"][:b["n_ops"][i]], np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)])
        rc, bins, fin, st = orc.cabac_decode_slice(data, sl_ops, init[i], orc.BYPASS_SPEC_OR)
        assert rc == orc.OK
        nw = (len(sl_ops) + 31) // 32
        got = np.zeros(nw, np.uint32)
        got[:len(bins)] = bins
        assert np.array_equal(got, b["bins"][i, :nw])
        assert np.array_equal(st, b["final_states"][i])
