#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

Provenance: the reference (Go) cannot be run here or on the GPU box and ships no vectors of its own,
so these fixtures are outputs of the CPU oracle (oracle/oracle.c, the literal restatement of the reference) on seeded
inputs.  The oracle itself is pinned by the hand-derived known-answer vectors of SURVEY.md Appendix B
(tests/test_oracle_kat.py).  The fixtures freeze that behaviour: tests/test_golden.py checks the oracle against them on
the CPU and the CUDA path against them on the GPU (without calling the oracle).

    python tools/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import harness as hz  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def annexb_fixture():
    """BASELINE configs[0] shape at 64 KiB: SPS + PPS + slices with EPB-bearing payloads, plus extension NAL types."""
    s = hz.build_stream_c1(1 << 16).copy()
    # a few extension-header NALs (types 14 / 20 / 21) so that the header fields are pinned too
    rng = np.random.default_rng(2026)
    extra = []
    for first, second in ((0x6E, 0x80), (0x74, 0x00), (0x75, 0xFF), (0x75, 0x7F)):
        body = rng.integers(4, 256, 40).astype(np.uint8)
        extra.append(np.concatenate([np.array([first, second, 0xA5, 0x5A], np.uint8), body, np.array([0, 0, 3, 1], np.uint8),
                                     np.array([0, 0, 0, 1], np.uint8)]))
    s = np.concatenate([s] + extra)
    nal, rbsp = orc.read_nal_units_arrays(s)
    np.savez_compressed(os.path.join(OUT, "annexb_c1_64k.npz"), stream=s, start=nal["start"], num_bytes=nal["num_bytes"],
                        rbsp_off=nal["rbsp_off"], rbsp_len=nal["rbsp_len"], type=nal["type"], ref_idc=nal["ref_idc"],
                        fzb=nal["fzb"], header_bytes=nal["header_bytes"], epb=nal["epb"], fields=nal["fields"],
                        field_names=np.array(orc._NAL_FIELDS), rbsp=rbsp)
    return len(nal["start"])


def cabac_fixture():
    """BASELINE configs[1] shape, small: 16 slices x up to 3000 ops of one shared schedule; both bypass forms."""
    n_active, n_ctx, n = 64, 64, 16
    ops = hz.gen_schedule(2, 3000, n_active)
    qp, idc = hz.slice_params(n, first=40)
    n_ops = np.array([3000, 2999, 2048, 1, 31, 32, 33, 383, 384, 385, 1500, 2976, 3000, 700, 64, 2000], np.uint32)
    g = hz.gen_cabac_slices(2, ops, n_ops, n_active, n_ctx, qp, idc)
    init = orc.ctx_init(qp, idc, n_ctx)
    stride = g["data"].shape[1]
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    out = dict(ops=ops, n_ops=n_ops, qp=qp, idc=idc, data=g["data"], lens=g["lens"], init_states=init, n_ctx=n_ctx)
    for name, flags in (("spec_or", orc.BYPASS_SPEC_OR), ("ref_shift", 0)):
        words = (len(ops) + 1 + 31) // 32
        bins = np.zeros((n, words), np.uint32)
        fin = np.zeros((n, 4), np.int64)      # codIRange, codIOffset, bitsRead, n_bins
        status = np.zeros(n, np.int64)
        states = np.zeros((n, n_ctx), np.uint8)
        for s in range(n):
            sl = np.concatenate([ops[:n_ops[s]], term])
            rc, b, f, st = orc.cabac_decode_slice(g["data"][s, :g["lens"][s]], sl, init[s], flags)
            bins[s, :len(b)] = b
            fin[s] = (f["codIRange"], f["codIOffset"], f["bitsRead"], f["n_bins"])
            status[s] = rc
            states[s] = st
        out.update({"bins_" + name: bins, "final_" + name: fin, "status_" + name: status, "states_" + name: states})
    assert stride == g["data"].shape[1]
    np.savez_compressed(os.path.join(OUT, "cabac_16x3000.npz"), **out)
    return n


def ctx_init_fixture():
    """BASELINE configs[2] KAT: every ctxIdx 0..1023 x SliceQPY -3..55 x cabac_init_idc {-1,0,1,2,3}, REF and SPEC."""
    qps = np.arange(-3, 56, dtype=np.int32)
    idcs = np.array([-1, 0, 1, 2, 3], np.int32)
    qp = np.repeat(qps, len(idcs))
    idc = np.tile(idcs, len(qps))
    np.savez_compressed(os.path.join(OUT, "ctx_init_sweep.npz"), qp=qp, idc=idc,
                        states_ref=orc.ctx_init(qp, idc, 1024, 0), states_spec=orc.ctx_init(qp, idc, 1024, orc.TABLES_SPEC))
    return len(qp)


def param_set_fixture():
    """Rows S1 / f4 / f1: seeded SPS, PPS and slice-header RBSPs (written + random bytes) with the oracle's NewSPS / NewPPS /
    NewSliceContext results, as structured arrays in the product's field order."""
    from h264decode_b200 import capi   # (dtypes only: needs the built library to import, not a GPU)
    from tests.test_param_sets import sps_cases, pps_cases
    from tests.test_slice_header import cases as header_cases, oracle_header, FIELD_MAP
    sc, pc = sps_cases(11, 40, 40), pps_cases(11, 40, 40)

    def pack(cs):
        off = np.cumsum([0] + [len(c[0]) + 3 for c in cs])[:-1]
        data = np.zeros(int(off[-1]) + len(cs[-1][0]) + 8, np.uint8)
        for o, c in zip(off, cs):
            data[o:o + len(c[0])] = c[0]
        return data, off.astype(np.uint64), np.array([len(c[0]) for c in cs], np.uint32)

    out = {}
    for name, cs, fn, dtype, scal_mine, scal_orc in (("sps", sc, orc.new_sps, capi.SPS_DTYPE, capi.SPS_SCALARS, orc._SPS_SCALARS),
                                                    ("pps", pc, orc.new_pps, capi.PPS_DTYPE, capi.PPS_SCALARS, orc._PPS_SCALARS)):
        data, off, ln = pack(cs)
        rec = np.zeros(len(cs), dtype)
        for i, (rb, _) in enumerate(cs):
            st, f = fn(rb)
            rec[i]["status"] = st
            if st != orc.OK:
                continue
            for a, b in zip(scal_mine, scal_orc):
                rec[i][a] = f[b]
            rec[i]["bits_read"] = f["bits_read"]
            if name == "sps":
                rec[i]["n_seq_scaling_list"] = len(f["SeqScalingList"])
                rec[i]["seq_scaling_list"][:len(f["SeqScalingList"])] = f["SeqScalingList"]
                rec[i]["n_offset_for_ref_frame"], rec[i]["n_hrd"] = f["n_OffsetForRefFrameList"], f["n_hrd"]
                k = min(f["n_OffsetForRefFrameList"], capi.SPS_MAX_REF_FRAMES)
                rec[i]["offset_for_ref_frame"][:k] = f["OffsetForRefFrameList"][:k]
                k = min(f["n_hrd"], capi.SPS_MAX_HRD)
                for a, b in (("bit_rate_value_minus1", "BitRateValueMinus1"), ("cpb_size_value_minus1", "CpbSizeValueMinus1"),
                             ("cbr", "Cbr")):
                    rec[i][a][:k] = f[b][:k]
        out.update({name + "_data": data, name + "_off": off, name + "_len": ln, name: rec})
    hc = header_cases(11, 60, 60)
    data, off, ln = pack([(c[3], None) for c in hc])
    hdr = np.zeros(len(hc), capi.SLICE_HEADER_DTYPE)
    ps = np.zeros((len(hc), len(capi.PARAM_SET_FIELDS)), np.int64)
    for i, (p, nal_type, ref_idc, rb, _) in enumerate(hc):
        rc, h = oracle_header(p, nal_type, ref_idc, rb)
        hdr[i]["status"] = rc
        ps[i] = [p.get(k, 0) for k in capi.PARAM_SET_FIELDS]
        if rc == orc.PANIC:
            continue
        for mine, theirs in FIELD_MAP:
            hdr[i][mine] = h[theirs]
    out.update(hdr_data=data, hdr_off=off, hdr_len=ln, hdr=hdr, hdr_param_sets=ps,
               hdr_nal_type=np.array([c[1] for c in hc], np.uint8), hdr_ref_idc=np.array([c[2] for c in hc], np.uint8))
    np.savez_compressed(os.path.join(OUT, "param_sets_headers.npz"), **out)
    return len(sc) + len(pc) + len(hc)


def glue_fixture():
    """Rows I5 / f3: CtxIdx over every offset the reference knows x binIdx -3..11, every NewBinarization row, every
    mb_type / sub_mb_type bin string."""
    from tests.test_ctx_glue import OFFSETS, BIN_IDX
    b, o = np.meshgrid(np.array(BIN_IDX, np.int64), np.array(OFFSETS, np.int64))
    b, o = b.ravel(), o.ravel()
    ctx_idx = np.array([orc.ctx_idx(int(x), 7, int(y)) for x, y in zip(b, o)], np.int64)
    se, st = np.meshgrid(np.arange(-1, 16, dtype=np.int32), np.arange(-1, 7, dtype=np.int32))
    se, st = se.ravel(), st.ravel()
    bz = np.array([[orc.new_binarization(int(x), int(y))[k] for k in orc.BINARIZATION_FIELDS] for x, y in zip(se, st)], np.int32)
    t_st, t_sub, t = np.meshgrid(np.arange(-1, 7, dtype=np.int32), np.arange(2, dtype=np.uint8), np.arange(-2, 34, dtype=np.int64))
    t_st, t_sub, t = t_st.ravel(), t_sub.ravel(), t.ravel()
    strs = [orc.mb_bin_string(int(a), int(c), int(s)) for a, s, c in zip(t_st, t_sub, t)]
    np.savez_compressed(os.path.join(OUT, "ctx_glue.npz"), bin_idx=b, offset=o, ctx_idx=ctx_idx, se=se, st=st, binarization=bz,
                        mb_st=t_st, mb_sub=t_sub, mb_type=t, mb_len=np.array([len(x) for x in strs], np.int32),
                        mb_bits=np.array([sum(v << k for k, v in enumerate(x)) for x in strs], np.uint32))
    return len(b) + len(se) + len(t)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    print("annexb:", annexb_fixture(), "NAL units")
    print("cabac:", cabac_fixture(), "slices")
    print("ctx_init:", ctx_init_fixture(), "(qp, idc) pairs")
    print("param sets + slice headers:", param_set_fixture(), "RBSPs")
    print("glue:", glue_fixture(), "queries")
    for f in sorted(os.listdir(OUT)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))
