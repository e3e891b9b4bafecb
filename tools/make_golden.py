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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    print("annexb:", annexb_fixture(), "NAL units")
    print("cabac:", cabac_fixture(), "slices")
    print("ctx_init:", ctx_init_fixture(), "(qp, idc) pairs")
    for f in sorted(os.listdir(OUT)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))
