"""Committed golden fixtures (tests/golden/, written by tools/make_golden.py): the oracle against them on the CPU,
the CUDA path (through the C ABI, without the oracle) against them on the GPU."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_reproduces_annexb_fixture():
    from oracle import oracle as orc
    g = load("annexb_c1_64k.npz")
    for literal in (False, True):
        nal, rbsp = orc.read_nal_units_arrays(g["stream"], literal=literal)
        for k in ("start", "num_bytes", "rbsp_off", "rbsp_len", "type", "ref_idc", "fzb", "header_bytes", "epb", "fields"):
            assert np.array_equal(nal[k], g[k]), k
        assert np.array_equal(rbsp, g["rbsp"])
    assert list(g["field_names"]) == list(orc._NAL_FIELDS)


def test_oracle_reproduces_cabac_fixture():
    from oracle import oracle as orc
    g = load("cabac_16x3000.npz")
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    for name, flags in (("spec_or", orc.BYPASS_SPEC_OR), ("ref_shift", 0)):
        for s in range(len(g["n_ops"])):
            sl = np.concatenate([g["ops"][:g["n_ops"][s]], term])
            rc, b, f, st = orc.cabac_decode_slice(g["data"][s, :g["lens"][s]], sl, g["init_states"][s], flags)
            assert rc == g["status_" + name][s]
            assert np.array_equal(b, g["bins_" + name][s, :len(b)])
            assert (f["codIRange"], f["codIOffset"], f["bitsRead"], f["n_bins"]) == tuple(g["final_" + name][s])
            assert np.array_equal(st, g["states_" + name][s])


def test_oracle_reproduces_ctx_init_fixture():
    from oracle import oracle as orc
    g = load("ctx_init_sweep.npz")
    assert np.array_equal(orc.ctx_init(g["qp"], g["idc"], 1024, 0), g["states_ref"])
    assert np.array_equal(orc.ctx_init(g["qp"], g["idc"], 1024, orc.TABLES_SPEC), g["states_spec"])
    assert not np.array_equal(g["states_ref"], g["states_spec"])  # the reference's table typos are visible


# ------------------------------------------------------------------------------------------------ GPU: the product
@pytest.fixture(scope="module")
def ctx():
    from h264decode_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_annexb_matches_fixture(ctx):
    from h264decode_b200 import capi
    g = load("annexb_c1_64k.npz")
    summ, nals, ext, rbsp = ctx.annexb_scan(g["stream"])
    assert summ["n_nals"] == len(g["start"]) and summ["rbsp_bytes"] == len(g["rbsp"])
    assert np.array_equal(nals["start"].astype(np.int64), g["start"])
    assert np.array_equal(nals["num_bytes"].astype(np.int64), g["num_bytes"])
    assert np.array_equal(nals["rbsp_len"].astype(np.int64), g["rbsp_len"])
    assert np.array_equal(nals["type"].astype(np.int64), g["type"])
    assert np.array_equal(nals["ref_idc"].astype(np.int64), g["ref_idc"])
    assert np.array_equal(nals["header_bytes"].astype(np.int64), g["header_bytes"])
    assert np.array_equal((nals["flags"] & capi.F_HAS_EPB) != 0, g["epb"] == 3)
    dense = np.concatenate([rbsp[int(o):int(o) + int(n)] for o, n in zip(nals["rbsp_off"], nals["rbsp_len"])])
    assert np.array_equal(dense, g["rbsp"])
    names = {n: i for i, n in enumerate(g["field_names"])}
    for mine, theirs in (("priority_id", "PriorityId"), ("view_id", "ViewId"), ("temporal_id", "TemporalId"),
                         ("svc_extension_flag", "SvcExtensionFlag"), ("avc_3d_extension_flag", "Avc3dExtensionFlag"),
                         ("view_idx", "ViewIdx"), ("dependency_id", "DependencyId"), ("quality_id", "QualityId")):
        assert np.array_equal(ext[mine].astype(np.int64), g["fields"][:, names[theirs]]), mine


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["spec_or", "ref_shift"])
def test_gpu_cabac_matches_fixture(ctx, name):
    from h264decode_b200 import capi
    g = load("cabac_16x3000.npz")
    n = len(g["n_ops"])
    stride = g["data"].shape[1]
    flags = capi.CABAC_FINAL_TERMINATE | (capi.BYPASS_SPEC_OR if name == "spec_or" else 0)
    off = np.arange(n, dtype=np.uint64) * stride
    for init in (g["init_states"], None):   # given states, and the in-kernel K4 rule from (qp, idc)
        bins, fin, fst = ctx.cabac_decode(g["data"].reshape(-1), off, g["lens"].astype(np.uint32), g["ops"], g["n_ops"],
                                          int(g["n_ctx"]), qp=g["qp"], idc=g["idc"], init_states=init, flags=flags)
        for s in range(n):
            panicked = g["status_" + name][s] != 0   # the reference would have read past the slice's last byte
            assert bool(fin["flags"][s] & capi.F_OVERRUN) == bool(panicked), s
            if panicked:                             # bins decoded before that point still agree
                nb = int(g["final_" + name][s][3])
                assert np.array_equal(bins[s][:nb // 32], g["bins_" + name][s, :nb // 32]), s
                continue
            total = int(g["n_ops"][s]) + 1
            nw = (total + 31) // 32
            got = bins[s][:nw].copy()
            if total % 32:
                got[-1] &= np.uint32((1 << (total % 32)) - 1)
            assert np.array_equal(got, g["bins_" + name][s, :nw]), s
            assert (fin["cod_i_range"][s], fin["cod_i_offset"][s], fin["bits_read"][s], fin["n_bins"][s]) == tuple(
                g["final_" + name][s]), s
            assert np.array_equal(fst[s], g["states_" + name][s]), s


@pytest.mark.gpu
def test_gpu_ctx_init_matches_fixture(ctx):
    g = load("ctx_init_sweep.npz")
    assert np.array_equal(ctx.ctx_init(g["qp"], g["idc"], 1024, 0), g["states_ref"])
    assert np.array_equal(ctx.ctx_init(g["qp"], g["idc"], 1024, 1), g["states_spec"])


# ------------------------------------------------------------------------------------------------ rows S1 / f1 / f3 / f4 / I5
def _records_equal(got, exp, dtype, what):
    """field-wise comparison of two structured arrays where the reference did not panic (after a panic only `status` is
    specified)"""
    assert np.array_equal(got["status"], exp["status"]), what
    ok = exp["status"] == 0
    for name in dtype.names:
        if name in ("status", "reserved"):
            continue
        assert np.array_equal(got[name][ok], exp[name][ok]), (what, name)


def test_oracle_reproduces_param_set_and_glue_fixtures():
    from oracle import oracle as orc
    from h264decode_b200 import capi
    g = load("param_sets_headers.npz")
    for name, fn, mine, theirs in (("sps", orc.new_sps, capi.SPS_SCALARS, orc._SPS_SCALARS),
                                   ("pps", orc.new_pps, capi.PPS_SCALARS, orc._PPS_SCALARS)):
        for i in range(len(g[name])):
            rb = g[name + "_data"][int(g[name + "_off"][i]):int(g[name + "_off"][i]) + int(g[name + "_len"][i])]
            st, f = fn(rb)
            assert st == g[name][i]["status"], (name, i)
            if st == orc.OK:
                assert [f[b] for a, b in zip(mine, theirs)] == [int(g[name][i][a]) for a, b in zip(mine, theirs)], (name, i)
                assert f["bits_read"] == int(g[name][i]["bits_read"])
    q = load("ctx_glue.npz")
    assert [orc.ctx_idx(int(b), 7, int(o)) for b, o in zip(q["bin_idx"], q["offset"])] == list(q["ctx_idx"])
    for i in range(len(q["se"])):
        bz = orc.new_binarization(int(q["se"][i]), int(q["st"][i]))
        assert [bz[k] for k in orc.BINARIZATION_FIELDS] == list(q["binarization"][i])
    # a few values read by hand from the reference's tables
    at = {(int(b), int(o)): int(c) for b, o, c in zip(q["bin_idx"], q["offset"], q["ctx_idx"])}
    assert at[(1, 3)] == 276 and at[(6, 3)] == 7 and at[(2, 69)] == 10000 and at[(-2, 21)] == -2 and at[(3, 40)] == 5


@pytest.mark.gpu
def test_gpu_param_sets_headers_and_glue_match_fixtures(ctx):
    from h264decode_b200 import capi
    g = load("param_sets_headers.npz")
    _records_equal(ctx.parse_sps(g["sps_data"], g["sps_off"], g["sps_len"]), g["sps"], capi.SPS_DTYPE, "sps")
    _records_equal(ctx.parse_pps(g["pps_data"], g["pps_off"], g["pps_len"]), g["pps"], capi.PPS_DTYPE, "pps")
    assert (g["sps"]["status"] == 0).sum() > 30 and (g["sps"]["status"] != 0).sum() > 10
    hdr = g["hdr"]
    got = np.zeros(len(hdr), capi.SLICE_HEADER_DTYPE)
    for i in range(len(hdr)):   # every header has its own parameter sets: one call each
        ps = capi.Context.param_sets(**dict(zip(capi.PARAM_SET_FIELDS, [int(x) for x in g["hdr_param_sets"][i]])))
        got[i] = ctx.slice_headers(ps, g["hdr_data"], g["hdr_off"][i:i + 1], g["hdr_len"][i:i + 1], g["hdr_nal_type"][i:i + 1],
                                   g["hdr_ref_idc"][i:i + 1])[0]
    assert np.array_equal(got["status"], hdr["status"])
    ok = hdr["status"] == 0
    for name in capi.SLICE_HEADER_DTYPE.names:
        if name not in ("status", "reserved"):
            assert np.array_equal(got[name][ok], hdr[name][ok]), name
    q = load("ctx_glue.npz")
    assert np.array_equal(ctx.ctx_idx(q["bin_idx"], np.full(len(q["bin_idx"]), 7), q["offset"]), q["ctx_idx"])
    bz = ctx.new_binarization(q["se"], q["st"])
    assert np.array_equal(np.stack([bz[k] for k in capi.BINARIZATION_FIELDS], 1), q["binarization"])
    ln, bits = ctx.mb_bin_string(q["mb_st"], q["mb_type"], q["mb_sub"])
    assert np.array_equal(ln, q["mb_len"]) and np.array_equal(bits, q["mb_bits"])
