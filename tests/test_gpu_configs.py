"""BASELINE.json configs[2] and configs[4] at (or near) their stated sizes, on one GPU: the full-size cases the oracle
cannot cover directly are checked through size-independent properties plus oracle / encoder samples."""
import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_config2_ctx_init_sweep_one_million_slices():
    """configs[2]: m/n init over all ctxIdx x SliceQPY 0-51 x cabac_init_idc {-1 (I/SI column), 0, 1, 2}, 1 048 576
    slices with N_CTX = 1024 (1 GiB of states written on the device).  Every row must be the oracle's row for its
    (qp, idc): checked on the device against the 208 distinct rows."""
    import torch
    from h264decode_b200 import capi
    dev = "cuda:0"
    n, n_ctx = 1 << 20, 1024
    qp, idc = hz.slice_params(n)
    ctx = capi.Context(0)
    try:
        p = capi.Context.slice_qp(qp, idc)
        d_p = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
        d_st = torch.empty((n, n_ctx), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx.ctx_init_dev(d_p.data_ptr(), n, n_ctx, d_st.data_ptr(), 0)
        ctx.sync()
        uq = np.arange(52, dtype=np.int32)
        rows = np.stack([orc.ctx_init(uq, np.full(52, c, np.int32), n_ctx) for c in (-1, 0, 1, 2)])  # [4][52][1024]
        d_rows = torch.from_numpy(rows.reshape(4 * 52, n_ctx)).to(dev)
        key = torch.from_numpy(((idc + 1) * 52 + qp).astype(np.int64)).to(dev)
        for lo in range(0, n, 1 << 18):   # 256 Ki slices at a time keeps the gathered copy small
            assert torch.equal(d_st[lo:lo + (1 << 18)], d_rows[key[lo:lo + (1 << 18)]])
        # the reference's fall-through: ctxIdx outside the populated ranges is state (62, 0) for every qp
        assert int(d_st[12345, 500]) == 62 and int(d_st[777, 1023]) == 62
    finally:
        ctx.close()


def test_config4_multi_camera_batch_rank_share():
    """configs[4], scaled to what one test can generate: many independent streams with skewed sizes, dealt to 8 ranks
    by LPT; this GPU plays rank 0 and pushes its share through the asynchronous stream API (three jobs in flight).
    NAL / slice counts, bin totals and every decoded bin (vs what the test encoder coded) must match."""
    from h264decode_b200 import capi, sharding
    rng = np.random.default_rng(4096)
    n_streams = 192
    # slice size = 1 KB * 2^(10 u^3), here as mean bins per slice of a stream
    mean_bins = (1024 * 2 ** (7 * rng.random(n_streams) ** 3) * 8 / 0.88).astype(np.int64)
    n_slices = rng.integers(2, 9, n_streams)
    sizes = mean_bins * n_slices
    parts = sharding.lpt_assign(sizes, 8)
    assert sharding.imbalance(sizes, parts) < 1.15
    mine = parts[0]
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    ctx = capi.Context(0)
    try:
        jobs = [hz.build_stream_cabac(int(n_slices[i]), int(mean_bins[i]), config=5, n_active=64, n_ctx=64,
                                      slices_per_frame=4, frames_per_params=2, id_base=100 * int(i)) for i in mine]
        pending, results = [], []
        for b in jobs:
            if len(pending) == 3:
                results.append(ctx.stream_wait(*pending.pop(0)))
            pending.append(ctx.stream_submit(b["stream"], b["ops"], b["n_ops"], b["qp"], b["idc"], b["n_ctx"], flags=flags))
        while pending:
            results.append(ctx.stream_wait(*pending.pop(0)))
        total = 0
        for b, r in zip(jobs, results):
            ns = len(b["n_ops"])
            assert len(r["final"]) == ns and not (r["final"]["flags"] & capi.F_OVERRUN).any()
            assert np.array_equal(r["final"]["n_bins"], b["n_ops"] + 1)
            for s in range(ns):   # the decoded bins are the ones the encoder coded
                nw = (int(b["n_ops"][s]) + 1) // 32
                assert np.array_equal(r["bins"][s][:nw], b["bins"][s, :nw])
            total += r["total_bins"]
        assert total == int(sum(int(b["n_ops"].sum()) + len(b["n_ops"]) for b in jobs))
        # one of them against the oracle as well
        b, r = jobs[0], results[0]
        onal, _ = orc.read_nal_units_arrays(b["stream"])
        assert np.array_equal(r["nals"]["start"].astype(np.int64), onal["start"])
        assert np.array_equal(r["nals"]["rbsp_len"].astype(np.int64), onal["rbsp_len"])
    finally:
        ctx.close()


def test_raw_stream_pipeline_at_scale():
    """H264B_STREAM_PARAM_SETS on a configs[3]-shaped stream (8000 slice NAL units, SPS + PPS every 25 frames): parameter
    sets, the sets each slice uses, every slice header (the payload bytes read as a header: arbitrary bits, so all three
    outcomes of the reference's walk occur) and, for the slices whose header the reference can walk, the engine's final
    state -- all against the oracle, through the asynchronous job API with three jobs in flight."""
    from h264decode_b200 import capi
    from tests.test_param_sets import compare_sps, compare_pps
    from tests.test_slice_header import compare as compare_header
    n_slices = 8000
    b = hz.build_stream_cabac(n_slices, 600, config=3, n_active=64, n_ctx=64, slices_per_frame=8, frames_per_params=25,
                              id_base=31)
    stream = b["stream"]
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE | capi.STREAM_PARAM_SETS
    ctx = capi.Context(0)
    try:
        tks = [ctx.stream_submit(stream, b["ops"], b["n_ops"], None, None, 64, flags=flags, max_slices=n_slices + 5,
                                 max_sps=64, max_pps=64) for _ in range(3)]
        rs = [ctx.stream_wait(*t) for t in tks]
    finally:
        ctx.close()
    onal, orbsp = orc.read_nal_units_arrays(stream)
    rb_of = lambda k: orbsp[int(onal["rbsp_off"][k]):int(onal["rbsp_off"][k]) + int(onal["rbsp_len"][k])]  # noqa: E731
    i_sps, i_pps = np.flatnonzero(onal["type"] == 7), np.flatnonzero(onal["type"] == 8)
    sl = np.flatnonzero((onal["type"] == 1) | (onal["type"] == 5))
    assert len(i_sps) == 40 and len(i_pps) == 40 and len(sl) == n_slices
    sps_f = [orc.new_sps(rb_of(k)) for k in i_sps]
    pps_f = [orc.new_pps(rb_of(k)) for k in i_pps]
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    for r in rs:
        assert np.array_equal(r["sps_nal"], i_sps) and np.array_equal(r["pps_nal"], i_pps)
        assert np.array_equal(r["slice_nal"], sl)
    r = rs[2]
    for j in range(40):
        compare_sps(r["sps"][j], *sps_f[j], j)
        compare_pps(r["pps"][j], *pps_f[j], j)
    exp_sps = np.searchsorted(i_sps, sl) - 1
    exp_pps = np.searchsorted(i_pps, sl) - 1
    assert np.array_equal(r["slice_sps"], exp_sps) and np.array_equal(r["slice_pps"], exp_pps)
    counts = {0: 0, 1: 0, 2: 0}
    decoded = 0
    for s, k in enumerate(sl):
        rc, h = orc.new_slice_header(sps_f[exp_sps[s]][1], pps_f[exp_pps[s]][1], int(onal["type"][k]), int(onal["ref_idc"][k]),
                                     rb_of(k))
        hdr = r["headers"][s]
        compare_header(hdr, int(hdr["status"]), rc, h, s)
        counts[rc] += 1
        if rc != orc.OK:
            assert r["final"]["flags"][s] & capi.F_OVERRUN, s   # not decoded: no data
            continue
        if decoded < 150:   # the engine on what follows the "header": same op schedule, (SliceQPY, idc) from the header
            t5 = h["SliceType"] % 5 if 0 <= h["SliceType"] <= 9 else -1
            idc = -1 if t5 in (2, 4) else max(min(h["CabacInit"], 1000), -2)
            qp = max(min(h["SliceQPy"], 1000000), -1000000)
            init = orc.ctx_init(np.array([qp], np.int32), np.array([idc], np.int32), 64)[0]
            skip = min((h["bits_read"] + 7) // 8, len(rb_of(k)))
            rc2, bins, fin, _ = orc.cabac_decode_slice(rb_of(k)[skip:], np.concatenate([b["ops"][:b["n_ops"][s]], term]), init,
                                                       orc.BYPASS_SPEC_OR)
            if rc2 == orc.OK:
                nw = (int(b["n_ops"][s]) + 1) // 32
                assert np.array_equal(r["bins"][s][:nw], bins[:nw]), s
                assert (r["final"]["cod_i_range"][s], r["final"]["cod_i_offset"][s], r["final"]["bits_read"][s]) == (
                    fin["codIRange"], fin["codIOffset"], fin["bitsRead"]), s
            else:
                assert r["final"]["flags"][s] & capi.F_OVERRUN, s
            decoded += 1
    assert counts[0] > 100 and counts[1] + counts[2] > 10 and decoded == 150, counts
    # the three jobs in flight saw the same stream: same results
    for other in rs[:2]:
        assert np.array_equal(other["headers"], r["headers"]) and np.array_equal(other["final"], r["final"])
        assert np.array_equal(other["bins_flat"], r["bins_flat"])


@pytest.mark.parametrize("bypass_spec_or", [True, False])
@pytest.mark.parametrize("tables_spec", [False, True])
def test_config1_full_size_64_slices_x_100k_bins(bypass_spec_or, tables_spec):
    """configs[1] at its stated size: 64 independent slices x 100 000 bins (test-encoder generated, one shared op
    schedule of mixed decision / bypass / terminate ops, 64 active contexts), every bin, every final
    (codIRange, codIOffset, bitsRead) and every final context state against the oracle, for both bypass forms and both
    table sets.  (The encoder codes for the SPEC_OR bypass form; under the reference's own bypass form the same bytes are
    just bytes, which the literal engine and the oracle must still decode alike, overruns included.)"""
    from h264decode_b200 import capi
    n_slices, n_bins, n_active, n_ctx = 64, 100_000, 64, 64
    fo = (orc.BYPASS_SPEC_OR if bypass_spec_or else 0) | (orc.TABLES_SPEC if tables_spec else 0)
    fg = (capi.BYPASS_SPEC_OR if bypass_spec_or else 0) | (capi.TABLES_SPEC if tables_spec else 0) | capi.CABAC_FINAL_TERMINATE
    ops = hz.gen_schedule(2, n_bins, n_active)
    n_ops = np.full(n_slices, n_bins, np.uint32)
    qp, idc = hz.slice_params(n_slices, first=7)
    g = hz.gen_cabac_slices(2, ops, n_ops, n_active, n_ctx, qp, idc, flags=hz.TABLES_SPEC if tables_spec else 0)
    stride = g["data"].shape[1]
    data = np.concatenate([g["data"].reshape(-1), np.zeros(64, np.uint8)])
    off = np.arange(n_slices, dtype=np.uint64) * stride
    length = g["lens"].astype(np.uint32)
    ctx = capi.Context(0)
    try:
        bins, fin, fst = ctx.cabac_decode(data, off, length, ops, n_ops, n_ctx, qp=qp, idc=idc, flags=fg)
    finally:
        ctx.close()
    init = orc.ctx_init(qp, idc, n_ctx, fo & orc.TABLES_SPEC)
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    full_ops = np.concatenate([ops, term])
    nw = (n_bins + 1 + 31) // 32
    for s in range(n_slices):
        rc, obins, ofin, ost = orc.cabac_decode_slice(data[int(off[s]):int(off[s]) + int(length[s])], full_ops, init[s], fo)
        assert bool(fin["flags"][s] & capi.F_OVERRUN) == (rc == orc.PANIC), s
        if rc == orc.PANIC:
            for w in range(ofin["n_bins"] // 32):
                assert bins[s][w] == obins[w], (s, w)
            continue
        got = bins[s][:nw].copy()
        got[-1] &= np.uint32((1 << ((n_bins + 1) % 32)) - 1)
        assert np.array_equal(got, obins[:nw]), "slice %d bins" % s
        assert (fin["cod_i_range"][s], fin["cod_i_offset"][s], fin["bits_read"][s], fin["n_bins"][s]) == (
            ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"], ofin["n_bins"]), s
        assert np.array_equal(fst[s], ost), "slice %d states" % s
        if bypass_spec_or:   # self-check: the decoder returns what the encoder coded
            assert np.array_equal(got, g["bins"][s, :nw]), s


def test_streams_of_very_short_nal_units_and_many_parameter_sets():
    """A stream whose NAL units average far less than 64 bytes (the NAL index of a stream job starts at n / 64 + 1024
    records) and one with more SPS / PPS NAL units than a job's default bound of 64: both entry points must deliver the
    oracle's units -- the job re-runs with the bounds its first pass reported -- instead of failing with E_CAPACITY."""
    from h264decode_b200 import capi
    rng = np.random.default_rng(77)
    parts = []
    for k in range(40000):   # 8- to 11-byte units: start code + header + 3..6 payload bytes
        parts.append(b"\x00\x00\x00\x01" + bytes([0x41 if k % 3 else 0x65]) + bytes(rng.integers(4, 255, int(rng.integers(3, 7))).astype(np.uint8)))
    short = np.frombuffer(b"".join(parts) + b"\x00\x00\x00\x01", np.uint8)
    sps, pps = hz.SPS_NAL, hz.PPS_NAL   # sets the reference parses
    many = []
    for k in range(150):
        many.append(b"\x00\x00\x00\x01" + sps + b"\x00\x00\x00\x01" + pps + b"\x00\x00\x00\x01\x65" + bytes(rng.integers(4, 255, 40).astype(np.uint8)))
    many = np.frombuffer(b"".join(many) + b"\x00\x00\x00\x01", np.uint8)
    ctx = capi.Context(0)
    try:
        for stream in (short, many):
            onal, orbsp = orc.read_nal_units_arrays(stream)
            summ, nals, ext, rbsp = ctx.annexb_scan(stream)
            assert summ["n_nals"] == len(onal["start"])
            assert np.array_equal(nals["start"].astype(np.int64), onal["start"])
            r = ctx.stream_decode(stream, np.zeros(0, np.uint16), None, np.zeros(1, np.int32), np.zeros(1, np.int32), 1)
            assert r["scan"]["n_nals"] == len(onal["start"])
            assert np.array_equal(r["nals"]["start"].astype(np.int64), onal["start"])
            assert np.array_equal(r["nals"]["rbsp_len"].astype(np.int64), onal["rbsp_len"])
        # the stream's own parameter sets, with the default bounds (64 SPS / 64 PPS) and a slice bound that is too small
        t = ctx.stream_submit(many, np.zeros(0, np.uint16), None, None, None, 1, flags=capi.STREAM_PARAM_SETS, max_slices=10)
        r = ctx.stream_wait(*t)
        assert len(r["sps"]) == 150 and len(r["pps"]) == 150 and len(r["headers"]) == 150
        assert np.array_equal(r["slice_sps"], np.arange(150)) and np.array_equal(r["slice_pps"], np.arange(150))
    finally:
        ctx.close()


def test_scheduler_batch_of_streams_over_several_workers():
    """h264b_scheduler: a batch of independent streams (skewed sizes, some with bytes in front of their first start code
    or without a closing start code) dealt to two workers -- here both on the one GPU -- each of which runs its share
    through one split + strip pass and five CABAC launches by slice length.  Every stream's NAL units must be the
    oracle's for that stream on its own, every slice's bins the ones the test encoder coded, whatever worker and launch
    they landed in."""
    from h264decode_b200 import capi
    rng = np.random.default_rng(5)
    n_streams = 40
    mean_bins = (1024 * 2 ** (6 * rng.random(n_streams) ** 3) * 8 / 0.88).astype(np.int64)
    per = rng.integers(1, 7, n_streams)
    built = [hz.build_stream_cabac(int(per[i]), int(mean_bins[i]), config=5, n_active=64, n_ctx=64, slices_per_frame=3,
                                   frames_per_params=2, id_base=1000 * i) for i in range(n_streams)]
    ops = max((b["ops"] for b in built), key=len)   # the schedules are prefixes of one another (same generator)
    for b in built:
        assert np.array_equal(b["ops"], ops[:len(b["ops"])])
    streams = []
    for i, b in enumerate(built):
        s = b["stream"]
        if i % 5 == 1:
            s = np.concatenate([np.array([7, 0, 0, 1, 9], np.uint8), s])          # junk in front of the first start code
        if i % 7 == 2:
            s = np.concatenate([s, np.array([0x65, 1, 2, 3, 4, 5], np.uint8)])    # an unterminated unit at the end
        streams.append(s)
    streams.append(np.array([1, 2, 3], np.uint8))                                 # no start code at all
    per = np.concatenate([per, [0]])
    n_ops = np.concatenate([b["n_ops"] for b in built])
    qp = np.concatenate([b["qp"] for b in built])
    idc = np.concatenate([b["idc"] for b in built])
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    sch = capi.Scheduler([0, 0])
    try:
        r = sch.run(streams, per, ops, n_ops, qp, idc, 64, flags=flags, group_bytes=96 << 10)
    finally:
        sch.close()
    assert set(r["stream_device"][:n_streams]) == {0, 1} and r["stream_device"][-1] == -1
    assert abs(int(r["device_bytes"][0]) - int(r["device_bytes"][1])) < 0.25 * r["device_bytes"].sum()
    row = 0
    for i, b in enumerate(built):
        onal, _ = orc.read_nal_units_arrays(streams[i])
        got = r["nals"][i]
        assert np.array_equal(got["start"].astype(np.int64), onal["start"]), i
        assert np.array_equal(got["rbsp_len"].astype(np.int64), onal["rbsp_len"]), i
        assert np.array_equal(got["type"].astype(np.int64), onal["type"]), i
        for s in range(int(per[i])):
            f = r["final"][row]
            assert f["n_bins"] == b["n_ops"][s] + 1 and not (f["flags"] & capi.F_OVERRUN)
            nw = (int(b["n_ops"][s]) + 1) // 32
            assert np.array_equal(r["bins"][row][:nw], b["bins"][s, :nw]), (i, s)
            assert 0 < r["slice_done_ms"][row] <= r["makespan_ms"]
            row += 1
    assert row == len(n_ops) and r["total_bins"] == int(n_ops.sum()) + len(n_ops)
    assert len(r["nals"][-1]) == 0
