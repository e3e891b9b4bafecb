"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the same inputs.

Bit-exact everywhere (all arithmetic on this path is integer/byte work).  Needs a B200: run with `-m gpu`.
"""
import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

H = bytes.fromhex


@pytest.fixture(scope="module")
def ctx():
    from h264decode_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def capi():
    from h264decode_b200 import capi as m
    return m


# =========================================================================================== K4 context init
def test_ctx_init_exhaustive_sweep(ctx):
    """75 populated ctxIdx x 52 QP x {-1,0,1,2} (SURVEY.md §8d C3 KAT), plus out-of-range qp / idc, REF and SPEC"""
    qps = list(range(-3, 56))
    idcs = [-1, 0, 1, 2, 3, -2, 77]
    qp = np.repeat(qps, len(idcs)).astype(np.int32)
    idc = np.tile(idcs, len(qps)).astype(np.int32)
    for flags_o, flags_g in [(0, 0), (orc.TABLES_SPEC, 1)]:
        for n_ctx in (1024, 460, 75, 16, 1):
            exp = orc.ctx_init(qp, idc, n_ctx, flags_o)
            got = ctx.ctx_init(qp, idc, n_ctx, flags_g)
            assert np.array_equal(got, exp), (flags_g, n_ctx)


def test_ctx_init_scalar_dropins(ctx):
    for m, n, q in [(20, -15, 0), (20, -15, 26), (-28, 127, 26), (-39, 127, 51), (57, 2, 51), (-7, 24, 26), (0, 0, 99),
                    (-4, 127, -5)]:
        assert ctx.pre_ctx_state(m, n, q) == orc.pre_ctx_state(m, n, q)
    for c in (0, 5, 10, 11, 30, 39, 40, 69, 70, 76, 92, 104, 105, 1023, 1024, -1):
        for idc in (-1, 0, 1, 2, 3, -7):
            for fo, fg in ((0, 0), (orc.TABLES_SPEC, 1)):
                assert ctx.mn(c, idc, fg) == orc.mn(c, idc, fo), (c, idc, fg)


def test_ctx_init_many_slices(ctx):
    qp, idc = hz.slice_params(20000)
    exp = orc.ctx_init(qp, idc, 1024)
    got = ctx.ctx_init(qp, idc, 1024)
    assert np.array_equal(got, exp)


# =========================================================================================== K1/K2 Annex-B scan/strip
def gather_rbsp(nals, rbsp):
    """concatenate the per-NAL RBSPs (dense, in NAL order) from the position-preserving buffer"""
    if len(nals) == 0:
        return np.zeros(0, np.uint8)
    lens = nals["rbsp_len"].astype(np.int64)
    offs = nals["rbsp_off"].astype(np.int64)
    dense_off = np.concatenate([[0], np.cumsum(lens)[:-1]])
    idx = np.repeat(offs - dense_off, lens) + np.arange(int(lens.sum()))
    return rbsp[idx]


def assert_rbsp_equal(nals, rbsp, onal, orbsp):
    got = gather_rbsp(nals, rbsp)
    assert len(got) == len(orbsp)
    if not np.array_equal(got, orbsp):
        bad = int(np.flatnonzero(got != orbsp)[0])
        k = int(np.searchsorted(np.cumsum(onal["rbsp_len"]), bad, side="right"))
        raise AssertionError("RBSP differs at dense byte %d (NAL %d, start %d)" % (bad, k, onal["start"][k]))


def check_scan(ctx, capi, stream):
    s = np.ascontiguousarray(stream, dtype=np.uint8)
    summ, nals, ext, rbsp = ctx.annexb_scan(s)
    onal, orbsp = orc.read_nal_units_arrays(s)
    n = len(onal["start"])
    assert summ["n_nals"] == n, (summ, n)
    assert summ["rbsp_bytes"] == len(orbsp)
    assert np.array_equal(nals["start"].astype(np.int64), onal["start"])
    assert np.array_equal(nals["num_bytes"].astype(np.int64), onal["num_bytes"])
    assert np.array_equal(nals["rbsp_len"].astype(np.int64), onal["rbsp_len"])
    # position-preserving layout: a NAL's RBSP sits at its body's own offset
    assert np.array_equal(nals["rbsp_off"].astype(np.int64), onal["start"] + onal["header_bytes"])
    assert np.array_equal(nals["forbidden_zero_bit"].astype(np.int64), onal["fzb"])
    assert np.array_equal(nals["ref_idc"].astype(np.int64), onal["ref_idc"])
    assert np.array_equal(nals["type"].astype(np.int64), onal["type"])
    assert np.array_equal(nals["header_bytes"].astype(np.int64), onal["header_bytes"])
    assert_rbsp_equal(nals, rbsp, onal, orbsp)
    assert np.array_equal((nals["flags"] & capi.F_HAS_EPB) != 0, onal["epb"] == 3)
    if n:
        assert summ["first_start"] == onal["start"][0]
        body = np.maximum(onal["num_bytes"] - onal["header_bytes"] - 2, 0)
        assert summ["n_epb"] == int((body - onal["rbsp_len"]).sum())
        f = {name: i for i, name in enumerate(orc._NAL_FIELDS)}
        pairs = [("svc_extension_flag", "SvcExtensionFlag"), ("avc_3d_extension_flag", "Avc3dExtensionFlag"),
                 ("idr_flag", "IdrFlag"), ("priority_id", "PriorityId"),
                 ("no_inter_layer_pred_flag", "NoInterLayerPredFlag"), ("dependency_id", "DependencyId"),
                 ("quality_id", "QualityId"), ("temporal_id", "TemporalId"),
                 ("use_ref_base_pic_flag", "UseRefBasePicFlag"), ("discardable_flag", "DiscardableFlag"),
                 ("output_flag", "OutputFlag"), ("reserved_three_2bits", "ReservedThree2Bits"),
                 ("non_idr_flag", "NonIdrFlag"), ("view_id", "ViewId"), ("anchor_pic_flag", "AnchorPicFlag"),
                 ("inter_view_flag", "InterViewFlag"), ("reserved_one_bit", "ReservedOneBit"), ("view_idx", "ViewIdx"),
                 ("depth_flag", "DepthFlag")]
        for g, o in pairs:
            assert np.array_equal(ext[g].astype(np.int64), onal["fields"][:, f[o]]), g
    return summ


def test_scan_kats(ctx, capi):
    from tests.test_oracle_kat import B2_STREAM
    from tests.test_hd_logic import random_stream  # noqa: F401
    check_scan(ctx, capi, np.frombuffer(B2_STREAM, np.uint8))
    for hexs in ["", "00", "00000001", "0000000100000001", "000000014100000000010000000165",
                 "00 00 00 00 00 01 00 00 03 00 00 00 01",
                 "00000001 6E 80 00 00 00 03 55 66 77 88 00 00 00 01",
                 "00000001 74 C5 A6 9F 11 22 33 44 00 00 00 01 75 FF 80 AA BB CC DD 00 00 00 01 75 7F 80 AA 00 00 00 01"
                 "6E 40 00 01 99 88 77 00 00 00 01",
                 "00000001 67 640028ACD94078022640 00000001 68 EE0F2C8B 00000001 65 8884 00000001"]:
        check_scan(ctx, capi, np.frombuffer(H(hexs), np.uint8))
    check_scan(ctx, capi, np.zeros(100000, np.uint8))
    check_scan(ctx, capi, np.tile(np.array([0, 0, 3], np.uint8), 20000))
    check_scan(ctx, capi, np.tile(np.array([0, 0, 0, 1], np.uint8), 20000))
    check_scan(ctx, capi, np.full(70000, 0xFF, np.uint8))
    # one NAL over several 16 KiB tiles that loses an EPB every third byte / only in its first tile / only late
    sc, hdr = np.array([0, 0, 0, 1], np.uint8), np.array([0x65], np.uint8)
    rng = np.random.default_rng(77)
    body = rng.integers(4, 256, 60000).astype(np.uint8)
    check_scan(ctx, capi, np.concatenate([sc, hdr, np.tile(np.array([0, 0, 3], np.uint8), 23000), sc, hdr, body, sc]))
    b2 = body.copy(); b2[100:103] = [0, 0, 3]
    check_scan(ctx, capi, np.concatenate([sc, hdr, b2, sc]))
    b3 = body.copy(); b3[40000:40003] = [0, 0, 3]; b3[16380:16383] = [0, 0, 3]
    check_scan(ctx, capi, np.concatenate([np.full(7, 9, np.uint8), sc, hdr, b3, sc, hdr, b2[:20000], sc]))


@pytest.mark.parametrize("seed", range(4))
def test_scan_random_streams(ctx, capi, seed):
    from tests.test_hd_logic import random_stream
    rng = np.random.default_rng(seed)
    T = 16384
    sizes = [1, 3, 4, 5, 15, 16, 17, 31, 33, 1000, 2047, 2048, 2049, 2064, 4095, T - 1, T, T + 1, 2 * T - 3, 2 * T + 5, 5 * T + 7,
             40 * T + 11]
    for n in sizes:
        for p_zero, p_sc in [(0.5, 0.02), (0.2, 0.002), (0.9, 0.0005), (0.02, 0.0001)]:
            check_scan(ctx, capi, random_stream(rng, n, p_zero, p_sc, ext_types=(seed % 2 == 0)))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_scan_chunk_boundaries(ctx, capi, seed):
    """2 KiB chunks are processed independently: NALs that span chunks, EPBs just before / after chunk boundaries,
    sparse streams where most chunks take the verbatim-copy kernel and a few the general one"""
    from tests.test_hd_logic import random_stream
    rng = np.random.default_rng(100 + seed)
    C = 2048
    for n in [1, C - 1, C, C + 1, 2 * C + 5, 3 * C - 2, 7 * C + 9, 64 * C, 64 * C + 1, 130 * C + 77, 400 * C + 3]:
        for p_zero, p_sc in [(0.5, 0.02), (0.2, 0.002), (0.9, 0.0005), (0.02, 0.0001), (0.004, 0.00002)]:
            check_scan(ctx, capi, random_stream(rng, n, p_zero, p_sc, ext_types=(seed % 2 == 0)))
    sc, hdr = np.array([0, 0, 0, 1], np.uint8), np.array([0x65], np.uint8)
    body = rng.integers(4, 256, 300000).astype(np.uint8)
    for at in ([100], [2040, 2047, 2050], [4093], [131070, 131073], [5000, 140000, 290000]):
        b = body.copy()
        for p in at:
            b[p:p + 3] = [0, 0, 3]
        check_scan(ctx, capi, np.concatenate([sc, hdr, b, sc, hdr, body[:5000], sc]))


def test_scan_carry_across_many_chunks(ctx, capi):
    """What a NAL has lost so far travels from chunk to chunk (look-back over the per-chunk counts) and, for chunks the
    copy kernel stored verbatim, through the segmented scan and the re-copy list.  Cases the random streams do not reach:
    a NAL that loses more than 2^15 bytes (the per-chunk count field is 15 bits wide, the carry is not), one early EPB
    in front of a megabyte of clean payload (hundreds of verbatim chunks to shift), dirty and clean chunks in turn,
    a long NAL whose every chunk is dirty, and the same with start codes sprinkled in."""
    rng = np.random.default_rng(77)
    sc, hdr = np.array([0, 0, 0, 1], np.uint8), np.array([0x65], np.uint8)
    clean = lambda n: rng.integers(4, 256, n).astype(np.uint8)
    # (a) 70 000 EPBs in one NAL, then clean payload inside the same NAL, then another NAL
    triples = np.tile(np.array([0, 0, 3], np.uint8), 70000)
    check_scan(ctx, capi, np.concatenate([sc, hdr, clean(100), triples, clean(50000), sc, hdr, clean(3000), sc]))
    # (b) one EPB at the front of 1.2 MB of clean payload; the NAL ends in the middle of a chunk
    b = clean(1200000 + 1234)
    b[10:13] = [0, 0, 3]
    check_scan(ctx, capi, np.concatenate([sc, hdr, b, sc, hdr, clean(100), sc]))
    # (c) an EPB in every fourth chunk only
    b = clean(64 * 2048 + 321)
    for k in range(0, 64, 4):
        at = k * 2048 + int(rng.integers(8, 2000))
        b[at:at + 3] = [0, 0, 3]
    check_scan(ctx, capi, np.concatenate([clean(7), sc, hdr, b, sc]))
    # (d) every chunk of a 300 KB NAL dirty, then the same payload cut into NAL units of odd lengths
    b = clean(300000)
    at = np.sort(rng.choice(np.arange(4, 300000 - 4, 97), 2500, replace=False))
    for p_ in at:
        b[p_:p_ + 3] = [0, 0, 3]
    check_scan(ctx, capi, np.concatenate([sc, hdr, b, sc]))
    parts, pos = [], 0
    while pos < len(b):
        ln = int(rng.integers(50, 9000))
        parts += [sc, hdr, b[pos:pos + ln]]
        pos += ln
    check_scan(ctx, capi, np.concatenate(parts + [sc]))


@pytest.mark.parametrize("T", [2048, 16384, 131072])
def test_scan_start_codes_across_tile_boundaries(ctx, capi, T):
    rng = np.random.default_rng(9)
    for shift in range(-8, 9):
        s = rng.integers(4, 256, 3 * T + 100).astype(np.uint8)
        s[0:5] = [0, 0, 0, 1, 0x65]
        for t in (T, 2 * T):
            p = t + shift
            s[p - 4:p] = [0, 0, 0, 1]          # start code ending just before p
            s[p] = [0x65, 0x6E, 0x75][shift % 3]
            s[p + 1] = 0x80 if shift % 2 else 0x00
            s[p + 2:p + 7] = [0, 0, 3, 0, 0]    # zeros + EPB candidates right behind the header
            s[p - 7:p - 4] = [0, 0, 3]          # an EPB just before the start code
        s[-4:] = [0, 0, 0, 1]
        check_scan(ctx, capi, s)


def test_scan_c1_stream(ctx, capi):
    """BASELINE configs[0]: synthetic 1 MB stream, SPS+PPS+slices, EPB-bearing payloads; NAL/RBSP/SPS/PPS bit-exact"""
    s = hz.build_stream_c1(1 << 20)
    summ = check_scan(ctx, capi, s)
    assert summ["n_epb"] > 1000
    _, nals, _, rbsp = ctx.annexb_scan(s)
    sps = rbsp[nals["rbsp_off"][0]:nals["rbsp_off"][0] + nals["rbsp_len"][0]]
    pps = rbsp[nals["rbsp_off"][1]:nals["rbsp_off"][1] + nals["rbsp_len"][1]]
    st, f = orc.new_sps(sps)
    assert st == orc.OK and f["Profile"] == 100 and f["PicWidthInMbsMinus1"] == 119
    st, f = orc.new_pps(pps)
    assert st == orc.OK and f["EntropyCodingMode"] == 1 and f["PicInitQpMinus26"] == -3


def test_scan_large_properties(ctx, capi):
    """64 MiB: size-independent properties (the oracle would take too long in literal mode; non-literal is fine)"""
    rng = np.random.default_rng(3)
    n = 64 << 20
    s = rng.integers(0, 256, n, dtype=np.uint8)
    pos = np.sort(rng.choice(n - 64, 3000, replace=False))
    pos = pos[np.diff(np.concatenate([[-100], pos])) > 16]
    for p in pos:
        s[p:p + 5] = [0, 0, 0, 1, 0x41]
    s[:5] = [0, 0, 0, 1, 0x65]
    s[-4:] = [0, 0, 0, 1]
    summ, nals, _, rbsp = ctx.annexb_scan(s, want_ext=False)
    onal, orbsp = orc.read_nal_units_arrays(s)
    assert summ["n_nals"] == len(onal["start"])
    assert np.array_equal(nals["start"].astype(np.int64), onal["start"])
    assert np.array_equal(nals["rbsp_len"].astype(np.int64), onal["rbsp_len"])
    assert_rbsp_equal(nals, rbsp, onal, orbsp)
    assert summ["rbsp_bytes"] == len(orbsp)
    # conservation: every NAL byte is a header byte, one of the 2 tail bytes, an EPB or an RBSP byte
    assert int(nals["num_bytes"].sum()) == int(nals["header_bytes"].sum()) + 2 * len(nals) + summ["n_epb"] + len(orbsp)
    assert summ["n_epb"] > 0


def test_scan_byte_ranges_of_one_stream(ctx, capi):
    """§8e: one stream cut at start codes into byte ranges (h264decode_b200/sharding.py), each range scanned on its own
    through the C ABI: the concatenation is the whole stream's scan, NAL by NAL, RBSP byte by byte"""
    from h264decode_b200 import sharding
    s = np.ascontiguousarray(hz.build_stream_c1(1 << 20), np.uint8)
    _, wn, wx, wr = ctx.annexb_scan(s)
    wn, wr = wn.copy(), wr.copy()
    for n_shards in (2, 8, 5):
        ranges = sharding.cut_byte_ranges(s, n_shards)
        k = 0
        for b, e in ranges:
            _, nals, _, rbsp = ctx.annexb_scan(s[b:e])
            for j in range(len(nals)):
                assert int(nals["start"][j]) + b == int(wn["start"][k])
                for f in ("num_bytes", "type", "ref_idc", "header_bytes", "rbsp_len"):
                    assert nals[f][j] == wn[f][k]
                assert np.array_equal(rbsp[nals["rbsp_off"][j]:nals["rbsp_off"][j] + nals["rbsp_len"][j]],
                                      wr[wn["rbsp_off"][k]:wn["rbsp_off"][k] + wn["rbsp_len"][k]])
                k += 1
        assert k == len(wn)
        assert max(e - b for b, e in ranges) < len(s) // n_shards + (1 << 17)   # balanced up to one NAL


# =========================================================================================== NewNalUnit (frames)
def test_nal_units_frames(ctx, capi):
    from tests.test_hd_logic import random_stream
    rng = np.random.default_rng(11)
    frames = [H(x) for x in ["67 42 00 00 03 01 AA BB 00 00 00 01", "00 00 03 07 00 00 03 00 00 00 01", "65 11 00 00 03",
                             "65 11 22 00 00 03 44", "65 11 22 33", "6E 80 00 00 00 03 55 66 77 88", "65", "65 11",
                             "74 C5 A6 9F 11 22 33 44", "75 FF 80 AA BB CC DD", "75 7F 80 AA BB CC DD"]]
    for _ in range(300):
        n = int(rng.integers(1, 9000 if rng.random() < 0.05 else 80))
        f = random_stream(rng, n, 0.5, 0.0)
        f[0] = [0x65, 0x6E, 0x74, 0x75, 0x00, 0x67][rng.integers(0, 6)]
        if rng.random() < 0.3 and n >= 3:
            f[-3:] = [0, 0, 3]
        frames.append(f.tobytes())
    nals, ext, outs = ctx.nal_units(frames)
    for i, f in enumerate(frames):
        st, o, rb = orc.new_nal_unit(f)
        if st == orc.PANIC:
            assert nals["flags"][i] & capi.F_OVERRUN
            continue
        assert not (nals["flags"][i] & capi.F_OVERRUN)
        assert outs[i] == rb, f.hex()
        assert (nals["type"][i], nals["ref_idc"][i], nals["header_bytes"][i], nals["num_bytes"][i]) == (
            o["Type"], o["RefIdc"], o["HeaderBytes"], o["NumBytes"])
        assert bool(nals["flags"][i] & capi.F_HAS_EPB) == (o["EmulationPreventionThreeByte"] == 3)
        assert ext["priority_id"][i] == o["PriorityId"] and ext["view_id"][i] == o["ViewId"]


# =========================================================================================== K3 CABAC engine
def oracle_decode_all(data, off, length, ops, n_ops, init, flags_o, final_term):
    out = []
    for s in range(len(off)):
        sl = ops[:n_ops[s]]
        if final_term:
            sl = np.concatenate([sl, np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)])
        out.append(orc.cabac_decode_slice(data[int(off[s]):int(off[s]) + int(length[s])], sl, init[s], flags_o))
    return out


def compare_cabac(capi, bins, fin, fst, oracle_out, n_ops, final_term):
    for s, (rc, obins, ofin, ost) in enumerate(oracle_out):
        total = int(n_ops[s]) + (1 if final_term else 0)
        assert bool(fin["flags"][s] & capi.F_OVERRUN) == (rc == orc.PANIC), s
        if rc == orc.PANIC:
            nb = ofin["n_bins"]           # bins decoded before the reference would have panicked still agree
            for w in range(nb // 32):
                assert bins[s][w] == obins[w]
            continue
        nw = (total + 31) // 32
        got = bins[s][:nw].copy()
        if total % 32:
            got[-1] &= np.uint32((1 << (total % 32)) - 1)
        assert np.array_equal(got, obins[:nw]), "slice %d bins" % s
        assert (fin["cod_i_range"][s], fin["cod_i_offset"][s], fin["bits_read"][s], fin["n_bins"][s]) == (
            ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"], ofin["n_bins"]), s
        if fst is not None:
            assert np.array_equal(fst[s], ost), "slice %d states" % s


def pack_slices(datas, rng=None, pad=8):
    """concatenate slices at arbitrary (unaligned) offsets"""
    off, length, parts, pos = [], [], [], 0
    for d in datas:
        gap = int(rng.integers(0, 7)) if rng is not None else 0
        parts.append(np.full(gap, 0x5A, np.uint8))
        pos += gap
        off.append(pos)
        length.append(len(d))
        parts.append(np.asarray(d, np.uint8))
        pos += len(d)
    parts.append(np.zeros(pad, np.uint8))
    return np.concatenate(parts), np.array(off, np.uint64), np.array(length, np.uint32)


@pytest.mark.parametrize("tables_spec", [False, True])
@pytest.mark.parametrize("n_slices,n_active,n_ctx,n_bins", [(64, 64, 64, 6000), (300, 460, 512, 1500), (5, 64, 1024, 400)])
def test_cabac_encoder_streams(ctx, capi, tables_spec, n_slices, n_active, n_ctx, n_bins):
    """BASELINE configs[1] shape (64 slices, shared schedule, mixed decision/bypass/terminate) at oracle-sized bins"""
    rng = np.random.default_rng(n_slices)
    fo = orc.BYPASS_SPEC_OR | (orc.TABLES_SPEC if tables_spec else 0)
    fg = capi.BYPASS_SPEC_OR | (capi.TABLES_SPEC if tables_spec else 0) | capi.CABAC_FINAL_TERMINATE
    ops = hz.gen_schedule(2, n_bins, n_active)
    n_ops = rng.integers(1, n_bins + 1, n_slices).astype(np.uint32)
    n_ops[0] = n_bins
    n_ops[-1] = 31
    qp, idc = hz.slice_params(n_slices, first=13)
    g = hz.gen_cabac_slices(2, ops, n_ops, n_active, n_ctx, qp, idc, flags=hz.TABLES_SPEC if tables_spec else 0)
    data, off, length = pack_slices([g["data"][s, :g["lens"][s]] for s in range(n_slices)], rng)
    bins, fin, fst = ctx.cabac_decode(data, off, length, ops, n_ops, n_ctx, qp=qp, idc=idc, flags=fg)
    init = orc.ctx_init(qp, idc, n_ctx, fo & orc.TABLES_SPEC)
    compare_cabac(capi, bins, fin, fst, oracle_decode_all(data, off, length, ops, n_ops, init, fo, True), n_ops, True)
    # self-check: decoded bins == the bins the encoder coded, final states == encoder states
    for s in range(n_slices):
        nw = (int(n_ops[s]) + 1 + 31) // 32
        got = bins[s, :nw].copy()
        if (int(n_ops[s]) + 1) % 32:
            got[-1] &= np.uint32((1 << ((int(n_ops[s]) + 1) % 32)) - 1)
        assert np.array_equal(got, g["bins"][s, :nw])
        assert np.array_equal(fst[s], g["final_states"][s])
    # explicit initial states instead of (qp, idc) give the same result
    bins2, fin2, fst2 = ctx.cabac_decode(data, off, length, ops, n_ops, n_ctx, init_states=init, flags=fg)
    assert np.array_equal(fin2, fin) and np.array_equal(fst2, fst)


@pytest.mark.parametrize("flags_o", [0, orc.TABLES_SPEC, orc.BYPASS_SPEC_OR])
def test_cabac_reference_bypass_and_garbage(ctx, capi, flags_o):
    """REF_SHIFT bypass (the reference's literal form, A5) and arbitrary bytes / op sequences, incl. overruns"""
    rng = np.random.default_rng(21 + flags_o)
    n_slices, n_ctx = 70, 32
    n_max = 3000
    kinds = rng.choice([0, 0, 0, 1, 2], n_max)
    ops = ((kinds << 14) | rng.integers(0, 40, n_max)).astype(np.uint16)   # some ctxIdx >= n_ctx -> treated as 0
    datas = [rng.integers(0, 256, int(rng.integers(1, 600))).astype(np.uint8) for _ in range(n_slices)]
    data, off, length = pack_slices(datas, rng, pad=64)
    n_ops = rng.integers(0, n_max + 1, n_slices).astype(np.uint32)
    init = rng.integers(0, 128, (n_slices, n_ctx)).astype(np.uint8)
    fg = (capi.TABLES_SPEC if flags_o & orc.TABLES_SPEC else 0) | (capi.BYPASS_SPEC_OR if flags_o & orc.BYPASS_SPEC_OR else 0)
    for final_term in (False, True):
        bins, fin, fst = ctx.cabac_decode(data, off, length, ops, n_ops, n_ctx, init_states=init,
                                          flags=fg | (capi.CABAC_FINAL_TERMINATE if final_term else 0))
        compare_cabac(capi, bins, fin, fst, oracle_decode_all(data, off, length, ops, n_ops, init, flags_o, final_term),
                      n_ops, final_term)


def test_engine_step_dropins(ctx, capi):
    """SURVEY.md Appendix B.6 through the scalar entry points"""
    for p, v, R, O in [(33, 0, 330, 200), (33, 0, 330, 280), (33, 1, 400, 333), (0, 0, 510, 270), (62, 1, 256, 250),
                       (63, 0, 300, 298)]:
        assert ctx.binary_decision(p, v, R, O) == orc.binary_decision(p, v, R, O)
        assert ctx.binary_decision(p, v, R, O, capi.TABLES_SPEC) == orc.binary_decision(p, v, R, O, orc.TABLES_SPEC)
    for p, v, b in [(0, 0, 1), (0, 1, 0), (0, 0, 0), (59, 1, 1), (59, 1, 0), (62, 0, 0), (63, 1, 1), (33, 0, 1)]:
        assert ctx.state_transition(p, v, b) == orc.state_transition(p, v, b)
        assert ctx.state_transition(p, v, b, capi.TABLES_SPEC) == orc.state_transition(p, v, b, orc.TABLES_SPEC)
    bits = H("B4 00 00")
    for R, O in [(256, 10), (255, 10), (128, 5), (6, 3), (2, 1)]:
        r = ctx.engine_step(4, R, O, bits)
        er, eo, nb = orc.renorm_d(bits, R, O)
        assert (r["R"], r["O"], r["bits_used"]) == (er, eo, nb)
    for R, O, bit in [(510, 100, 0), (510, 100, 1), (300, 200, 0), (300, 200, 1), (256, 255, 1), (300, (1 << 62) + 5, 1)]:
        d = bytes([0x80 if bit else 0])
        r = ctx.engine_step(capi.OP_BYPASS, R, O, d)
        assert (r["O"], r["bin"]) == orc.decode_bypass(d, R, O)
        r = ctx.engine_step(capi.OP_BYPASS, R, O, d, flags=capi.BYPASS_SPEC_OR)
        assert (r["O"], r["bin"]) == orc.decode_bypass(d, R, O, orc.BYPASS_SPEC_OR)
    for R, O in [(510, 507), (510, 508), (258, 100), (256, 253), (256, 254)]:
        r = ctx.engine_step(capi.OP_TERMINATE, R, O, b"\x80")
        assert (r["R"], r["O"], r["bin"]) == orc.decode_terminate(b"\x80", R, O)
    r = ctx.engine_step(5, 0, 0, H("A5 C3"))
    assert (r["R"], r["O"], r["bits_used"]) == (510, 0b101001011, 9)
    # composed DecodeDecision = BinaryDecision -> StateTransitionProcess -> RenormD
    r = ctx.engine_step(capi.OP_DECISION, 330, 280, H("B4"), p_state=33, val_mps=0)
    b, R2, O2 = orc.binary_decision(33, 0, 330, 280)
    p2, v2 = orc.state_transition(33, 0, b)
    R3, O3, used = orc.renorm_d(H("B4"), R2, O2)
    assert (r["bin"], r["R"], r["O"], r["p_state"], r["val_mps"], r["bits_used"]) == (b, R3, O3, p2, v2, used)


# =========================================================================================== whole front end
def test_stream_decode_matches_oracle_pipeline(ctx, capi):
    b = hz.build_stream_cabac(200, 3000, n_active=64, n_ctx=64, slices_per_frame=8, frames_per_params=5)
    r = ctx.stream_decode(b["stream"], b["ops"], b["n_ops"], b["qp"], b["idc"], b["n_ctx"],
                          flags=capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE)
    onal, orbsp = orc.read_nal_units_arrays(b["stream"])
    assert r["scan"]["n_nals"] == len(onal["start"])
    sl = np.flatnonzero((onal["type"] == 1) | (onal["type"] == 5))
    assert np.array_equal(r["slice_nal"].astype(np.int64), sl)
    init = orc.ctx_init(b["qp"], b["idc"], b["n_ctx"])
    off = onal["rbsp_off"][sl].astype(np.uint64)
    length = onal["rbsp_len"][sl].astype(np.uint32)
    exp = oracle_decode_all(orbsp, off, length, b["ops"], b["n_ops"], init, orc.BYPASS_SPEC_OR, True)
    compare_cabac(capi, r["bins"], r["final"], None, exp, b["n_ops"], True)
    assert r["total_bins"] == int(b["n_ops"].sum()) + len(sl)
    for i in range(len(sl)):   # and the decoded bins are the ones the encoder coded
        nw = (int(b["n_ops"][i]) + 1) // 32
        assert np.array_equal(r["bins"][i][:nw], b["bins"][i, :nw])


def test_stream_submit_wait_jobs_in_flight(ctx, capi):
    """h264b_stream_submit / h264b_stream_wait: jobs overlap (three in flight), results stay per job and match the
    oracle; a fourth submit without a wait is refused; max_slices larger than the slices the stream holds is fine"""
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    jobs = [hz.build_stream_cabac(n, 1500 + 100 * k, n_active=64, n_ctx=64, slices_per_frame=4, frames_per_params=3,
                                  id_base=7000 * k) for k, n in enumerate([40, 9, 64, 1, 33, 17])]

    def check(b, r):
        onal, orbsp = orc.read_nal_units_arrays(b["stream"])
        assert r["scan"]["n_nals"] == len(onal["start"])
        assert np.array_equal(r["nals"]["start"].astype(np.int64), onal["start"])
        sl = np.flatnonzero((onal["type"] == 1) | (onal["type"] == 5))
        assert np.array_equal(r["slice_nal"].astype(np.int64), sl)
        init = orc.ctx_init(b["qp"], b["idc"], b["n_ctx"])
        exp = oracle_decode_all(orbsp, onal["rbsp_off"][sl].astype(np.uint64), onal["rbsp_len"][sl].astype(np.uint32),
                                b["ops"], b["n_ops"], init, orc.BYPASS_SPEC_OR, True)
        compare_cabac(capi, r["bins"], r["final"], None, exp, b["n_ops"], True)
        assert r["total_bins"] == int(b["n_ops"][:len(sl)].sum()) + len(sl)

    def submit(b, extra=0):
        qp = np.concatenate([b["qp"], np.full(extra, 26, np.int32)])
        idc = np.concatenate([b["idc"], np.zeros(extra, np.int32)])
        n_ops = np.concatenate([b["n_ops"], np.full(extra, 10, np.uint32)])
        return ctx.stream_submit(b["stream"], b["ops"], n_ops, qp, idc, b["n_ctx"], flags=flags)

    t0 = submit(jobs[0])
    t1 = submit(jobs[1], extra=5)
    t2 = submit(jobs[2])
    with pytest.raises(capi.H264BError):
        submit(jobs[3])
    check(jobs[0], ctx.stream_wait(*t0))
    t3 = submit(jobs[3], extra=2)
    check(jobs[1], ctx.stream_wait(*t1))
    t4 = submit(jobs[4])
    check(jobs[2], ctx.stream_wait(*t2))
    t5 = submit(jobs[5])
    check(jobs[3], ctx.stream_wait(*t3))
    check(jobs[4], ctx.stream_wait(*t4))
    check(jobs[5], ctx.stream_wait(*t5))
    with pytest.raises(capi.H264BError):
        ctx.stream_wait(*t5)
