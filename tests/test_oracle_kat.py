"""Pin the CPU oracle (oracle/oracle.c) against the hand-derived known-answer vectors of SURVEY.md Appendix B.

The reference ships no golden vectors and cannot be run (no Go toolchain), so these
vectors -- derived independently from the cited reference lines -- are the pins.  CPU only.
"""
import numpy as np
import pytest

from oracle import oracle as orc

H = bytes.fromhex


# ---------------------------------------------------------------- B.1  NewNalUnit (nalUnit.go:75-131)
@pytest.mark.parametrize("frame,typ,ref_idc,hdr,rbsp", [
    ("67 42 00 00 03 01 AA BB 00 00 00 01", 7, 3, 1, "42 00 00 01 AA BB 00 00"),
    ("00 00 03 07 00 00 03 00 00 00 01", 0, 0, 1, "00 03 07 00 00 00 00"),
    ("65 11 00 00 03", 5, 3, 1, "11 00 00"),
    ("65 11 22 00 00 03 44", 5, 3, 1, "11 22 00 00"),
    ("65 11 22 33", 5, 3, 1, "11"),
    ("6E 80 00 00 00 03 55 66 77 88", 14, 3, 4, "00 03 55 66"),
])
def test_b1_new_nal_unit(frame, typ, ref_idc, hdr, rbsp):
    st, nal, out = orc.new_nal_unit(H(frame))
    assert st == orc.OK
    assert (nal["Type"], nal["RefIdc"], nal["HeaderBytes"]) == (typ, ref_idc, hdr)
    assert out == H(rbsp)
    assert nal["NumBytes"] == len(H(frame))


def test_new_nal_unit_ext_headers():
    # type 20, svc_extension_flag=1: 3 ext bytes, fields per nalUnit.go:39-51
    st, nal, _ = orc.new_nal_unit(H("74 C5 A6 9F 11 22 33 44"))
    assert st == orc.OK and nal["Type"] == 20 and nal["HeaderBytes"] == 4
    # bits after the type: 1 | 1 | 000101 | 1 | 010 | 0110 | 100 | 1 | 1 | 1 | 11
    assert nal["SvcExtensionFlag"] == 1 and nal["IdrFlag"] == 1 and nal["PriorityId"] == 5
    assert nal["NoInterLayerPredFlag"] == 1 and nal["DependencyId"] == 2 and nal["QualityId"] == 6
    assert nal["TemporalId"] == 4 and nal["UseRefBasePicFlag"] == 1 and nal["DiscardableFlag"] == 1
    assert nal["OutputFlag"] == 1 and nal["ReservedThree2Bits"] == 3
    # type 21 with avc_3d_extension_flag=1: 2 ext bytes (nalUnit.go:53-61,95-97)
    st, nal, rbsp = orc.new_nal_unit(H("75 FF 80 AA BB CC DD"))
    assert st == orc.OK and nal["Type"] == 21 and nal["HeaderBytes"] == 3 and nal["Avc3dExtensionFlag"] == 1
    assert rbsp == H("AA BB")
    # type 21 with the flag clear: MVC extension, 3 ext bytes (nalUnit.go:98-101)
    st, nal, rbsp = orc.new_nal_unit(H("75 7F 80 AA BB CC DD"))
    assert nal["HeaderBytes"] == 4 and nal["Avc3dExtensionFlag"] == 0 and rbsp == H("BB")
    # type 14 with svc flag 0 -> MVC: non_idr 1, priority 6 bits, view_id 10 bits...
    st, nal, _ = orc.new_nal_unit(H("6E 40 00 01 99 88 77"))
    assert nal["SvcExtensionFlag"] == 0 and nal["NonIdrFlag"] == 1 and nal["HeaderBytes"] == 4


def test_new_nal_unit_panics_on_short_header():
    st, _, _ = orc.new_nal_unit(b"")
    assert st == orc.PANIC
    st, _, _ = orc.new_nal_unit(H("6E 80 00"))  # SVC header needs 4 bytes
    assert st == orc.PANIC


# ---------------------------------------------------------------- B.2  stream split (server.go:64-111)
B2_STREAM = H("FF 00 00 00 01 67 42 00 00 03 01 AA 00 00 00 01 68 CE 00 00 03 00 00 03 80 00 00 00 00 01 "
              "65 88 00 00 01 99 00 00 00 01 41 9A")


@pytest.mark.parametrize("literal", [False, True])
def test_b2_stream_split(literal):
    cnt, recs, rbsp = orc.read_nal_units(B2_STREAM, literal=literal)
    assert cnt == 3
    exp = [(5, 16, 11, 7, "42 00 00 01 AA 00 00"),
           (16, 30, 14, 8, "CE 00 00 00 00 80 00 00 00"),
           (30, 40, 10, 5, "88 00 00 01 99 00 00")]
    for r, (so, eo, nb, typ, rb) in zip(recs, exp):
        assert (r.start_offset, r.end_offset, r.nal.NumBytes, r.nal.Type) == (so, eo, nb, typ)
        assert bytes(rbsp[r.rbsp_off:r.rbsp_off + r.nal.rbsp_len]) == H(rb)


def test_stream_edge_cases():
    assert orc.read_nal_units(b"")[0] == 0
    assert orc.read_nal_units(H("00 00 00 01"))[0] == 0              # one start code: no complete NAL
    assert orc.read_nal_units(H("00 00 00 01 65 88"))[0] == 0        # the last NAL is never emitted
    cnt, recs, rbsp = orc.read_nal_units(H("00 00 00 01 00 00 00 01"))  # back-to-back start codes: 4-byte NAL
    assert cnt == 1 and recs[0].nal.NumBytes == 4 and recs[0].nal.Type == 0 and recs[0].nal.rbsp_len == 1
    assert bytes(rbsp) == H("00")
    # 00 00 00 00 01: the extra zero stays with the previous NAL
    cnt, recs, _ = orc.read_nal_units(H("00 00 00 01 41 00 00 00 00 01 00 00 00 01"))
    assert cnt == 2 and recs[0].nal.NumBytes == 6 and recs[1].nal.NumBytes == 4


# ---------------------------------------------------------------- B.3  parameter-set template
def test_b3_sps_pps_template():
    stream = H("00000001 67 640028ACD94078022640 00000001 68 EE0F2C8B 00000001 65 8884 00000001")
    cnt, recs, rbsp = orc.read_nal_units(stream)
    assert cnt == 3 and [r.nal.Type for r in recs] == [7, 8, 5]
    sps_rbsp = bytes(rbsp[recs[0].rbsp_off:recs[0].rbsp_off + recs[0].nal.rbsp_len])
    assert sps_rbsp == H("640028ACD940780226400000")
    st, sps = orc.new_sps(sps_rbsp)
    assert st == orc.OK
    exp = dict(Profile=100, Level=40, ID=0, ChromaFormat=1, BitDepthLumaMinus8=0, BitDepthChromaMinus8=0,
               Log2MaxFrameNumMinus4=0, PicOrderCountType=0, Log2MaxPicOrderCntLSBMin4=2, MaxNumRefFrames=4,
               PicWidthInMbsMinus1=119, PicHeightInMapUnitsMinus1=67, FrameMbsOnly=1, Direct8x8Inference=1,
               FrameCropping=0, VuiParametersPresent=0)
    for k, v in exp.items():
        assert sps[k] == v, k
    pps_rbsp = bytes(rbsp[recs[1].rbsp_off:recs[1].rbsp_off + recs[1].nal.rbsp_len])
    assert pps_rbsp == H("EE0F2C8B0000")
    st, pps = orc.new_pps(pps_rbsp, sps["ChromaFormat"])
    assert st == orc.OK
    exp = dict(ID=0, SPSID=0, EntropyCodingMode=1, NumSliceGroupsMinus1=0, PicInitQpMinus26=-3, PicInitQsMinus26=0,
               ChromaQpIndexOffset=-2, DeblockingFilterControlPresent=1, Transform8x8Mode=1,
               PicScalingMatrixPresent=0)
    for k, v in exp.items():
        assert pps[k] == v, k
    # the PPS without the High-profile tail makes the reference panic in MoreRBSPData (A11)
    st, _ = orc.new_pps(H("EE0F2C80") + b"\x00\x00", 1)
    assert st == orc.PANIC


def test_pps_panic_branches():
    # pic_scaling_matrix_present -> write into a nil slice (pps.go:103)
    st, _ = orc.new_pps(H("EE0F2CC0") + b"\x80\x00\x00", 1)
    assert st == orc.PANIC


# ---------------------------------------------------------------- B.4  Exp-Golomb (bit_reader.go:62-64,158-161)
def _eg_bits(code_num):
    v = code_num + 1
    nb = v.bit_length()
    return "0" * (nb - 1) + format(v, "b")


@pytest.mark.parametrize("k,ue,se", [(0, 0, 0), (1, 1, 0), (2, 2, -1), (3, 3, 1), (4, 4, -2), (5, 5, 2), (6, 6, -3),
                                     (7, 7, 3)])
def test_b4_exp_golomb(k, ue, se):
    bits = _eg_bits(k) + "1"
    bits += "0" * (-len(bits) % 8)
    data = int(bits, 2).to_bytes(len(bits) // 8, "big")
    assert orc.Bits(data).ue() == ue
    assert orc.Bits(data).se() == se


def test_bitreader_overrun_panics():
    b = orc.Bits(b"\xA5")
    assert b.next_field(8) == 0xA5 and not b.panicked    # reading exactly to the end is fine
    b.next_field(1)
    assert b.panicked
    b = orc.Bits(b"\x00\x00")                             # golomb with no terminating 1
    b.ue()
    assert b.panicked


# ---------------------------------------------------------------- B.5  context init (cabac.go:118-121,158-164)
@pytest.mark.parametrize("m,n,qp,pre,ps,mps", [
    (20, -15, 0, 1, 62, 0), (20, -15, 26, 17, 46, 0), (20, -15, 51, 48, 15, 0), (-28, 127, 26, 81, 17, 1),
    (-4, 127, 51, 114, 50, 1), (-39, 127, 51, 2, 61, 0), (0, 0, 33, 1, 62, 0), (57, 2, 51, 126, 62, 1),
    (-7, 24, 26, 12, 51, 0),
])
def test_b5_ctx_init(m, n, qp, pre, ps, mps):
    assert orc.pre_ctx_state(m, n, qp) == pre
    st = orc.ctx_state(pre)
    assert (st & 63, st >> 6) == (ps, mps)


def test_mn_lookup_rules():
    assert orc.mn(0, -1) == (20, -15) and orc.mn(0, 0) == (0, 0)       # ctx 0..10 only have the key -1 (A14)
    assert orc.mn(11, 0) == (23, 33) and orc.mn(11, -1) == (0, 0)
    assert orc.mn(30, 0) == (-4, 127) and orc.mn(30, 0, orc.TABLES_SPEC) == (-46, 127)   # A3
    assert orc.mn(76, -1) == (-7, 24) and orc.mn(76, 7) == (-7, 24)    # I/SI column for any idc outside 0..2
    assert orc.mn(92, 1) == (-36, 127)
    assert orc.mn(40, 0) == (0, 0) and orc.mn(69, 1) == (0, 0) and orc.mn(105, 2) == (0, 0)
    assert orc.mn(11, 3) == (0, 0)
    out = orc.ctx_init([26, 51], [0, -1], 128)
    assert out.shape == (2, 128)
    assert out[0, 40] == 62 and out[1, 500 % 128] in range(128)
    m, n = orc.mn(76, -1)
    assert out[1, 76] == orc.ctx_state(orc.pre_ctx_state(m, n, 51))


# ---------------------------------------------------------------- B.6  engine primitives (REF tables)
@pytest.mark.parametrize("p,v,R,O,exp", [
    (33, 0, 330, 200, (0, 269, 200)), (33, 0, 330, 280, (1, 61, 11)), (33, 1, 400, 332, (1, 333, 332)),
    (33, 1, 400, 333, (0, 67, 0)), (0, 0, 510, 269, (0, 270, 269)), (0, 0, 510, 270, (1, 240, 0)),
    (62, 1, 256, 249, (1, 250, 249)), (62, 1, 256, 250, (0, 6, 0)), (63, 0, 300, 297, (0, 298, 297)),
    (63, 0, 300, 298, (1, 2, 0)),
])
def test_b6_binary_decision(p, v, R, O, exp):
    assert orc.binary_decision(p, v, R, O) == exp


@pytest.mark.parametrize("p,v,b,exp", [
    (0, 0, 1, (0, 1)), (0, 1, 0, (0, 0)), (0, 0, 0, (1, 0)), (59, 1, 1, (61, 1)), (59, 1, 0, (37, 1)),
    (62, 0, 0, (62, 0)), (63, 1, 1, (63, 1)), (63, 1, 0, (63, 1)), (33, 0, 1, (25, 0)),
])
def test_b6_state_transition(p, v, b, exp):
    assert orc.state_transition(p, v, b) == exp


def test_b6_spec_tables_differ_only_where_documented():
    assert orc.binary_decision(33, 0, 330, 280, orc.TABLES_SPEC) == (0, 299, 280)   # LPS width 31, not 61 (A1)
    assert orc.state_transition(59, 1, 1, orc.TABLES_SPEC) == (60, 1)


RENORM_BITS = H("B4 00 00")  # 1011 0100 0...


@pytest.mark.parametrize("R,O,exp", [(256, 10, (256, 10)), (255, 10, (510, 21)), (128, 5, (256, 11)),
                                     (6, 3, (384, 237)), (2, 1, (256, 218))])
def test_b6_renorm(R, O, exp):
    r, o, nbits = orc.renorm_d(RENORM_BITS, R, O)
    assert (r, o) == exp
    k = 0
    while (R << k) < 256:
        k += 1
    assert nbits == k


@pytest.mark.parametrize("R,O,bit,ref,spec", [
    (510, 100, 0, (200, 0), (200, 0)), (510, 100, 1, (400, 0), (201, 0)), (300, 200, 0, (100, 1), (100, 1)),
    (300, 200, 1, (500, 1), (101, 1)), (256, 255, 1, (764, 1), (255, 1)),
])
def test_b6_bypass(R, O, bit, ref, spec):
    data = bytes([0x80 if bit else 0x00])
    assert orc.decode_bypass(data, R, O) == ref
    assert orc.decode_bypass(data, R, O, orc.BYPASS_SPEC_OR) == spec


def test_bypass_ref_shift_wraps_like_go_int64():
    o = (1 << 62) + 5
    got, b = orc.decode_bypass(b"\x80", 300, o)   # (o << 1) << 1 wraps to 20 in 64-bit two's complement
    assert (got, b) == (20, 0)
    got, b = orc.decode_bypass(b"\x00", 300, o)   # o << 1 = -2^63 + 10 : negative, so signed compare says < R
    assert (got, b) == (-(1 << 63) + 10, 0)


@pytest.mark.parametrize("R,O,exp", [(510, 507, (508, 507, 0)), (510, 508, (508, 508, 1)), (258, 100, (256, 100, 0)),
                                     (256, 253, (508, 507, 0)), (256, 254, (254, 254, 1))])
def test_b6_terminate(R, O, exp):
    assert orc.decode_terminate(b"\x80", R, O) == exp


def test_b6_init_engine():
    assert orc.init_decoding_engine(H("A5 C3")) == (510, 0b101001011)


def test_decode_slice_panics_past_end():
    ops = np.array([orc.make_op(orc.OP_BYPASS)] * 40, dtype=np.uint16)
    rc, bins, fin, _ = orc.cabac_decode_slice(H("00 00 00"), ops, np.zeros(4, np.uint8), orc.BYPASS_SPEC_OR)
    assert rc == orc.PANIC and fin["flags"] == 1 and fin["n_bins"] == 24 - 9
    rc, bins, fin, _ = orc.cabac_decode_slice(H("00"), ops, np.zeros(4, np.uint8))
    assert rc == orc.PANIC and fin["n_bins"] == 0
