"""ctypes binding of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package h264decode_b200.  Parity unpinned by the reference
(see oracle.h); pinned by SURVEY.md Appendix B vectors in tests/test_oracle_kat.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

OK, PANIC, CAPACITY = 0, 1, 2
TABLES_SPEC = 1
BYPASS_SPEC_OR = 2
OP_DECISION, OP_BYPASS, OP_TERMINATE = 0, 1, 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "ref_tables.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_NAL_FIELDS = [
    "NumBytes", "ForbiddenZeroBit", "RefIdc", "Type", "SvcExtensionFlag", "Avc3dExtensionFlag", "IdrFlag",
    "PriorityId", "NoInterLayerPredFlag", "DependencyId", "QualityId", "TemporalId", "UseRefBasePicFlag",
    "DiscardableFlag", "OutputFlag", "ReservedThree2Bits", "HeaderBytes", "NonIdrFlag", "ViewId", "AnchorPicFlag",
    "InterViewFlag", "ReservedOneBit", "ViewIdx", "DepthFlag", "EmulationPreventionThreeByte", "rbsp_len",
]


class NalUnit(C.Structure):
    _fields_ = [(n, C.c_int64) for n in _NAL_FIELDS]

    def as_dict(self):
        return {n: getattr(self, n) for n in _NAL_FIELDS}


class StreamNal(C.Structure):
    _fields_ = [("start_offset", C.c_int64), ("end_offset", C.c_int64), ("rbsp_off", C.c_int64), ("nal", NalUnit)]


class BitReader(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("len", C.c_int64), ("byteOffset", C.c_int64), ("bitOffset", C.c_int64),
                ("bitsRead", C.c_int64), ("panicked", C.c_int)]


class CabacFinal(C.Structure):
    _fields_ = [("codIRange", C.c_int64), ("codIOffset", C.c_int64), ("bitsRead", C.c_int64),
                ("flags", C.c_uint32), ("n_bins", C.c_uint32)]


_SPS_SCALARS = [
    "Profile", "Constraint0", "Constraint1", "Constraint2", "Constraint3", "Constraint4", "Constraint5", "Level", "ID",
    "ChromaFormat", "UseSeparateColorPlane", "BitDepthLumaMinus8", "BitDepthChromaMinus8",
    "QPrimeYZeroTransformBypass", "SeqScalingMatrixPresent", "Log2MaxFrameNumMinus4", "PicOrderCountType",
    "Log2MaxPicOrderCntLSBMin4", "DeltaPicOrderAlwaysZero", "OffsetForNonRefPic", "OffsetForTopToBottomField",
    "NumRefFramesInPicOrderCntCycle", "MaxNumRefFrames", "GapsInFrameNumValueAllowed", "PicWidthInMbsMinus1",
    "PicHeightInMapUnitsMinus1", "FrameMbsOnly", "MBAdaptiveFrameField", "Direct8x8Inference", "FrameCropping",
    "FrameCropLeftOffset", "FrameCropRightOffset", "FrameCropTopOffset", "FrameCropBottomOffset",
    "VuiParametersPresent", "AspectRatioInfoPresent", "AspectRatio", "SarWidth", "SarHeight", "OverscanInfoPresent",
    "OverscanAppropriate", "VideoSignalTypePresent", "VideoFormat", "VideoFullRange", "ColorDescriptionPresent",
    "ColorPrimaries", "TransferCharacteristics", "MatrixCoefficients", "ChromaLocInfoPresent",
    "ChromaSampleLocTypeTopField", "ChromaSampleLocTypeBottomField", "CpbCntMinus1", "BitRateScale", "CpbSizeScale",
    "InitialCpbRemovalDelayLengthMinus1", "CpbRemovalDelayLengthMinus1", "DpbOutputDelayLengthMinus1",
    "TimeOffsetLength", "TimingInfoPresent", "NumUnitsInTick", "TimeScale", "NalHrdParametersPresent",
    "FixedFrameRate", "VclHrdParametersPresent", "LowHrdDelay", "PicStructPresent", "BitstreamRestriction",
    "MotionVectorsOverPicBoundaries", "MaxBytesPerPicDenom", "MaxBitsPerMbDenom", "Log2MaxMvLengthHorizontal",
    "Log2MaxMvLengthVertical", "MaxDecFrameBuffering", "MaxNumReorderFrames",
]
MAX_LIST = 256


class SPS(C.Structure):
    _fields_ = ([(n, C.c_int64) for n in _SPS_SCALARS] + [
        ("n_SeqScalingList", C.c_int64), ("SeqScalingList", C.c_int64 * 12),
        ("n_OffsetForRefFrameList", C.c_int64), ("OffsetForRefFrameList", C.c_int64 * MAX_LIST),
        ("n_hrd", C.c_int64), ("BitRateValueMinus1", C.c_int64 * MAX_LIST), ("CpbSizeValueMinus1", C.c_int64 * MAX_LIST),
        ("Cbr", C.c_int64 * MAX_LIST), ("bits_read", C.c_int64)])

    def as_dict(self):
        d = {n: getattr(self, n) for n in _SPS_SCALARS}
        d["SeqScalingList"] = list(self.SeqScalingList[:self.n_SeqScalingList])
        nr, nh = min(self.n_OffsetForRefFrameList, MAX_LIST), min(self.n_hrd, MAX_LIST)
        d["n_OffsetForRefFrameList"], d["n_hrd"] = self.n_OffsetForRefFrameList, self.n_hrd  # len() of the Go lists
        d["OffsetForRefFrameList"] = list(self.OffsetForRefFrameList[:nr])                    # (first MAX_LIST kept)
        d["BitRateValueMinus1"] = list(self.BitRateValueMinus1[:nh])
        d["CpbSizeValueMinus1"] = list(self.CpbSizeValueMinus1[:nh])
        d["Cbr"] = list(self.Cbr[:nh])
        d["bits_read"] = self.bits_read
        return d


_PPS_SCALARS = [
    "ID", "SPSID", "EntropyCodingMode", "NumSliceGroupsMinus1", "BottomFieldPicOrderInFramePresent",
    "SliceGroupMapType", "SliceGroupChangeDirection", "SliceGroupChangeRateMinus1", "PicSizeInMapUnitsMinus1",
    "NumRefIdxL0DefaultActiveMinus1", "NumRefIdxL1DefaultActiveMinus1", "WeightedPred", "WeightedBipred",
    "PicInitQpMinus26", "PicInitQsMinus26", "ChromaQpIndexOffset", "DeblockingFilterControlPresent",
    "ConstrainedIntraPred", "RedundantPicCntPresent", "Transform8x8Mode", "PicScalingMatrixPresent",
    "SecondChromaQpIndexOffset", "bits_read",
]


class PPS(C.Structure):
    _fields_ = [(n, C.c_int64) for n in _PPS_SCALARS]

    def as_dict(self):
        return {n: getattr(self, n) for n in _PPS_SCALARS}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    u8p, i64, i64p = C.c_void_p, C.c_int64, C.POINTER(C.c_int64)
    L.orc_is_start_sequence.argtypes = [u8p, i64]
    L.orc_is_start_sequence.restype = C.c_int
    L.orc_new_nal_unit.argtypes = [u8p, i64, i64, C.POINTER(NalUnit), u8p]
    L.orc_new_nal_unit.restype = C.c_int
    L.orc_read_nal_units.argtypes = [u8p, i64, C.c_void_p, i64, u8p, i64, i64p, C.c_int]
    L.orc_read_nal_units.restype = i64
    L.orc_br_init.argtypes = [C.POINTER(BitReader), u8p, i64]
    L.orc_br_init.restype = None
    L.orc_br_next_field.argtypes = [C.POINTER(BitReader), i64]
    L.orc_br_next_field.restype = i64
    L.orc_br_read_one_bit.argtypes = [C.POINTER(BitReader)]
    L.orc_br_read_one_bit.restype = i64
    L.orc_br_golomb.argtypes = [C.POINTER(BitReader), i64p, i64]
    L.orc_br_golomb.restype = i64
    L.orc_ue.argtypes = [i64p, i64]
    L.orc_ue.restype = i64
    L.orc_se.argtypes = [i64p, i64]
    L.orc_se.restype = i64
    L.orc_init_decoding_engine.argtypes = [C.POINTER(BitReader), i64p, i64p]
    L.orc_init_decoding_engine.restype = None
    L.orc_binary_decision.argtypes = [C.c_uint32, i64, i64, i64p, i64p, i64p]
    L.orc_binary_decision.restype = None
    L.orc_state_transition.argtypes = [C.c_uint32, i64p, i64p, i64]
    L.orc_state_transition.restype = None
    L.orc_renorm_d.argtypes = [C.POINTER(BitReader), i64p, i64p]
    L.orc_renorm_d.restype = None
    L.orc_decode_bypass.argtypes = [C.c_uint32, C.POINTER(BitReader), i64, i64p, i64p]
    L.orc_decode_bypass.restype = None
    L.orc_decode_terminate.argtypes = [C.POINTER(BitReader), i64p, i64p, i64p]
    L.orc_decode_terminate.restype = None
    L.orc_cabac_decode_slice.argtypes = [C.c_uint32, u8p, i64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p,
                                         C.POINTER(CabacFinal)]
    L.orc_cabac_decode_slice.restype = C.c_int
    L.orc_decode_mb_types.argtypes = [C.c_uint32, C.c_int32, u8p, i64, i64, C.c_void_p, i64, C.c_void_p, i64p,
                                      C.POINTER(CabacFinal)]
    L.orc_decode_mb_types.restype = C.c_int
    L.orc_clip3.argtypes = [i64, i64, i64]
    L.orc_clip3.restype = i64
    L.orc_pre_ctx_state.argtypes = [i64, i64, i64]
    L.orc_pre_ctx_state.restype = i64
    L.orc_slice_qpy.argtypes = [i64, i64]
    L.orc_slice_qpy.restype = i64
    L.orc_mn.argtypes = [C.c_uint32, i64, i64, i64p, i64p]
    L.orc_mn.restype = None
    L.orc_ctx_state.argtypes = [i64]
    L.orc_ctx_state.restype = C.c_uint8
    L.orc_ctx_init.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, i64, i64, C.c_void_p]
    L.orc_ctx_init.restype = None
    L.orc_new_sps.argtypes = [u8p, i64, C.POINTER(SPS)]
    L.orc_new_sps.restype = C.c_int
    L.orc_new_pps.argtypes = [i64, u8p, i64, C.POINTER(PPS)]
    L.orc_new_pps.restype = C.c_int
    _lib = L
    return L


def _u8(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


# ------------------------------------------------------------------ NAL / RBSP
def new_nal_unit(frame, num_bytes_in_nal=None):
    """NewNalUnit(frame, numBytesInNal) -> (status, NalUnit dict, rbsp bytes)"""
    f = _u8(frame)
    n = len(f) if num_bytes_in_nal is None else num_bytes_in_nal
    out = NalUnit()
    rbsp = np.zeros(max(len(f), 1), dtype=np.uint8)
    st = lib().orc_new_nal_unit(f.ctypes.data, len(f), n, C.byref(out), rbsp.ctypes.data)
    return st, out.as_dict(), bytes(rbsp[:out.rbsp_len])


def read_nal_units(stream, literal=False, cap=None):
    """The handleConnection/readNalUnit loop.  Returns (count_or_negative_status, records, rbsp ndarray)."""
    s = _u8(stream)
    cap = cap if cap is not None else len(s) // 4 + 2
    recs = (StreamNal * cap)()
    rbsp = np.zeros(len(s) + 16, dtype=np.uint8)
    tot = C.c_int64(0)
    cnt = lib().orc_read_nal_units(s.ctypes.data, len(s), C.cast(recs, C.c_void_p), cap, rbsp.ctypes.data, len(rbsp),
                                   C.byref(tot), 1 if literal else 0)
    n = max(cnt, 0)
    return cnt, [recs[i] for i in range(n)], rbsp[:tot.value]


def read_nal_units_arrays(stream, literal=False, cap=None):
    """Same as read_nal_units but returns numpy arrays: dict(start, num_bytes, rbsp_off, rbsp_len, fzb, ref_idc,
    type, header_bytes) and the concatenated rbsp."""
    cnt, recs, rbsp = read_nal_units(stream, literal=literal, cap=cap)
    if cnt < 0:
        raise RuntimeError("oracle status %d" % cnt)
    arr = np.frombuffer(b"".join(bytes(r) for r in recs), dtype=np.int64).reshape(cnt, -1) if cnt else \
        np.zeros((0, 3 + len(_NAL_FIELDS)), dtype=np.int64)
    f = {n: i + 3 for i, n in enumerate(_NAL_FIELDS)}
    return {
        "start": arr[:, 0].copy(), "end": arr[:, 1].copy(), "rbsp_off": arr[:, 2].copy(),
        "num_bytes": arr[:, f["NumBytes"]].copy(), "rbsp_len": arr[:, f["rbsp_len"]].copy(),
        "fzb": arr[:, f["ForbiddenZeroBit"]].copy(), "ref_idc": arr[:, f["RefIdc"]].copy(),
        "type": arr[:, f["Type"]].copy(), "header_bytes": arr[:, f["HeaderBytes"]].copy(),
        "epb": arr[:, f["EmulationPreventionThreeByte"]].copy(), "fields": arr[:, 3:].copy(),
    }, rbsp


# ------------------------------------------------------------------ bit reader helpers
class Bits:
    """BitReader over a bytes object (keeps the buffer alive)."""

    def __init__(self, data):
        self.buf = _u8(data).copy()
        self.br = BitReader()
        lib().orc_br_init(C.byref(self.br), self.buf.ctypes.data, len(self.buf))

    def next_field(self, n):
        return lib().orc_br_next_field(C.byref(self.br), n)

    def one_bit(self):
        return lib().orc_br_read_one_bit(C.byref(self.br))

    def golomb(self):
        bits = (C.c_int64 * 130)()
        nb = lib().orc_br_golomb(C.byref(self.br), bits, 130)
        return bits, nb

    def ue(self):
        bits, nb = self.golomb()
        return lib().orc_ue(bits, nb)

    def se(self):
        bits, nb = self.golomb()
        return lib().orc_se(bits, nb)

    @property
    def panicked(self):
        return bool(self.br.panicked)

    @property
    def bits_read(self):
        return self.br.bitsRead


# ------------------------------------------------------------------ engine primitives
def binary_decision(pstate, valmps, R, O, flags=0):
    r, o, b = C.c_int64(R), C.c_int64(O), C.c_int64(0)
    lib().orc_binary_decision(flags, pstate, valmps, C.byref(r), C.byref(o), C.byref(b))
    return b.value, r.value, o.value


def state_transition(pstate, valmps, binval, flags=0):
    p, v = C.c_int64(pstate), C.c_int64(valmps)
    lib().orc_state_transition(flags, C.byref(p), C.byref(v), binval)
    return p.value, v.value


def renorm_d(data, R, O):
    bits = Bits(data)
    r, o = C.c_int64(R), C.c_int64(O)
    lib().orc_renorm_d(C.byref(bits.br), C.byref(r), C.byref(o))
    return r.value, o.value, bits.bits_read


def decode_bypass(data, R, O, flags=0):
    bits = Bits(data)
    o, b = C.c_int64(O), C.c_int64(0)
    lib().orc_decode_bypass(flags, C.byref(bits.br), R, C.byref(o), C.byref(b))
    return o.value, b.value


def decode_terminate(data, R, O):
    bits = Bits(data)
    r, o, b = C.c_int64(R), C.c_int64(O), C.c_int64(0)
    lib().orc_decode_terminate(C.byref(bits.br), C.byref(r), C.byref(o), C.byref(b))
    return r.value, o.value, b.value


def init_decoding_engine(data):
    bits = Bits(data)
    r, o = C.c_int64(0), C.c_int64(0)
    lib().orc_init_decoding_engine(C.byref(bits.br), C.byref(r), C.byref(o))
    return r.value, o.value


def make_op(kind, ctx=0):
    return (kind << 14) | (ctx & 0x3FF)


def cabac_decode_slice(data, ops, ctx_state, flags=0):
    """Returns (status, packed bins uint32[], final dict, ctx_state after)."""
    d = _u8(data)
    ops = np.ascontiguousarray(ops, dtype=np.uint16)
    st = np.ascontiguousarray(ctx_state, dtype=np.uint8).copy()
    bins = np.zeros((len(ops) + 31) // 32 + 1, dtype=np.uint32)
    fin = CabacFinal()
    rc = lib().orc_cabac_decode_slice(flags, d.ctypes.data, len(d), ops.ctypes.data, len(ops), st.ctypes.data, len(st),
                                      bins.ctypes.data, C.byref(fin))
    return rc, bins[:(len(ops) + 31) // 32], dict(codIRange=fin.codIRange, codIOffset=fin.codIOffset,
                                                  bitsRead=fin.bitsRead, flags=fin.flags, n_bins=fin.n_bins), st


def decode_mb_types(data, kind, n_mb, ctx_state, flags=0):
    """The mb_type walk (kind 0: I slice, 1: P / SP slice) -> (status, mb_types uint8[n_done], final dict, ctx_state after)"""
    d = _u8(data)
    st = np.ascontiguousarray(ctx_state, dtype=np.uint8).copy()
    out = np.zeros(max(n_mb, 1), dtype=np.uint8)
    fin = CabacFinal()
    nd = C.c_int64(0)
    rc = lib().orc_decode_mb_types(flags, kind, d.ctypes.data, len(d), n_mb, st.ctypes.data, len(st), out.ctypes.data,
                                   C.byref(nd), C.byref(fin))
    return rc, out[:nd.value], dict(codIRange=fin.codIRange, codIOffset=fin.codIOffset, bitsRead=fin.bitsRead,
                                    flags=fin.flags, n_bins=fin.n_bins), st


# ------------------------------------------------------------------ context init
def pre_ctx_state(m, n, qp):
    return lib().orc_pre_ctx_state(m, n, qp)


def ctx_state(pre):
    return lib().orc_ctx_state(pre)


def mn(ctx_idx, idc, flags=0):
    m, n = C.c_int64(0), C.c_int64(0)
    lib().orc_mn(flags, ctx_idx, idc, C.byref(m), C.byref(n))
    return m.value, n.value


def ctx_init(qp, idc, n_ctx, flags=0):
    qp = np.ascontiguousarray(qp, dtype=np.int32)
    idc = np.ascontiguousarray(idc, dtype=np.int32)
    out = np.zeros((len(qp), n_ctx), dtype=np.uint8)
    lib().orc_ctx_init(flags, qp.ctypes.data, idc.ctypes.data, len(qp), n_ctx, out.ctypes.data)
    return out


# ------------------------------------------------------------------ SPS / PPS
def new_sps(rbsp):
    r = _u8(rbsp)
    s = SPS()
    st = lib().orc_new_sps(r.ctypes.data, len(r), C.byref(s))
    return st, s.as_dict()


def new_pps(rbsp, sps_chroma_format=1):
    r = _u8(rbsp)
    p = PPS()
    st = lib().orc_new_pps(sps_chroma_format, r.ctypes.data, len(r), C.byref(p))
    return st, p.as_dict()


# ------------------------------------------------------------------ slice header (next row f1)
HANG = 2
_SLICE_HEADER_FIELDS = [
    "FirstMbInSlice", "SliceType", "PPSID", "ColorPlaneID", "FieldPic", "BottomField", "IDRPicID", "PicOrderCntLsb",
    "DeltaPicOrderCntBottom", "DeltaPicOrderCnt0", "DeltaPicOrderCnt1", "RedundantPicCnt", "DirectSpatialMvPred",
    "NumRefIdxActiveOverride", "NumRefIdxL0ActiveMinus1", "NumRefIdxL1ActiveMinus1", "RefPicListModificationFlagL0",
    "RefPicListModificationFlagL1", "ModificationOfPicNums", "AbsDiffPicNumMinus1", "LongTermPicNum",
    "LumaLog2WeightDenom", "ChromaLog2WeightDenom", "NLumaWeightL0", "NChromaWeightL0", "NLumaWeightL1",
    "NChromaWeightL1", "NoOutputOfPriorPicsFlag", "LongTermReferenceFlag", "AdaptiveRefPicMarkingModeFlag",
    "MemoryManagementControlOperation", "DifferenceOfPicNumsMinus1", "LongTermFrameIdx", "MaxLongTermFrameIdxPlus1",
    "CabacInit", "SliceQpDelta", "SpForSwitch", "SliceQsDelta", "DisableDeblockingFilter", "SliceAlphaC0OffsetDiv2",
    "SliceBetaOffsetDiv2", "SliceGroupChangeCycle", "ChromaArrayType", "SliceQPy", "bits_read"]


class SliceHeader(C.Structure):
    _fields_ = [(n, C.c_int64) for n in _SLICE_HEADER_FIELDS]


def new_slice_header(sps_fields, pps_fields, nal_type, nal_ref_idc, rbsp):
    """sps_fields / pps_fields: dicts of the orc_sps / orc_pps scalar names NewSliceContext reads.
    Returns (status, dict of header fields)."""
    sps, pps = SPS(), PPS()
    for k, v in sps_fields.items():      # (whole as_dict() results are fine: lists and counts are skipped)
        if k in _SPS_SCALARS:
            setattr(sps, k, int(v))
    for k, v in pps_fields.items():
        if k in _PPS_SCALARS:
            setattr(pps, k, int(v))
    d = _u8(rbsp)
    h = SliceHeader()
    L = lib()
    L.orc_new_slice_header.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    rc = L.orc_new_slice_header(C.byref(sps), C.byref(pps), int(nal_type), int(nal_ref_idc), d.ctypes.data, len(d),
                                C.byref(h))
    return rc, {n: getattr(h, n) for n in _SLICE_HEADER_FIELDS}


# ------------------------------------------------------------------ syntax-element glue (rows I5 / f3)
NA_CTX_ID = 10000
BINARIZATION_FIELDS = ["syntax_element", "prefix_suffix", "fixed_length", "unary", "truncated_unary", "cmax", "uegk",
                       "cmax_value", "max_is_prefix_suffix", "max_prefix", "max_suffix", "off_is_prefix_suffix",
                       "off_prefix", "off_suffix", "use_decode_bypass", "reserved"]


def ctx_idx(bin_idx, max_bin_idx_ctx, ctx_idx_offset):
    L = lib()
    L.orc_ctx_idx.restype = C.c_int64
    L.orc_ctx_idx.argtypes = [C.c_int64] * 3
    return L.orc_ctx_idx(int(bin_idx), int(max_bin_idx_ctx), int(ctx_idx_offset))


def new_binarization(se, st):
    out = (C.c_int32 * 16)()
    L = lib()
    L.orc_new_binarization.restype = None
    L.orc_new_binarization.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
    L.orc_new_binarization(int(se), int(st), out)
    return dict(zip(BINARIZATION_FIELDS, list(out)))


def init_cabac(bin_idx, max_prefix, off_prefix, pic_init_qp_minus26, slice_qp_delta, flags=0):
    L = lib()
    L.orc_init_cabac.restype = None
    L.orc_init_cabac.argtypes = [C.c_uint32] + [C.c_int64] * 5 + [C.POINTER(C.c_int64)] * 3
    p, v, c = C.c_int64(), C.c_int64(), C.c_int64()
    L.orc_init_cabac(flags, int(bin_idx), int(max_prefix), int(off_prefix), int(pic_init_qp_minus26), int(slice_qp_delta),
                     C.byref(p), C.byref(v), C.byref(c))
    return p.value, v.value, c.value


def mb_bin_string(st, mb_type, sub=False):
    bits = (C.c_int32 * 8)()
    L = lib()
    L.orc_mb_bin_string.restype = C.c_int32
    L.orc_mb_bin_string.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_void_p]
    n = L.orc_mb_bin_string(int(st), int(mb_type), int(bool(sub)), bits)
    return list(bits)[:n]


def bin_string_match(bin_string, bits):
    a = (C.c_int32 * max(len(bin_string), 1))(*bin_string)
    b = (C.c_int32 * max(len(bits), 1))(*bits)
    L = lib()
    L.orc_bin_string_match.restype = C.c_int32
    L.orc_bin_string_match.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    return L.orc_bin_string_match(a, len(bin_string), b, len(bits))
