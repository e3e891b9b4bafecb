"""Parameter sets (rows S1 / f4: NewSPS h264/sps.go:192-437, NewPPS h264/pps.go:40-133): the product's walk
(param_sets.cuh) against the oracle's literal restatement -- on the CPU through the emulation library (same source as
the kernels), on the GPU through the C ABI, and inside the stream job (H264B_STREAM_PARAM_SETS: SPS / PPS NAL units ->
active parameter sets per slice -> slice headers -> CABAC, nothing leaving the device in between).  Inputs: parameter
sets written by a bit writer that follows what the REFERENCE reads (every branch: scaling lists, POC types, cropping,
VUI with both HRDs, slice groups, the 8x8 tail), the Appendix-B.3 template, and random bytes (the walks must agree on
garbage too, including where the reference would panic)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from tests.test_slice_header import BitWriter, write_header, compare as compare_header, SPS_KEYS, PPS_KEYS

SPS_MAX_HRD, SPS_MAX_REF = 64, 256


def H(s):
    return bytes.fromhex(s.replace(" ", ""))


def _se_code(rng, lim=40):
    return int(rng.integers(0, lim))


def write_scaling_list(w, rng, size):
    """se() deltas until next_scale hits 0 or the list ends (what scalingList consumes)"""
    last = nxt = 8
    for _ in range(size):
        if nxt != 0:
            k = _se_code(rng, 12)
            w.ue(k)
            delta = (-1) ** (k + 1) * (k // 2)
            nxt = (last + delta + 256) % 256
        last = last if nxt == 0 else nxt


def write_hrd(w, rng):
    cpb = int(rng.integers(0, 4))
    w.ue(cpb)
    w.u(4, int(rng.integers(0, 16)))
    w.u(4, int(rng.integers(0, 16)))
    for _ in range(cpb + 1):
        w.ue(int(rng.integers(0, 100000)))
        w.ue(int(rng.integers(0, 100000)))
        w.u(1, int(rng.integers(0, 2)))
        for _ in range(4):
            w.u(5, int(rng.integers(0, 32)))


def write_sps(rng, panic_ok=False):
    """an SPS the reference's walk accepts; returns (rbsp bytes, bits written)"""
    w = BitWriter()
    profile = int(rng.choice([66, 77, 88, 100, 110, 122, 244, 44, 83, 86, 118, 128, 138, 139, 134, 135]))
    w.u(8, profile)
    w.u(8, int(rng.integers(0, 256)) & 0xFC)      # constraint flags + reserved_zero_2bits
    w.u(8, int(rng.integers(9, 52)))
    w.ue(int(rng.integers(0, 32)))
    chroma = int(rng.integers(0, 4))
    w.ue(chroma)                                   # read for every profile
    if profile in (100, 110, 122, 244, 44, 83, 86, 118, 128, 138, 139, 134, 135):
        if chroma == 3:
            w.u(1, int(rng.integers(0, 2)))
        w.ue(int(rng.integers(0, 7)))
        w.ue(int(rng.integers(0, 7)))
        w.u(1, int(rng.integers(0, 2)))
        sm = int(rng.random() < 0.5)
        w.u(1, sm)
        if sm:
            for i in range(8 if chroma != 3 else 12):
                ok_index = (i < 2) or (6 <= i < 8)      # other present lists index past the default matrices
                present = int(rng.random() < 0.6) if (ok_index or panic_ok) else 0
                w.u(1, present)
                if present:
                    if not ok_index:
                        return w.bytes(), None
                    write_scaling_list(w, rng, 16 if i < 6 else 64)
    w.ue(int(rng.integers(0, 13)))
    poc = int(rng.integers(0, 3))
    w.ue(poc)
    if poc == 0:
        w.ue(int(rng.integers(0, 13)))
    elif poc == 1:
        w.u(1, int(rng.integers(0, 2)))
        w.ue(_se_code(rng))
        w.ue(_se_code(rng))
        n = int(rng.integers(0, 6)) if rng.random() < 0.9 else int(rng.integers(250, 300))
        w.ue(n)
        for _ in range(n):
            w.ue(_se_code(rng, 2000))
    w.ue(int(rng.integers(0, 17)))
    w.u(1, int(rng.integers(0, 2)))
    w.ue(int(rng.integers(0, 512)))
    w.ue(int(rng.integers(0, 512)))
    fmo = int(rng.integers(0, 2))
    w.u(1, fmo)
    if not fmo:
        w.u(1, int(rng.integers(0, 2)))
    w.u(1, int(rng.integers(0, 2)))
    crop = int(rng.integers(0, 2))
    w.u(1, crop)
    if crop:
        for _ in range(4):
            w.ue(int(rng.integers(0, 64)))
    vui = int(rng.random() < 0.6)
    w.u(1, vui)
    if vui:
        ar = int(rng.integers(0, 2))
        w.u(1, ar)
        if ar:
            w.u(8, int(rng.choice([1, 14, 255])))       # 255 = Extended_SAR in the standard; the reference tests 999
        ov = int(rng.integers(0, 2))
        w.u(1, ov)
        if ov:
            w.u(1, int(rng.integers(0, 2)))
        vs = int(rng.integers(0, 2))
        w.u(1, vs)
        if vs:
            w.u(3, int(rng.integers(0, 8)))
            w.u(1, int(rng.integers(0, 2)))
            cd = int(rng.integers(0, 2))
            w.u(1, cd)
            if cd:
                for _ in range(3):
                    w.u(8, int(rng.integers(0, 256)))
        cl = int(rng.integers(0, 2))
        w.u(1, cl)
        if cl:
            w.ue(int(rng.integers(0, 6)))
            w.ue(int(rng.integers(0, 6)))
        ti = int(rng.integers(0, 2))
        w.u(1, ti)
        if ti:
            w.u(32, int(rng.integers(0, 1 << 32)))
            w.u(32, int(rng.integers(0, 1 << 32)))
            w.u(1, int(rng.integers(0, 2)))
        nal_hrd = int(rng.integers(0, 2))
        w.u(1, nal_hrd)
        if nal_hrd:
            write_hrd(w, rng)
        vcl_hrd = int(rng.integers(0, 2))
        w.u(1, vcl_hrd)
        if vcl_hrd:
            write_hrd(w, rng)
        if nal_hrd or vcl_hrd:
            w.u(1, int(rng.integers(0, 2)))
        w.u(1, int(rng.integers(0, 2)))
        br = int(rng.integers(0, 2))
        w.u(1, br)
        if br:
            w.u(1, int(rng.integers(0, 2)))
            for _ in range(6):
                w.ue(int(rng.integers(0, 17)))
    return w.bytes(), len(w.bits)


def write_pps(rng, entropy=None, qp_minus26_code=None, simple=False, raw=False):
    """a PPS the reference's walk accepts (slice-group map types 3..5 only, no scaling matrix, 8x8 tail present)"""
    w = BitWriter()
    w.ue(int(rng.integers(0, 256)))
    w.ue(int(rng.integers(0, 32)))
    w.u(1, int(rng.integers(0, 2)) if entropy is None else entropy)
    w.u(1, int(rng.integers(0, 2)))
    groups = 0 if simple else int(rng.integers(0, 3))
    w.ue(groups)
    if groups > 0:
        w.ue(int(rng.integers(3, 6)))
        w.u(1, int(rng.integers(0, 2)))
        w.ue(int(rng.integers(1, 40)))
    w.ue(int(rng.integers(0, 32)))
    w.ue(int(rng.integers(0, 32)))
    w.u(1, 0 if simple else int(rng.integers(0, 2)))
    w.u(2, 0 if simple else int(rng.integers(0, 3)))
    w.ue(_se_code(rng, 50) if qp_minus26_code is None else qp_minus26_code)
    w.ue(_se_code(rng, 50))
    w.ue(_se_code(rng, 24))
    w.u(1, int(rng.integers(0, 2)))
    w.u(1, int(rng.integers(0, 2)))
    w.u(1, 0 if simple else int(rng.integers(0, 2)))
    w.u(1, int(rng.integers(0, 2)))    # transform_8x8_mode
    w.u(1, 0)                           # pic_scaling_matrix_present (1 panics)
    zeros = int(rng.integers(0, 9))     # MoreRBSPData then eats zeros and the stop bit
    w.bits += [0] * zeros
    if raw:
        return w
    return w.bytes(), len(w.bits) + 1


def garbage(rng, k):
    n = int(rng.integers(0, 48))
    d = rng.integers(0, 256, n).astype(np.uint8)
    d[rng.random(n) < [0.0, 0.3, 0.8][k % 3]] = 0
    if k % 5 == 0 and n:
        d[0] = int(rng.choice([100, 110, 244]))     # a High-type profile: reaches the scaling-list branch
    return d


def compare_sps(got, exp_rc, exp, what):
    from h264decode_b200 import capi
    assert int(got["status"]) == exp_rc, (what, int(got["status"]), exp_rc)
    if exp_rc != orc.OK:
        return
    for mine, theirs in zip(capi.SPS_SCALARS, orc._SPS_SCALARS):
        assert int(got[mine]) == exp[theirs], (what, mine)
    assert int(got["bits_read"]) == exp["bits_read"], what
    ns = int(got["n_seq_scaling_list"])
    assert list(got["seq_scaling_list"][:ns]) == exp["SeqScalingList"], what
    nr, nh = int(got["n_offset_for_ref_frame"]), int(got["n_hrd"])
    assert nr == exp["n_OffsetForRefFrameList"] and nh == exp["n_hrd"], what
    assert list(got["offset_for_ref_frame"][:min(nr, SPS_MAX_REF)]) == exp["OffsetForRefFrameList"][:SPS_MAX_REF], what
    for mine, theirs in (("bit_rate_value_minus1", "BitRateValueMinus1"), ("cpb_size_value_minus1", "CpbSizeValueMinus1"),
                         ("cbr", "Cbr")):
        assert list(got[mine][:min(nh, SPS_MAX_HRD)]) == exp[theirs][:SPS_MAX_HRD], (what, mine)


def compare_pps(got, exp_rc, exp, what):
    from h264decode_b200 import capi
    assert int(got["status"]) == exp_rc, (what, int(got["status"]), exp_rc)
    if exp_rc != orc.OK:
        return
    for mine, theirs in zip(capi.PPS_SCALARS, orc._PPS_SCALARS):
        assert int(got[mine]) == exp[theirs], (what, mine)
    assert int(got["bits_read"]) == exp["bits_read"], what


def sps_cases(seed, n_written, n_garbage):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_written):
        rbsp, nbits = write_sps(rng, panic_ok=rng.random() < 0.15)
        out.append((np.frombuffer(rbsp, np.uint8), nbits))
    out += [(garbage(rng, k), None) for k in range(n_garbage)]
    return out


def pps_cases(seed, n_written, n_garbage):
    rng = np.random.default_rng(1000 + seed)
    out = []
    for _ in range(n_written):
        rbsp, nbits = write_pps(rng)
        out.append((np.frombuffer(rbsp, np.uint8), nbits))
    out += [(garbage(rng, k), None) for k in range(n_garbage)]
    return out


# ------------------------------------------------------------------------------------------------ CPU: emulation
@pytest.fixture(scope="module")
def emul():
    from tests import test_hd_logic
    import os
    import subprocess
    deps = [test_hd_logic.SRC] + [os.path.join(test_hd_logic.HERE, "..", "h264decode_b200", "csrc", f)
                                  for f in ("annexb_local.cuh", "cabac_lane.cuh", "slice_header.cuh", "param_sets.cuh",
                                            "tables.inc")] + [os.path.join(test_hd_logic.HERE, "..", "include", "h264b200.h")]
    out = test_hd_logic.OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-x", "c++", test_hd_logic.SRC,
                               "-o", out])
    L = C.CDLL(out)
    for f in (L.emul_parse_sps, L.emul_parse_pps):
        f.restype = None
        f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.emul_make_param_sets.restype = None
    L.emul_make_param_sets.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def emul_parse(fn, dtype, rbsp):
    out = np.zeros(1, dtype)
    buf = np.ascontiguousarray(rbsp, dtype=np.uint8)
    fn(buf.ctypes.data if len(buf) else None, len(buf), out.ctypes.data)
    return out[0]


@pytest.mark.parametrize("seed", range(3))
def test_sps_walk_matches_oracle_cpu(emul, seed):
    from h264decode_b200 import capi
    n_ok = n_panic = n_hrd = n_ref = 0
    for rbsp, nbits in sps_cases(seed, 250, 500):
        rc, exp = orc.new_sps(rbsp)
        got = emul_parse(emul.emul_parse_sps, capi.SPS_DTYPE, rbsp)
        compare_sps(got, rc, exp, (seed, rbsp.tobytes().hex()))
        if nbits is not None:
            assert rc == orc.OK and exp["bits_read"] == nbits   # the writer and the walk agree on the extent
        n_ok += rc == orc.OK
        n_panic += rc == orc.PANIC
        n_hrd += rc == orc.OK and exp["n_hrd"] > 0
        n_ref += rc == orc.OK and exp["n_OffsetForRefFrameList"] > SPS_MAX_REF
    assert n_ok > 200 and n_panic > 100 and n_hrd > 30 and n_ref > 0, (n_ok, n_panic, n_hrd, n_ref)


@pytest.mark.parametrize("seed", range(3))
def test_pps_walk_matches_oracle_cpu(emul, seed):
    from h264decode_b200 import capi
    n_ok = n_panic = 0
    for rbsp, nbits in pps_cases(seed, 250, 600):
        rc, exp = orc.new_pps(rbsp)
        got = emul_parse(emul.emul_parse_pps, capi.PPS_DTYPE, rbsp)
        compare_pps(got, rc, exp, (seed, rbsp.tobytes().hex()))
        if nbits is not None:
            assert rc == orc.OK and exp["bits_read"] == nbits
        n_ok += rc == orc.OK
        n_panic += rc == orc.PANIC
    assert n_ok > 250 and n_panic > 100, (n_ok, n_panic)


def test_parameter_set_known_answers(emul):
    """SURVEY.md Appendix B.3 (hand-derived) through the product's walk, plus the reference's panic branches"""
    from h264decode_b200 import capi
    sps = emul_parse(emul.emul_parse_sps, capi.SPS_DTYPE, np.frombuffer(H("640028ACD940780226400000"), np.uint8))
    exp = dict(profile=100, level=40, id=0, chroma_format=1, bit_depth_luma_minus8=0, bit_depth_chroma_minus8=0,
               log2_max_frame_num_minus4=0, pic_order_count_type=0, log2_max_pic_order_cnt_lsb_min4=2, max_num_ref_frames=4,
               pic_width_in_mbs_minus1=119, pic_height_in_map_units_minus1=67, frame_mbs_only=1, direct_8x8_inference=1,
               frame_cropping=0, vui_parameters_present=0, status=0)
    for k, v in exp.items():
        assert int(sps[k]) == v, k
    pps = emul_parse(emul.emul_parse_pps, capi.PPS_DTYPE, np.frombuffer(H("EE0F2C8B0000"), np.uint8))
    exp = dict(id=0, sps_id=0, entropy_coding_mode=1, num_slice_groups_minus1=0, pic_init_qp_minus26=-3, pic_init_qs_minus26=0,
               chroma_qp_index_offset=-2, deblocking_filter_control_present=1, transform_8x8_mode=1,
               pic_scaling_matrix_present=0, status=0)
    for k, v in exp.items():
        assert int(pps[k]) == v, k
    # without the High-profile tail MoreRBSPData runs off the end (A11); pic_scaling_matrix_present writes a nil slice
    for bad in (H("EE0F2C80") + b"\x00\x00", H("EE0F2CC0") + b"\x80\x00\x00"):
        assert int(emul_parse(emul.emul_parse_pps, capi.PPS_DTYPE, np.frombuffer(bad, np.uint8))["status"]) == orc.PANIC
    # slice_group_map_type 0 / 2 / 6 write through nil slices (pps.go:61,65,74)
    for t in (0, 2, 6):
        w = BitWriter()
        w.ue(0), w.ue(0), w.u(1, 1), w.u(1, 0), w.ue(1), w.ue(t), w.ue(3), w.ue(3)
        rb = np.frombuffer(w.bytes(), np.uint8)
        rc, _ = orc.new_pps(rb)
        assert rc == orc.PANIC and int(emul_parse(emul.emul_parse_pps, capi.PPS_DTYPE, rb)["status"]) == orc.PANIC
    # the fields the slice-header walk reads
    ps = capi.ParamSets()
    a, b = np.array([sps]), np.array([pps])
    emul.emul_make_param_sets(a.ctypes.data, b.ctypes.data, C.byref(ps))
    assert (ps.chroma_format, ps.frame_mbs_only, ps.log2_max_pic_order_cnt_lsb_min4, ps.entropy_coding_mode,
            ps.pic_init_qp_minus26, ps.deblocking_filter_control_present) == (1, 1, 2, 1, -3, 1)


def test_golomb_code_words_longer_than_64_bits(emul):
    """ue() of a code word with 64 or more leading zeros wraps modulo 2^64 (bitVal shifts past the word), se() goes
    through float64: both walks agree, and PPS slice_group_map_type 6 with a size that wrapped to -1 does not panic"""
    from h264decode_b200 import capi
    for zeros, suffix in ((64, 0), (64, 5), (70, (1 << 64) + 12345), (63, (1 << 62) + 1)):
        w = BitWriter()
        w.ue(0), w.ue(0), w.u(1, 1), w.u(1, 0), w.ue(1), w.ue(6)
        w.bits += [0] * zeros + [1]
        w.u(zeros, suffix & ((1 << zeros) - 1))
        for _ in range(2):
            w.ue(1)
        w.u(1, 0), w.u(2, 0)
        w.bits += [0] * 66 + [1] + [1] * 65 + [0]          # se() of a huge odd code word
        w.ue(3), w.ue(4), w.u(3, 0), w.u(2, 0)
        rb = np.frombuffer(w.bytes(), np.uint8)
        rc, exp = orc.new_pps(rb)
        got = emul_parse(emul.emul_parse_pps, capi.PPS_DTYPE, rb)
        compare_pps(got, rc, exp, (zeros, suffix))
        assert (rc == orc.OK) == (exp["PicSizeInMapUnitsMinus1"] < 0)


# ------------------------------------------------------------------------------------------------ GPU: the C ABI
def _pack(cases):
    """all RBSPs in one buffer, ragged and unaligned"""
    off, ln, pos = [], [], 0
    for rbsp, _ in cases:
        off.append(pos)
        ln.append(len(rbsp))
        pos += len(rbsp) + len(off) % 3
    data = np.zeros(pos + 8, np.uint8)
    for o, (rbsp, _) in zip(off, cases):
        data[o:o + len(rbsp)] = rbsp
    return data, off, ln


@pytest.mark.gpu
def test_parse_sps_pps_gpu_match_oracle():
    from h264decode_b200 import capi
    ctx = capi.Context(0)
    try:
        sc = sps_cases(7, 300, 700)
        data, off, ln = _pack(sc)
        got = ctx.parse_sps(data, off, ln)
        n_ok = 0
        for i, (rbsp, _) in enumerate(sc):
            rc, exp = orc.new_sps(rbsp)
            compare_sps(got[i], rc, exp, i)
            n_ok += rc == orc.OK
        assert n_ok > 250
        pc = pps_cases(7, 300, 700)
        data, off, ln = _pack(pc)
        got = ctx.parse_pps(data, off, ln)
        n_ok = 0
        for i, (rbsp, _) in enumerate(pc):
            rc, exp = orc.new_pps(rbsp)
            compare_pps(got[i], rc, exp, i)
            n_ok += rc == orc.OK
        assert n_ok > 300
        assert len(ctx.parse_sps(np.zeros(0, np.uint8), [], [])) == 0
    finally:
        ctx.close()


def _ps_dict(sps, pps):
    d = {k: sps[v] for k, v in SPS_KEYS.items()}
    d.update({k: pps[v] for k, v in PPS_KEYS.items()})
    return d


@pytest.mark.gpu
def test_stream_pipeline_takes_parameter_sets_from_the_stream():
    """H264B_STREAM_PARAM_SETS: several (SPS, PPS) generations in one stream, a PPS-only update, slices before any
    parameter set, slices between an SPS and its PPS, and a PPS the reference cannot parse.  Every slice's header must be
    walked with the sets handleConnection would hold at that point (server.go:147-162), and the CABAC stage must decode
    the slices the oracle can decode, with SliceQPY / cabac_init_idc from the stream."""
    import harness as hz
    from h264decode_b200 import capi
    rng = np.random.default_rng(77)
    n_active = n_ctx = 64
    ops = hz.gen_schedule(3, 1200, n_active)
    SC = b"\x00\x00\x00\x01"
    parts, plan = [], []                        # plan: per slice (sps rbsp or None, pps rbsp or None, expected decode)
    cur_sps = cur_pps = None
    events = (["slice"] * 2 + ["sps", "slice", "pps", "slice", "slice", "pps", "slice", "sps", "pps"] + ["slice"] * 5 +
              ["badpps", "slice", "pps", "slice", "sps", "slice", "pps"] + ["slice"] * 9)
    slices = []
    for ev in events:
        if ev == "sps":
            while True:
                rb, nb = write_sps(rng)
                st, f = orc.new_sps(np.frombuffer(rb, np.uint8))
                if nb is not None and st == orc.OK and len(hz.escape(np.frombuffer(rb, np.uint8))) == len(rb):
                    break
            cur_sps, cur_pps = rb, None
            parts += [SC, b"\x67", rb]
        elif ev in ("pps", "badpps"):
            rb = write_pps(rng, entropy=1)[0] if ev == "pps" else H("EE0F2CC0") + b"\x80"
            rb = hz.escape(np.frombuffer(rb, np.uint8)).tobytes()
            if cur_sps is not None:
                cur_pps = rb
            else:
                cur_pps = None
            parts += [SC, b"\x68", rb]
        else:
            slices.append((cur_sps, cur_pps, len(parts)))
            parts += [SC, None, None]           # filled below, once the slice's parameters are known
    n = len(slices)
    n_ops = rng.integers(40, 1200, n).astype(np.uint32)
    slice_types = rng.integers(0, 10, n)
    qp_delta_code = rng.integers(0, 30, n)
    idc_hdr = rng.integers(0, 3, n)
    qp = np.zeros(n, np.int32)
    idc = np.zeros(n, np.int32)
    decodable = np.zeros(n, bool)
    hdr_bytes = []
    for s, (sps_rb, pps_rb, _) in enumerate(slices):
        ps = None
        if sps_rb is not None and pps_rb is not None:
            st1, f1 = orc.new_sps(np.frombuffer(sps_rb, np.uint8))
            # (the stream carries the escaped PPS; the walk sees the stripped one plus the two bytes readNalUnit leaves)
            st2, f2 = orc.new_pps(np.frombuffer(_unescape(pps_rb), np.uint8))
            if st1 == orc.OK and st2 == orc.OK:
                ps = _ps_dict(f1, f2)
        nal_type, ref_idc = (5, 3) if s % 5 == 0 else (1, int(rng.integers(0, 4)))
        if ps is None or ps["slice_group_change_rate_minus1"] == 0 and ps["num_slice_groups_minus1"] > 0:
            hdr_bytes.append((nal_type, ref_idc, bytes(rng.integers(1, 255, 12).astype(np.uint8))))
            continue
        dq = (-1) ** (int(qp_delta_code[s]) + 1) * (int(qp_delta_code[s]) // 2)
        w = write_header(rng, ps, nal_type, ref_idc, int(slice_types[s]), qp_delta=dq, cabac_init_idc=int(idc_hdr[s]),
                         raw=True)
        bits = w.bits + [1] * (-len(w.bits) % 8)
        hdr_bytes.append((nal_type, ref_idc, bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))))
        qp[s] = 26 + ps["pic_init_qp_minus26"] + dq
        idc[s] = -1 if slice_types[s] % 5 in (2, 4) else idc_hdr[s]
        decodable[s] = True
    g = hz.gen_cabac_slices(3, ops, n_ops, n_active, n_ctx, qp, idc)
    for s, (_, _, at) in enumerate(slices):
        nal_type, ref_idc, hb = hdr_bytes[s]
        body = np.concatenate([np.frombuffer(hb, np.uint8), g["data"][s, :g["lens"][s]]])
        parts[at + 1] = bytes([(ref_idc << 5) | nal_type])
        parts[at + 2] = hz.escape(body).tobytes()
    stream = np.frombuffer(b"".join(parts) + SC, np.uint8)

    ctx = capi.Context(0)
    try:
        flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE | capi.STREAM_PARAM_SETS
        t = ctx.stream_submit(stream, ops, n_ops, None, None, n_ctx, flags=flags, max_slices=n + 2, max_sps=8, max_pps=8)
        r = ctx.stream_wait(*t)
        # bounds that fall short are raised to what the run reported and the job runs again: same result
        r2 = ctx.stream_wait(*ctx.stream_submit(stream, ops, n_ops, None, None, n_ctx, flags=flags, max_slices=n + 2,
                                                max_sps=2, max_pps=1))
        assert np.array_equal(r2["sps"], r["sps"]) and np.array_equal(r2["pps"], r["pps"])
        assert np.array_equal(r2["slice_sps"], r["slice_sps"]) and np.array_equal(r2["slice_pps"], r["slice_pps"])
        assert np.array_equal(r2["final"], r["final"]) and np.array_equal(r2["bins_flat"], r["bins_flat"])
    finally:
        ctx.close()
    onal, orbsp = orc.read_nal_units_arrays(stream)
    rb_of = lambda k: orbsp[int(onal["rbsp_off"][k]):int(onal["rbsp_off"][k]) + int(onal["rbsp_len"][k])]
    i_sps, i_pps = np.flatnonzero(onal["type"] == 7), np.flatnonzero(onal["type"] == 8)
    assert np.array_equal(r["sps_nal"], i_sps) and np.array_equal(r["pps_nal"], i_pps)
    for j, k in enumerate(i_sps):
        compare_sps(r["sps"][j], *orc.new_sps(rb_of(k)), ("sps", j))
    for j, k in enumerate(i_pps):
        compare_pps(r["pps"][j], *orc.new_pps(rb_of(k)), ("pps", j))
    sl = np.flatnonzero((onal["type"] == 1) | (onal["type"] == 5))
    assert len(sl) == n == len(r["final"])
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    n_dec = 0
    for s, k in enumerate(sl):
        bs, bp = i_sps[i_sps < k], i_pps[i_pps < k]
        e_sps = len(bs) - 1
        e_pps = len(bp) - 1 if (len(bs) and len(bp) and bp[-1] > bs[-1]) else -1
        assert (int(r["slice_sps"][s]), int(r["slice_pps"][s])) == (e_sps, e_pps), s
        hdr = r["headers"][s]
        usable = False
        if e_sps >= 0 and e_pps >= 0:
            st1, f1 = orc.new_sps(rb_of(bs[-1]))
            st2, f2 = orc.new_pps(rb_of(bp[-1]))
            usable = st1 == orc.OK and st2 == orc.OK
        if not usable:
            assert int(hdr["status"]) == capi.SH_PANIC and r["final"]["flags"][s] & capi.F_OVERRUN, s
            assert not decodable[s]
            continue
        rc, h = orc.new_slice_header(f1, f2, int(onal["type"][k]), int(onal["ref_idc"][k]), rb_of(k))
        compare_header(hdr, int(hdr["status"]), rc, h, s)
        if rc != orc.OK:
            continue
        assert decodable[s] and h["SliceQPy"] == qp[s]
        skip = (h["bits_read"] + 7) // 8
        init = orc.ctx_init(np.array([h["SliceQPy"]], np.int32), np.array([idc[s]], np.int32), n_ctx)[0]
        rc, bins, fin, _ = orc.cabac_decode_slice(rb_of(k)[skip:], np.concatenate([ops[:n_ops[s]], term]), init,
                                                  orc.BYPASS_SPEC_OR)
        nw = (int(n_ops[s]) + 1) // 32
        assert np.array_equal(r["bins"][s][:nw], bins[:nw]) and np.array_equal(r["bins"][s][:nw], g["bins"][s, :nw]), s
        assert (r["final"]["cod_i_range"][s], r["final"]["cod_i_offset"][s], r["final"]["bits_read"][s]) == (
            fin["codIRange"], fin["codIOffset"], fin["bitsRead"]), s
        n_dec += 1
    assert n_dec >= 10 and n_dec == int(decodable.sum()), (n_dec, decodable.sum())


def _unescape(b):
    """what NewNalUnit leaves of an escaped payload that is followed by a start code: EPBs removed + 00 00 (A8)"""
    out = bytearray()
    z = 0
    for x in b:
        if z >= 2 and x == 3:
            z = 0
            continue
        out.append(x)
        z = z + 1 if x == 0 else 0
    return bytes(out) + b"\x00\x00"


@pytest.mark.gpu
def test_c1_stream_parameter_sets_on_the_device():
    """BASELINE configs[0]: the 1 MB stream's SPS / PPS fields from the device parse == the oracle's"""
    import harness as hz
    from h264decode_b200 import capi
    s = hz.build_stream_c1(1 << 20)
    ctx = capi.Context(0)
    try:
        _, nals, _, rbsp = ctx.annexb_scan(s)
        k7, k8 = np.flatnonzero(nals["type"] == 7), np.flatnonzero(nals["type"] == 8)
        assert len(k7) == 1 and len(k8) == 1
        sps = ctx.parse_sps(rbsp, nals["rbsp_off"][k7], nals["rbsp_len"][k7])[0]
        pps = ctx.parse_pps(rbsp, nals["rbsp_off"][k8], nals["rbsp_len"][k8])[0]
    finally:
        ctx.close()
    onal, orbsp = orc.read_nal_units_arrays(s)
    rb = lambda k: orbsp[int(onal["rbsp_off"][k]):int(onal["rbsp_off"][k]) + int(onal["rbsp_len"][k])]
    compare_sps(sps, *orc.new_sps(rb(k7[0])), "c1 sps")
    compare_pps(pps, *orc.new_pps(rb(k8[0])), "c1 pps")
    assert int(sps["profile"]) == 100 and int(sps["pic_width_in_mbs_minus1"]) == 119 and int(pps["pic_init_qp_minus26"]) == -3


@pytest.mark.gpu
def test_parameter_sets_carry_across_batches():
    """Batched ingest: a stream cut at any NAL boundary into two jobs, the second one inheriting the sets in force
    (job.initial_sps / initial_pps), yields the slice headers of the uncut stream -- including a PPS that arrives in the
    second batch for an SPS of the first, and slices with no parameter sets at all.  (No CABAC ops: headers only.)"""
    import harness as hz
    from h264decode_b200 import capi
    rng = np.random.default_rng(99)
    SC = b"\x00\x00\x00\x01"
    nal_list, ps = [], None
    cur_sps = cur_pps = None
    for ev in ["slice", "sps", "slice", "pps", "slice", "slice", "pps", "slice", "sps", "pps", "slice", "slice", "slice"]:
        if ev == "sps":
            while True:
                rb, nb = write_sps(rng)
                if nb is not None and orc.new_sps(np.frombuffer(rb, np.uint8))[0] == orc.OK:
                    break
            cur_sps, cur_pps = orc.new_sps(np.frombuffer(rb, np.uint8))[1], None
            nal_list.append(b"\x67" + hz.escape(np.frombuffer(rb, np.uint8)).tobytes())
        elif ev == "pps":
            rb = write_pps(rng, entropy=1)[0]
            esc = hz.escape(np.frombuffer(rb, np.uint8)).tobytes()
            if cur_sps is not None:
                cur_pps = orc.new_pps(np.frombuffer(_unescape(esc), np.uint8))[1]
            nal_list.append(b"\x68" + esc)
        else:
            if cur_sps is not None and cur_pps is not None:
                p = _ps_dict(cur_sps, cur_pps)
                hb, _ = write_header(rng, p, 1, 2, int(rng.integers(0, 10)))
            else:
                hb = bytes(rng.integers(1, 255, 9).astype(np.uint8))
            nal_list.append(b"\x41" + hz.escape(np.frombuffer(hb, np.uint8)).tobytes())
    whole = np.frombuffer(b"".join(SC + x for x in nal_list) + SC, np.uint8)
    empty_ops = np.zeros(0, np.uint16)
    flags = capi.STREAM_PARAM_SETS

    def run(ctx, stream, **kw):
        return ctx.stream_wait(*ctx.stream_submit(stream, empty_ops, None, None, None, 64, flags=flags, max_slices=32,
                                                  max_sps=8, max_pps=8, **kw))

    ctx = capi.Context(0)
    try:
        ref = run(ctx, whole)
        n_slices_all = len(ref["headers"])
        assert n_slices_all == 8 and int(ref["headers"]["status"][0]) == capi.SH_PANIC
        assert (ref["headers"]["status"] == capi.SH_OK).sum() >= 6
        n_inherited = 0
        for cut in range(1, len(nal_list)):
            a = np.frombuffer(b"".join(SC + x for x in nal_list[:cut]) + SC, np.uint8)
            b = np.frombuffer(b"".join(SC + x for x in nal_list[cut:]) + SC, np.uint8)
            r1 = run(ctx, a)
            init_sps = init_pps = None
            if len(r1["sps"]):
                init_sps = r1["sps"][-1].copy()
                later = np.flatnonzero(r1["pps_nal"] > r1["sps_nal"][-1])
                if len(later):
                    init_pps = r1["pps"][later[-1]].copy()
            r2 = run(ctx, b, initial_sps=init_sps, initial_pps=init_pps)
            got = np.concatenate([r1["headers"], r2["headers"]])
            assert len(got) == n_slices_all, cut
            assert np.array_equal(got["status"], ref["headers"]["status"]), cut
            ok = got["status"] == capi.SH_OK
            for name in capi.SLICE_HEADER_DTYPE.names:
                if name not in ("reserved",):
                    assert np.array_equal(got[name][ok], ref["headers"][name][ok]), (cut, name)
            n_inherited += int((r2["slice_sps"] == -2).sum())
            assert np.all((r2["slice_pps"] != -2) | (r2["slice_sps"] == -2))
        assert n_inherited > 10
    finally:
        ctx.close()
