"""The C++ host layer (host/h264.hpp: the reference's Go API re-stated over the C ABI) driven by
tests/native/host_test.cpp.  CPU: it builds, links against the library and refuses to run without a GPU.  GPU: NAL
units (whole buffer, batched ingest through a pipe, single frames), context init and the per-call engine methods
against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as hz
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "host_test.cpp")
EXE = os.path.join(HERE, "native", "_build", "host_test")


@pytest.fixture(scope="module")
def exe():
    from h264decode_b200 import build
    build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "host", "h264.hpp"), os.path.join(ROOT, "include", "h264b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", SRC, "-o", EXE, "-L" + os.path.join(ROOT, "h264decode_b200"),
                               "-lh264b200", "-Wl,-rpath,$ORIGIN/../../../h264decode_b200", "-lpthread"])
    return EXE


def run(exe, *args):
    p = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    return p.returncode, [l.split() for l in p.stdout.splitlines()]


def fnv(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_host_layer_builds_and_has_no_cpu_path(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, out = run(exe, "ctx", 1, 2, 3)
    assert rc == 1 and out[0][0] == "error" and "status" in out[0]


def expect_nal_lines(stream):
    nal, rbsp = orc.read_nal_units_arrays(stream)
    f = {name: i for i, name in enumerate(orc._NAL_FIELDS)}
    lines = []
    for k in range(len(nal["start"])):
        rb = rbsp[int(nal["rbsp_off"][k]):int(nal["rbsp_off"][k]) + int(nal["rbsp_len"][k])]
        fl = nal["fields"][k]
        lines.append(["nal", str(int(nal["start"][k])), str(int(nal["num_bytes"][k])), str(int(nal["fzb"][k])),
                      str(int(nal["ref_idc"][k])), str(int(nal["type"][k])), str(int(nal["header_bytes"][k])), str(len(rb)),
                      "%016x" % fnv(rb), str(int(nal["epb"][k])), str(int(fl[f["SvcExtensionFlag"]])),
                      str(int(fl[f["Avc3dExtensionFlag"]])), str(int(fl[f["PriorityId"]])), str(int(fl[f["ViewId"]])),
                      str(int(fl[f["TemporalId"]])), str(int(fl[f["ViewIdx"]]))])
    return lines


def make_stream(tmp_path, n=300000, seed=3):
    from tests.test_hd_logic import random_stream
    rng = np.random.default_rng(seed)
    parts = [hz.build_stream_c1(n // 2), random_stream(rng, n // 4, 0.3, 0.003, ext_types=True),
             hz.build_stream_cabac(24, 1500, slices_per_frame=4, frames_per_params=2)["stream"]]
    s = np.concatenate(parts)
    path = os.path.join(str(tmp_path), "stream.bin")
    s.tofile(path)
    return s, path


@pytest.mark.gpu
def test_read_nal_units_whole_buffer(exe, tmp_path):
    s, path = make_stream(tmp_path)
    rc, out = run(exe, "nals", path)
    assert rc == 0
    assert out == expect_nal_lines(s)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,chunk,room", [(65536, 4000, 1 << 20), (10007, 333, 4096), (1 << 20, 65536, 1 << 20),
                                              (3001, 17, 64)])
def test_batched_ingest_through_a_pipe(exe, tmp_path, batch, chunk, room):
    """handleConnection as modified: the same NAL units, same absolute offsets, whatever the batch / read sizes;
    NAL units larger than a batch and than the carry room included"""
    s, path = make_stream(tmp_path, n=120000 if batch < 20000 else 300000)
    rc, out = run(exe, "ingest", path, batch, chunk, room)
    assert rc == 0, out[-3:]
    exp = expect_nal_lines(s)
    assert out[-1] == ["units", str(len(exp))]
    assert out[:-1] == exp


@pytest.mark.gpu
def test_new_nal_unit_single_frames(exe, tmp_path):
    for i, hexs in enumerate(["67 42 00 00 03 01 AA BB 00 00 00 01", "65 11 22 00 00 03 44", "6E 80 00 00 00 03 55 66 77 88",
                              "75 FF 80 AA BB CC DD", "65", "74 C5 A6"]):
        f = bytes.fromhex(hexs)
        path = os.path.join(str(tmp_path), "f%d.bin" % i)
        open(path, "wb").write(f)
        rc, out = run(exe, "frame", path)
        st, o, rb = orc.new_nal_unit(f)
        if st == orc.PANIC:
            assert out == [["panic"]]
            continue
        assert rc == 0
        assert out[0][2:9] == [str(o["NumBytes"]), str(o["ForbiddenZeroBit"]), str(o["RefIdc"]), str(o["Type"]),
                               str(o["HeaderBytes"]), str(len(rb)), "%016x" % fnv(rb)]


@pytest.mark.gpu
def test_ctx_functions(exe):
    triples = [(20, -15, 0), (-28, 127, 26), (57, 2, 51), (0, 0, 99), (-4, 127, -5)]
    rc, out = run(exe, "ctx", *[x for t in triples for x in t])
    assert rc == 0
    pre = [l for l in out if l[0] == "pre"]
    assert [int(l[1]) for l in pre] == [orc.pre_ctx_state(*t) for t in triples]
    for l in out:
        if l[0] == "mn":
            assert (int(l[3]), int(l[4])) == orc.mn(int(l[1]), int(l[2])), l
    init = [l for l in out if l[0] == "init"][0]
    st = orc.ctx_init(np.array([26, 0, 51, 30], np.int32), np.array([0, -1, 2, 1], np.int32), 128)
    assert init[1] == "%016x" % fnv(st.tobytes()) and init[2:] == ["28", "51"]


@pytest.mark.gpu
def test_engine_methods_per_call(exe, tmp_path):
    """InitDecodingEngine / DecodeDecision / DecodeBypass (REF form) / DecodeTerminate / BinaryDecision /
    StateTransitionProcess, one call at a time on a shared BitReader, like the Go methods"""
    rng = np.random.default_rng(12)
    data = rng.integers(0, 256, 64).astype(np.uint8)
    path = os.path.join(str(tmp_path), "bits.bin")
    data.tofile(path)
    rc, out = run(exe, "engine", path)
    assert rc == 0, out
    L = orc.lib()
    bits = orc.Bits(data)
    R, O = C.c_int64(0), C.c_int64(0)
    L.orc_init_decoding_engine(C.byref(bits.br), C.byref(R), C.byref(O))
    assert out[0] == ["init", str(R.value), str(O.value), str(bits.bits_read)]
    state = np.array([20 | (1 << 6)], np.uint8)
    for i in range(24):
        b = C.c_int64(0)
        if i % 5 == 3:
            L.orc_decode_bypass(0, C.byref(bits.br), R.value, C.byref(O), C.byref(b))
        elif i % 11 == 10:
            L.orc_decode_terminate(C.byref(bits.br), C.byref(R), C.byref(O), C.byref(b))
        else:
            L.orc_decode_decision(0, C.byref(bits.br), state.ctypes.data, C.byref(R), C.byref(O), C.byref(b))
        assert out[1 + i] == ["step", str(i), str(b.value), str(R.value), str(O.value), str(int(state[0] & 63)),
                              str(int(state[0] >> 6)), str(bits.bits_read)], i
    bv, r2, o2 = orc.binary_decision(int(state[0] & 63), int(state[0] >> 6), R.value, O.value)
    assert out[25] == ["core", str(bv), str(r2), str(o2)]
    p2, v2 = orc.state_transition(int(state[0] & 63), int(state[0] >> 6), 1 - int(state[0] >> 6))
    assert out[26] == ["trans", str(p2), str(v2)]


@pytest.mark.gpu
def test_parameter_sets_and_slice_headers_like_handle_connection(exe, tmp_path):
    """h264::NewSPS / NewPPS / SliceHeaders driven the way handleConnection dispatches NAL units (server.go:145-162)"""
    from tests.test_param_sets import write_sps, write_pps, H
    from tests.test_slice_header import write_header, SPS_KEYS, PPS_KEYS
    import harness as hz
    rng = np.random.default_rng(5)
    SC = b"\x00\x00\x00\x01"
    parts, exp = [], []
    ps = None
    for ev in ["sps", "pps", "slice", "slice", "badpps", "slice", "pps", "slice", "sps", "pps", "slice", "slice"]:
        if ev == "sps":
            rb = write_sps(rng)[0]
            parts += [SC, b"\x67", hz.escape(np.frombuffer(rb, np.uint8)).tobytes()]
        elif ev in ("pps", "badpps"):
            rb = write_pps(rng, entropy=1)[0] if ev == "pps" else H("EE0F2CC0") + b"\x80"
            parts += [SC, b"\x68", hz.escape(np.frombuffer(rb, np.uint8)).tobytes()]
        else:
            # the header is written against the parameter sets in force (parsed back by the oracle below)
            onal, orbsp = orc.read_nal_units_arrays(np.frombuffer(b"".join(parts) + SC, np.uint8))
            rb_of = lambda k: orbsp[int(onal["rbsp_off"][k]):int(onal["rbsp_off"][k]) + int(onal["rbsp_len"][k])]
            k7, k8 = np.flatnonzero(onal["type"] == 7), np.flatnonzero(onal["type"] == 8)
            st1, f1 = orc.new_sps(rb_of(k7[-1]))
            st2, f2 = orc.new_pps(rb_of(k8[-1]))
            if st1 == orc.OK and st2 == orc.OK and k8[-1] > k7[-1]:
                ps = {k: f1[v] for k, v in SPS_KEYS.items()}
                ps.update({k: f2[v] for k, v in PPS_KEYS.items()})
                hb, _ = write_header(rng, ps, 1, 2, int(rng.integers(0, 10)))
            else:
                hb = bytes(rng.integers(1, 255, 10).astype(np.uint8))
            parts += [SC, b"\x41", hz.escape(np.frombuffer(hb, np.uint8)).tobytes()]
    stream = np.frombuffer(b"".join(parts) + SC, np.uint8)
    path = os.path.join(str(tmp_path), "psets.bin")
    stream.tofile(path)
    rc, out = run(exe, "psets", path)
    assert rc == 0, out
    onal, orbsp = orc.read_nal_units_arrays(stream)
    sps = pps = None
    for k in range(len(onal["start"])):
        rb = orbsp[int(onal["rbsp_off"][k]):int(onal["rbsp_off"][k]) + int(onal["rbsp_len"][k])]
        t = int(onal["type"][k])
        if t == 7:
            st, f = orc.new_sps(rb)
            sps, pps = (f if st == orc.OK else None), None
            exp.append(["sps"] + [str(f[n]) for n in ("Profile", "Level", "PicWidthInMbsMinus1", "PicHeightInMapUnitsMinus1",
                                                      "PicOrderCountType", "n_hrd", "bits_read")] if st == orc.OK else ["panic", "7"])
        elif t == 8:
            st, f = orc.new_pps(rb)
            pps = f if (st == orc.OK and sps is not None) else None
            exp.append(["pps"] + [str(f[n]) for n in ("ID", "EntropyCodingMode", "PicInitQpMinus26", "ChromaQpIndexOffset",
                                                      "Transform8x8Mode", "bits_read")] if st == orc.OK else ["panic", "8"])
        elif t in (1, 5) and sps is not None and pps is not None:
            st, h = orc.new_slice_header(sps, pps, t, int(onal["ref_idc"][k]), rb)
            exp.append(["slice", str(h["SliceType"]), str(h["SliceQPy"]), str(h["CabacInit"]), str(h["bits_read"])]
                       if st == orc.OK else ["panic", str(t)])
    assert out == exp
    assert sum(l[0] == "slice" for l in out) >= 4 and ["panic", "8"] in out


@pytest.mark.gpu
def test_glue_functions_per_call(exe):
    """h264::CtxIdx / NewBinarization / InitCabac, one call at a time like the Go functions they mirror"""
    rc, out = run(exe, "glue")
    assert rc == 0, out
    it = iter(out)
    for off in (3, 17, 21, 69, 276):
        for b in (-1, 0, 1, 2, 5, 9):
            assert next(it) == ["ctxidx", str(b), str(off), str(orc.ctx_idx(b, 6, off))]
    for se in range(15):
        for st in (0, 2, 4, 1):
            bz = orc.new_binarization(se, st)
            p, v, _ = orc.init_cabac(0, bz["max_prefix"], bz["off_prefix"], -3, 5)
            assert next(it) == ["bin", str(se), str(st), str(bz["prefix_suffix"]), str(bz["max_prefix"]),
                                str(bz["off_prefix"]), str(bz["use_decode_bypass"]), str(p), str(v)]
    bz = orc.new_binarization(5, 0)
    p, v, _ = orc.init_cabac(2, bz["max_prefix"], bz["off_prefix"], 4, -9)
    assert next(it) == ["init", str(p), str(v)]


@pytest.mark.gpu
@pytest.mark.parametrize("batch,chunk", [(1 << 20, 65536), (700, 97), (257, 31)])
def test_batched_ingest_with_handle_connection_dispatch(exe, tmp_path, batch, chunk):
    """ByteStreamReader::Run with IngestHandlers: NewSPS / NewPPS / slice headers come out of the same device job as the
    split, and the parameter sets in force carry from batch to batch -- whatever the batch size, the lines are those of
    one pass over the whole stream"""
    from tests.test_param_sets import write_sps, write_pps, H, _unescape
    from tests.test_slice_header import write_header, SPS_KEYS, PPS_KEYS
    rng = np.random.default_rng(8)
    SC = b"\x00\x00\x00\x01"
    parts, exp = [], []
    sps = pps = None
    for ev in ["slice", "sps", "slice", "pps", "slice", "slice", "badpps", "slice", "pps", "slice", "sps", "pps"] + ["slice"] * 6:
        if ev == "sps":
            while True:
                rb, nb = write_sps(rng)
                st, f = orc.new_sps(np.frombuffer(rb, np.uint8))
                if nb is not None and st == orc.OK:
                    break
            sps, pps = f, None
            parts += [SC, b"\x67", hz.escape(np.frombuffer(rb, np.uint8)).tobytes()]
            exp.append(["sps"] + [str(f[n]) for n in ("Profile", "Level", "PicWidthInMbsMinus1", "PicHeightInMapUnitsMinus1",
                                                      "PicOrderCountType", "n_hrd", "bits_read")])
        elif ev in ("pps", "badpps"):
            rb = write_pps(rng, entropy=1)[0] if ev == "pps" else H("EE0F2CC0") + b"\x80"
            esc = hz.escape(np.frombuffer(rb, np.uint8)).tobytes()
            st, f = orc.new_pps(np.frombuffer(_unescape(esc), np.uint8))
            pps = f if (st == orc.OK and sps is not None) else None
            parts += [SC, b"\x68", esc]
            exp.append(["pps"] + [str(f[n]) for n in ("ID", "EntropyCodingMode", "PicInitQpMinus26", "ChromaQpIndexOffset",
                                                      "Transform8x8Mode", "bits_read")] if st == orc.OK else ["panic", "8"])
        else:
            if sps is not None and pps is not None:
                ps = {k: sps[v] for k, v in SPS_KEYS.items()}
                ps.update({k: pps[v] for k, v in PPS_KEYS.items()})
                hb, _ = write_header(rng, ps, 1, 2, int(rng.integers(0, 10)))
                esc = hz.escape(np.frombuffer(hb, np.uint8)).tobytes()
                st, h = orc.new_slice_header(sps, pps, 1, 2, np.frombuffer(_unescape(esc), np.uint8))
                exp.append(["slice", str(h["SliceType"]), str(h["SliceQPy"]), str(h["CabacInit"]), str(h["bits_read"])]
                           if st == orc.OK else ["panic", "1"])
            else:
                esc = bytes(rng.integers(1, 255, 10).astype(np.uint8))
                exp.append(["panic", "1"])
            parts += [SC, b"\x41", esc]
    stream = np.frombuffer(b"".join(parts) + SC, np.uint8)
    path = os.path.join(str(tmp_path), "ingest_psets.bin")
    stream.tofile(path)
    rc, out = run(exe, "ingest_psets", path, batch, chunk)
    assert rc == 0, out[-3:]
    assert out[-1] == ["units", str(len(exp))]
    assert out[:-1] == exp
    assert sum(l[0] == "slice" for l in exp) >= 8


def test_cut_byte_ranges_matches_the_python_twin(exe, tmp_path):
    """host/h264.hpp CutByteRanges == h264decode_b200.sharding.cut_byte_ranges (pure host logic: runs without a GPU)"""
    from h264decode_b200 import sharding
    rng = np.random.default_rng(5)
    for t in range(12):
        n = int(rng.integers(0, 40000))
        s = rng.integers(0, 256, n, dtype=np.uint8)
        s[rng.random(n) < 0.4] = 0
        for pos in rng.integers(0, max(1, n - 4), max(1, n // 700)) if n >= 4 else []:
            s[pos:pos + 4] = [0, 0, 0, 1]
        path = os.path.join(str(tmp_path), "r%d.bin" % t)
        s.tofile(path)
        for n_ranges in (1, 2, 7):
            rc, out = run(exe, "ranges", path, n_ranges)
            assert rc == 0
            assert [(int(a), int(b)) for _, a, b in out] == sharding.cut_byte_ranges(s, n_ranges)


@pytest.mark.gpu
@pytest.mark.parametrize("n_ranges", [2, 5])
def test_read_nal_units_by_byte_ranges(exe, tmp_path, n_ranges):
    s, path = make_stream(tmp_path)
    rc, out = run(exe, "nals_ranges", path, n_ranges)
    assert rc == 0
    assert out == expect_nal_lines(s)


@pytest.mark.gpu
def test_scheduler_wrapper_runs_a_batch_of_streams(exe, tmp_path):
    """h264::Scheduler (host/h264.hpp) over h264b_scheduler_run: a batch of streams of different sizes over two workers;
    per stream the oracle's NAL unit count, per slice the bins the test encoder coded and the oracle's final engine state."""
    import struct
    from h264decode_b200 import capi
    rng = np.random.default_rng(11)
    n_streams = 9
    mean_bins = (1024 * 2 ** (5 * rng.random(n_streams) ** 3) * 8 / 0.88).astype(np.int64)
    per = rng.integers(1, 5, n_streams)
    built = [hz.build_stream_cabac(int(per[i]), int(mean_bins[i]), config=5, n_active=64, n_ctx=64, slices_per_frame=2,
                                   frames_per_params=2, id_base=100 * i) for i in range(n_streams)]
    ops = max((b["ops"] for b in built), key=len)
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    blob = [struct.pack("<5I", n_streams, 64, len(ops), flags, 2)]
    for i, b in enumerate(built):
        s = np.ascontiguousarray(b["stream"])
        blob += [struct.pack("<QI", len(s), int(per[i])), s.tobytes()]
    n_ops = np.concatenate([b["n_ops"] for b in built]).astype(np.uint32)
    qp = capi.Context.slice_qp(np.concatenate([b["qp"] for b in built]), np.concatenate([b["idc"] for b in built]))
    blob += [np.ascontiguousarray(ops).astype(np.uint16).tobytes(), n_ops.tobytes(), qp.tobytes()]
    path = os.path.join(str(tmp_path), "batch.bin")
    with open(path, "wb") as f:
        f.write(b"".join(blob))
    rc, out = run(exe, "sched", path)
    assert rc == 0, out
    lines = {(l[0], int(l[1])): l for l in out if l[0] in ("stream", "slice")}
    row = 0
    workers = set()
    for i, b in enumerate(built):
        onal, _ = orc.read_nal_units_arrays(b["stream"])
        l = lines[("stream", i)]
        workers.add(int(l[3]))
        assert int(l[5]) == len(onal["start"])
        for s in range(int(per[i])):
            l = lines[("slice", row)]
            n = int(b["n_ops"][s]) + 1
            assert int(l[3]) == n and int(l[9]) == 0 and int(l[13]) == 1
            acc = 0
            for w in b["bins"][s, :n // 32]:
                acc = (acc * 1000003 + int(w)) & 0xFFFFFFFFFFFFFFFF
            assert int(l[11]) == acc, (i, s)
            row += 1
    assert workers == {0, 1}
    assert out[-1][:3] == ["total", "bins", str(int(n_ops.sum()) + len(n_ops))]


def test_plan_batch_wrapper_matches_the_python_binding(exe, tmp_path):
    """h264::PlanBatch (host/h264.hpp) == capi.scheduler_plan on the same batch: host only, runs without a GPU"""
    import struct
    from h264decode_b200 import capi
    rng = np.random.default_rng(21)
    n_streams, per = 48, 4
    sizes = (200 * 2 ** (8 * rng.random(n_streams) ** 3)).astype(np.int64)
    streams = []
    for i in range(n_streams):
        body = rng.integers(4, 256, int(sizes[i]), dtype=np.uint8)
        streams.append(np.concatenate([np.array([0, 0, 0, 1], np.uint8), body, np.array([0, 0, 0, 1], np.uint8)]))
    streams[5] = np.array([9, 9, 9], np.uint8)                      # no NAL unit in it
    n_ops = (100 * 2 ** (10 * rng.random(n_streams * per) ** 3)).astype(np.uint32)
    n_ops_max = int(n_ops.max())
    blob = [struct.pack("<5I", n_streams, 64, n_ops_max, 0, 1)]
    for s in streams:
        blob += [struct.pack("<QI", len(s), per), s.tobytes()]
    blob += [np.zeros(n_ops_max, np.uint16).tobytes(), n_ops.tobytes()]
    path = os.path.join(str(tmp_path), "plan.bin")
    with open(path, "wb") as f:
        f.write(b"".join(blob))
    for nd, group in ((1, 1024), (3, 2048), (8, 1 << 30)):
        rc, out = run(exe, "plan", path, nd, group)
        assert rc == 0, out
        dev, pas, cls = capi.scheduler_plan(streams, [per] * n_streams, n_ops, n_ops_max, nd, 148, group)
        got_s = [(int(l[3]), int(l[5])) for l in out if l[0] == "stream"]
        got_c = [int(l[3]) for l in out if l[0] == "slice"]
        assert got_s == [(int(d), int(p)) for d, p in zip(dev, pas)]
        assert got_c == [int(c) for c in cls]
        assert dev[5] == -1 and set(dev) - {-1} == set(range(min(nd, n_streams - 1)))
