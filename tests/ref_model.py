"""ref_model.py -- an INDEPENDENT second model of the reference's hot path, for pinning oracle/oracle.c.

Test infrastructure only.  Written directly from the Go sources of mrmod/h264decode (not from oracle.c, not from
tools/extract_tables.py), in plain Python with Go `int` emulated as a wrapping int64, with its own reader of the Go
table literals.  tests/test_ref_model.py fuzzes the oracle against it for SURVEY.md section 8 rows N1-N6, E1-E7, I1-I4.

Reference lines each piece follows (paths relative to the reference root):
  isStartSequence                     h264/server.go:28-39
  readNalUnit / BufferToReader        h264/server.go:64-111, h264/bit_reader.go:27-39
  NewNalUnit + extension headers      h264/nalUnit.go:39-71,75-131
  BitReader.Read / NextField / ...    h264/bit_reader.go:162-172,228-236,263-332
  initDecodingEngine                  h264/cabac.go:439-446
  DecodeBypass / DecodeTerminate      h264/cabac.go:468-499
  RenormD                             h264/cabac.go:503-511
  BinaryDecision arithmetic core      h264/cabac.go:525-536
  StateTransitionProcess              h264/cabac.go:544-553
  SliceQPy / PreCtxState / Clip3      h264/cabac.go:113-139, state split :158-164
  rangeTabLPS, stateTransxTab         h264/rangeTabLPS.go, h264/stateTransxTab.go
  MNVars, CodedblockPatternMN         h264/mn_vars.go:15-175,184-440

The tables are read from the Go files when /root/reference is present (this container) and from the committed
snapshot tests/golden/go_tables.json otherwise (the GPU box); tools/make_go_tables.py writes the snapshot with this
module's own reader, and a CPU test checks the snapshot against the Go files whenever they are there.
"""
import json
import os
import re

REF_ROOT = "/root/reference"
SNAPSHOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "go_tables.json")
MASK64 = (1 << 64) - 1


class GoPanic(Exception):
    """A Go runtime panic (index out of range)."""


def i64(x):
    """Go int on amd64: wraps silently at 64 bits."""
    x &= MASK64
    return x - (1 << 64) if x >> 63 else x


# ----------------------------------------------------------------------------------------------------- Go literals
def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _ints(text):
    return [int(t) for t in re.findall(r"-?\d+", text)]


def read_go_tables(root=REF_ROOT):
    """Returns dict(range_tab_lps={p: [4]}, trans={p: (lps, mps)}, mn_vars={ctx: {idc: (m, n)}},
    cbp={ctx: ([(m, n)] * 3, (m, n) fallthrough)}) read from the Go composite literals."""
    h = os.path.join(root, "h264")
    out = {}
    # rangeTabLPS: `N: {a, b, c, d},` rows inside `map[int][]int{ ... }`
    src = _strip_comments(open(os.path.join(h, "rangeTabLPS.go")).read())
    body = src[src.index("map[int][]int{") + len("map[int][]int{"):]
    out["range_tab_lps"] = {int(k): _ints(v) for k, v in re.findall(r"(\d+)\s*:\s*\{([^}]*)\}", body)}
    # stateTransxTab: `N: {lps, mps},` rows (field order of StateTransx: TransIdxLPS, TransIdxMPS)
    src = _strip_comments(open(os.path.join(h, "stateTransxTab.go")).read())
    m = re.search(r"type\s+StateTransx\s+struct\s*\{\s*(\w+)\s*,\s*(\w+)\s+int", src)
    assert m and (m.group(1), m.group(2)) == ("TransIdxLPS", "TransIdxMPS"), "unexpected StateTransx field order"
    body = src[src.index("map[int]StateTransx{") + len("map[int]StateTransx{"):]
    out["trans"] = {int(k): tuple(_ints(v)) for k, v in re.findall(r"(\d+)\s*:\s*\{([^}]*)\}", body)}
    # MNVars: ctxIdx: map[int]MN{ key: MN{m, n}, ... } with key a number or NoCabacInitIdc (= -1)
    src = _strip_comments(open(os.path.join(h, "mn_vars.go")).read())
    no_idc = int(re.search(r"const\s+NoCabacInitIdc\s*=\s*(-?\d+)", src).group(1))
    start = src.index("MNVars = map[int]map[int]MN{")
    end = src.index("func MNSecond")
    mn_vars = {}
    for ctx, inner in re.findall(r"(\d+)\s*:\s*map\[int\]MN\{(.*?)\}\s*,?\s*(?=\d+\s*:\s*map\[int\]MN|\}\s*\))",
                                 src[start:end], flags=re.S):
        row = {}
        for key, mm, nn in re.findall(r"(NoCabacInitIdc|-?\d+)\s*:\s*MN\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", inner):
            row[no_idc if key == "NoCabacInitIdc" else int(key)] = (int(mm), int(nn))
        mn_vars[int(ctx)] = row
    out["mn_vars"] = mn_vars
    # CodedblockPatternMN: `case N:` ... `[]MN{ MN{..}, MN{..}, MN{..} }[cabacInitIdc]` ... `return MN{m, n}`
    fn = src[src.index("func CodedblockPatternMN"):]
    cbp = {}
    cases = re.split(r"\bcase\s+(\d+)\s*:", fn)
    for k in range(1, len(cases), 2):
        ctx, body = int(cases[k]), cases[k + 1]
        lst = re.search(r"\[\]MN\{(.*?)\}\s*\[cabacInitIdc\]", body, flags=re.S)
        cols = [(int(a), int(b)) for a, b in re.findall(r"MN\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", lst.group(1))]
        rest = body[lst.end():]
        ret = re.search(r"return\s+MN\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", rest)
        cbp[ctx] = (cols, (int(ret.group(1)), int(ret.group(2))))
    out["cbp"] = cbp
    return out


def _to_json(t):
    return {"range_tab_lps": {str(k): v for k, v in t["range_tab_lps"].items()},
            "trans": {str(k): list(v) for k, v in t["trans"].items()},
            "mn_vars": {str(c): {str(i): list(v) for i, v in row.items()} for c, row in t["mn_vars"].items()},
            "cbp": {str(c): [[list(x) for x in cols], list(ret)] for c, (cols, ret) in t["cbp"].items()}}


def _from_json(j):
    return {"range_tab_lps": {int(k): list(v) for k, v in j["range_tab_lps"].items()},
            "trans": {int(k): tuple(v) for k, v in j["trans"].items()},
            "mn_vars": {int(c): {int(i): tuple(v) for i, v in row.items()} for c, row in j["mn_vars"].items()},
            "cbp": {int(c): ([tuple(x) for x in v[0]], tuple(v[1])) for c, v in j["cbp"].items()}}


def load_tables():
    if os.path.isdir(os.path.join(REF_ROOT, "h264")):
        return read_go_tables()
    return _from_json(json.load(open(SNAPSHOT)))


_T = None


def tables():
    global _T
    if _T is None:
        _T = load_tables()
    return _T


def spec_tables():
    """The documented corrections A1 / A2 of SURVEY.md Appendix A applied to the Go tables (H264B_TABLES_SPEC)."""
    t = tables()
    r = {k: list(v) for k, v in t["range_tab_lps"].items()}
    r[33] = [26, 31, 37, 43]
    tr = dict(t["trans"])
    tr[59] = (tr[59][0], 60)
    return r, tr


# ----------------------------------------------------------------------------------------------------- BitReader
def bit_array(b):
    """degolomb.BitArray (un-vendored dependency): element i = bit 7 - i, the only order under which NewNalUnit's
    forbidden_zero_bit(1) / nal_ref_idc(2) / nal_unit_type(5) come out of byte 0 (nalUnit.go:82-84)."""
    return [(b >> (7 - i)) & 1 for i in range(8)]


def bit_val(bits):
    t = 0
    for i, b in enumerate(bits):
        if b == 1:
            sh = (len(bits) - 1) - i
            t = i64(t + (i64(1 << sh) if sh < 64 else 0))  # Go: 1 << uint(k) is 0 for k >= 64; int wraps
    return t


class BitReader:
    def __init__(self, data):
        self.bytes = bytes(data)
        self.byteOffset = 0
        self.bitOffset = 0
        self.bitsRead = 0

    def set_offset(self):
        self.byteOffset = self.bitsRead // 8
        self.bitOffset = self.bitsRead % 8

    def read(self, n):
        """(*BitReader).Read: fills n ints; indexing past the last byte panics (the EOF tests are off by one)."""
        buf = [0] * n
        if self.byteOffset > len(self.bytes):
            raise GoPanic("EOF error path (returns an error; NextField then yields -1)")
        i = 0
        if n == 0:  # the Go loop would index buf[0] of an empty slice
            raise GoPanic("index out of range [0] with length 0")
        while True:
            if self.byteOffset >= len(self.bytes):
                raise GoPanic("index out of range")
            for bit in bit_array(self.bytes[self.byteOffset])[self.bitOffset:8]:
                buf[i] = bit
                i += 1
                self.bitsRead += 1
                self.set_offset()
                if i >= n:
                    return buf

    def next_field(self, n):
        return bit_val(self.read(n))

    def read_one_bit(self):
        return self.read(1)[0]

    def peek_bytes(self, n):
        if len(self.bytes) >= self.byteOffset + n:
            return self.bytes[self.byteOffset:self.byteOffset + n]
        return None

    def read_byte(self):
        if len(self.bytes) > self.byteOffset:
            b = self.bytes[self.byteOffset]
            self.byteOffset += 1
            return b
        return None


# ----------------------------------------------------------------------------------------------------- NAL units
NAL_FIELDS = ["NumBytes", "ForbiddenZeroBit", "RefIdc", "Type", "SvcExtensionFlag", "Avc3dExtensionFlag", "IdrFlag",
              "PriorityId", "NoInterLayerPredFlag", "DependencyId", "QualityId", "TemporalId", "UseRefBasePicFlag",
              "DiscardableFlag", "OutputFlag", "ReservedThree2Bits", "HeaderBytes", "NonIdrFlag", "ViewId",
              "AnchorPicFlag", "InterViewFlag", "ReservedOneBit", "ViewIdx", "DepthFlag",
              "EmulationPreventionThreeByte"]


def new_nal_unit(frame, num_bytes_in_nal):
    n = {k: 0 for k in NAL_FIELDS}
    n["NumBytes"] = num_bytes_in_nal
    n["HeaderBytes"] = 1
    b = BitReader(frame)
    n["ForbiddenZeroBit"] = b.next_field(1)
    n["RefIdc"] = b.next_field(2)
    n["Type"] = b.next_field(5)
    if n["Type"] in (14, 20, 21):
        if n["Type"] != 21:
            n["SvcExtensionFlag"] = b.next_field(1)
        else:
            n["Avc3dExtensionFlag"] = b.next_field(1)
        if n["SvcExtensionFlag"] == 1:
            for name, w in (("IdrFlag", 1), ("PriorityId", 6), ("NoInterLayerPredFlag", 1), ("DependencyId", 3),
                            ("QualityId", 4), ("TemporalId", 3), ("UseRefBasePicFlag", 1), ("DiscardableFlag", 1),
                            ("OutputFlag", 1), ("ReservedThree2Bits", 2)):
                n[name] = b.next_field(w)
            n["HeaderBytes"] += 3
        elif n["Avc3dExtensionFlag"] == 1:
            for name, w in (("ViewIdx", 8), ("DepthFlag", 1), ("NonIdrFlag", 1), ("TemporalId", 3),
                            ("AnchorPicFlag", 1), ("InterViewFlag", 1)):
                n[name] = b.next_field(w)
            n["HeaderBytes"] += 2
        else:
            for name, w in (("NonIdrFlag", 1), ("PriorityId", 6), ("ViewId", 10), ("TemporalId", 3),
                            ("AnchorPicFlag", 1), ("InterViewFlag", 1), ("ReservedOneBit", 1)):
                n[name] = b.next_field(w)
            n["HeaderBytes"] += 3
    rbsp = bytearray()
    i = n["HeaderBytes"]
    while i < n["NumBytes"]:
        nxt = b.peek_bytes(3)
        if nxt is None:
            break
        if i + 2 < n["NumBytes"] and nxt[0] == 0 and nxt[1] == 0 and nxt[2] == 3:
            three = [b.read_byte(), b.read_byte(), b.read_byte()]
            rbsp += bytes(three[:2])
            i += 2
            n["EmulationPreventionThreeByte"] = three[2]
        else:
            one = b.read_byte()
            if one is None:
                break
            rbsp.append(one)
        i += 1
    return n, bytes(rbsp)


def is_start_sequence(packet):
    return len(packet) >= 4 and bytes(packet[-4:]) == b"\x00\x00\x00\x01"


def read_nal_units(stream):
    """The handleConnection loop over readNalUnit: returns [(startOffset, endOffset, NalUnit fields, rbsp)].
    The H264Reader grows `buf` one byte per BufferToReader(1); byteOffset counts the bytes taken."""
    stream = bytes(stream)
    buf = bytearray()
    pos = 0  # h.byteOffset == len(buf): every byte read is appended

    def buffer_one():
        nonlocal pos
        if pos >= len(stream):
            return False  # io.EOF
        buf.append(stream[pos])
        pos += 1
        return True

    out = []
    while True:
        ok = True
        while not is_start_sequence(buf):
            if not buffer_one():
                ok = False
                break
        if not ok:
            break
        start = pos
        so = pos
        while so == start or not is_start_sequence(buf):
            so = pos
            if not buffer_one():
                ok = False
                break
        if not ok:
            break
        end = pos
        frame = bytes(buf[start:])
        if len(frame) < 8:
            # server.go:108 slices [0:8]: legal while cap >= 8 (A13); the model keeps going like the oracle
            pass
        fields, rbsp = new_nal_unit(frame, len(frame))
        out.append((start, end, fields, rbsp))
    return out


# ----------------------------------------------------------------------------------------------------- CABAC engine
def init_decoding_engine(br):
    return 510, br.next_field(9)


def decode_bypass(br, R, O, spec_or=False):
    O = i64(O << 1)
    bit = br.read_one_bit()
    O = i64(O | bit) if spec_or else i64(O << bit)
    if O >= R:
        return i64(O - R), 1
    return O, 0


def renorm_d(br, R, O):
    while R < 256:
        R = i64(R << 1)
        O = i64(O << 1)
        O = O | br.read_one_bit()
    return R, O


def decode_terminate(br, R, O):
    R = i64(R - 2)
    if O >= R:
        return R, O, 1
    R, O = renorm_d(br, R, O)
    return R, O, 0


def binary_decision(p_state_idx, val_mps, R, O, range_tab):
    q = (R >> 6) & 3
    lps = range_tab[p_state_idx][q]
    R = i64(R - lps)
    if O >= R:
        return 1 - val_mps, lps, i64(O - R)
    return val_mps, R, O


def state_transition(p_state_idx, val_mps, bin_val, trans):
    if bin_val == val_mps:
        return trans[p_state_idx][1], val_mps
    if p_state_idx == 0:
        val_mps = 1 - val_mps
    return trans[p_state_idx][0], val_mps


def decode_decision(br, state, R, O, range_tab, trans):
    """9.3.3.2.1 as the reference's pieces compose: BinaryDecision core, StateTransitionProcess, RenormD.
    state = (pStateIdx, valMPS)."""
    b, R, O = binary_decision(state[0], state[1], R, O, range_tab)
    state = state_transition(state[0], state[1], b, trans)
    R, O = renorm_d(br, R, O)
    return b, state, R, O


OP_DECISION, OP_BYPASS, OP_TERMINATE = 0, 1, 2


def decode_slice(data, ops, states, spec_or=False, spec_tables_=False):
    """ops: [(kind, ctxIdx)]; states: list of state bytes (pStateIdx | valMPS << 6), modified in place.
    Returns (bins, R, O, bitsRead, panicked): an op that runs off the end of the data leaves everything as it was
    before that op (the checker's convention for a Go panic) and ends the slice."""
    range_tab, trans = spec_tables() if spec_tables_ else (tables()["range_tab_lps"], tables()["trans"])
    br = BitReader(data)
    bins = []
    try:
        R, O = init_decoding_engine(br)
    except GoPanic:
        return bins, 0, 0, br.bitsRead, True
    for kind, ctx in ops:
        save = (R, O, br.bitsRead)
        try:
            if kind == OP_DECISION:
                if ctx >= len(states):
                    ctx = 0
                s = states[ctx]
                b, (p, v), R2, O2 = decode_decision(br, (s & 63, (s >> 6) & 1), R, O, range_tab, trans)
                states[ctx] = p | (v << 6)
                R, O = R2, O2
            elif kind == OP_BYPASS:
                O, b = decode_bypass(br, R, O, spec_or)
            else:
                R, O, b = decode_terminate(br, R, O)
        except GoPanic:
            R, O, br.bitsRead = save
            return bins, R, O, br.bitsRead, True
        bins.append(b)
    return bins, R, O, br.bitsRead, False


# ----------------------------------------------------------------------------------------------------- context init
def clip3(x, y, z):
    if z < x:
        return x
    if z > y:
        return y
    return z


def slice_qpy(pic_init_qp_minus26, slice_qp_delta):
    return 26 + pic_init_qp_minus26 + slice_qp_delta


def pre_ctx_state(m, n, slice_qpy_):
    return clip3(1, 126, ((m * clip3(0, 51, slice_qpy_)) >> 4) + n)  # Python's >> floors like Go's


def state_split(pre):
    if pre <= 63:
        return 63 - pre, 0
    return pre - 64, 1


def mn_lookup(ctx_idx, cabac_init_idc):
    """(m, n) of a context: CodedblockPatternMN for the ctxIdx it switches on, MNVars[ctxIdx][idc] otherwise, with
    Go's zero value MN{0, 0} for a missing key / an unknown ctxIdx."""
    t = tables()
    if ctx_idx in t["cbp"]:
        cols, ret = t["cbp"][ctx_idx]
        if 0 <= cabac_init_idc <= 2:
            return cols[cabac_init_idc]
        return ret
    return t["mn_vars"].get(ctx_idx, {}).get(cabac_init_idc, (0, 0))


def ctx_state_byte(ctx_idx, cabac_init_idc, qp):
    m, n = mn_lookup(ctx_idx, cabac_init_idc)
    p, v = state_split(pre_ctx_state(m, n, qp))
    return p | (v << 6)
