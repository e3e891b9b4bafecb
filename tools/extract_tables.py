#!/usr/bin/env python3
"""Extract the CABAC data tables of the reference into C include files.

The reference keeps Table 9-44 (rangeTabLPS), Table 9-45 (state transitions) and the
(m, n) context-initialisation constants as Go map literals:

  /root/reference/h264/rangeTabLPS.go:5-70      rangeTabLPS      map[int][]int
  /root/reference/h264/stateTransxTab.go:9-74   stateTransxTab   map[int]StateTransx
  /root/reference/h264/mn_vars.go:15-175        MNVars           map[ctxIdx]map[idc]MN
  /root/reference/h264/mn_vars.go:184-440       CodedblockPatternMN switch (ctxIdx 70..104)

This script parses those literals (data only; no reference code is copied) and writes the same
numbers as flat C arrays, three times: for the CPU oracle (oracle/ref_tables.h), for the CUDA product
(h264decode_b200/csrc/tables.inc) and for the synthetic-data harness (harness/harness_tables.h).  The
outputs are deliberately separate files so that neither the product nor the harness includes anything
under oracle/.

Two variants are emitted:
  REF  : the numbers exactly as the reference has them, typos included (A1..A4) -- parity default.
  SPEC : A1 (rangeTabLPS row 33), A2 (transIdxMPS[59]) and A3 (the two-digit negative m values whose
         second digit the reference left in a "Second M" comment) corrected.  A4 is only suspected
         and is left untouched.

(m, n) lookup convention used by both outputs: column c in 0..3 where c = cabac_init_idc + 1, i.e.
c = 0 is the reference's key NoCabacInitIdc (-1) / the I,SI column of CodedblockPatternMN.
Anything the reference does not define is MN{0,0} (Go zero value for a missing map key,
mn_vars.go:439 fall-through).

Run from the repo root in a container that has /root/reference:
    python tools/extract_tables.py            # rewrite the two generated files
    python tools/extract_tables.py --check    # exit 1 if the committed files differ
"""
import argparse
import os
import re
import sys

REF = os.environ.get("H264_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CTX = 1024


def parse_range_tab(path):
    rows = {}
    for line in open(path):
        m = re.match(r"\s*(\d+):\s*\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\},", line)
        if m:
            rows[int(m.group(1))] = [int(m.group(i)) for i in range(2, 6)]
    assert sorted(rows) == list(range(64)), "rangeTabLPS must have rows 0..63"
    return [rows[i] for i in range(64)]


def parse_state_trans(path):
    rows = {}
    for line in open(path):
        m = re.match(r"\s*(\d+):\s*\{(\d+),\s*(\d+)\},", line)
        if m:
            rows[int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
    assert sorted(rows) == list(range(64)), "stateTransxTab must have rows 0..63"
    return [rows[i][0] for i in range(64)], [rows[i][1] for i in range(64)]


def parse_mn(path):
    """Returns (ref, spec): dict[(ctxIdx, col)] -> (m, n), col = idc + 1 (0 = NoCabacInitIdc / I column)."""
    text = open(path).read()
    ref, spec = {}, {}
    # ---- MNVars (ctxIdx 0..39)
    body = text[text.index("MNVars = map[int]map[int]MN{"):text.index("func MNSecond")]
    cur = None
    for line in body.splitlines():
        m = re.match(r"\s*(\d+):\s*map\[int\]MN\{(.*)$", line)
        if m:
            cur = int(m.group(1))
            rest = m.group(2)
            one = re.match(r"\s*NoCabacInitIdc:\s*MN\{(-?\d+),\s*(-?\d+)\}\},", rest)
            if one:
                ref[(cur, 0)] = spec[(cur, 0)] = (int(one.group(1)), int(one.group(2)))
                cur = None
            continue
        m = re.match(r"\s*(\d+):\s*MN\{(-?\d+),\s*(-?\d+)\},\s*(?://\s*Second M:\s*(\d))?", line)
        if m and cur is not None:
            idc, mm, nn = int(m.group(1)), int(m.group(2)), int(m.group(3))
            ref[(cur, idc + 1)] = (mm, nn)
            if m.group(4) is not None:  # A3: the dropped second digit of a negative m
                assert mm < 0
                mm = -(abs(mm) * 10 + int(m.group(4)))
            spec[(cur, idc + 1)] = (mm, nn)
    # ---- CodedblockPatternMN (ctxIdx 70..104)
    body = text[text.index("func CodedblockPatternMN"):]
    for m in re.finditer(
            r"case (\d+):.*?MN\{(-?\d+),\s*(-?\d+)\},\s*MN\{(-?\d+),\s*(-?\d+)\},\s*MN\{(-?\d+),\s*(-?\d+)\},"
            r".*?return MN\{(-?\d+),\s*(-?\d+)\}", body, re.S):
        g = [int(x) for x in m.groups()]
        ctx = g[0]
        for idc in range(3):
            ref[(ctx, idc + 1)] = spec[(ctx, idc + 1)] = (g[1 + 2 * idc], g[2 + 2 * idc])
        ref[(ctx, 0)] = spec[(ctx, 0)] = (g[7], g[8])
    ctxs = sorted({c for c, _ in ref})
    assert ctxs == list(range(0, 40)) + list(range(70, 105)), ctxs
    return ref, spec


def c_array(name, ctype, values, per_line=16):
    out = ["static const %s %s[%d] = {" % (ctype, name, len(values))]
    for i in range(0, len(values), per_line):
        out.append("    " + ", ".join("%d" % v for v in values[i:i + per_line]) + ",")
    out.append("};")
    return "\n".join(out)


def render(prefix, guard):
    rt = parse_range_tab(os.path.join(REF, "h264/rangeTabLPS.go"))
    lps, mps = parse_state_trans(os.path.join(REF, "h264/stateTransxTab.go"))
    mn_ref, mn_spec = parse_mn(os.path.join(REF, "h264/mn_vars.go"))

    rt_spec = [list(r) for r in rt]
    rt_spec[33] = [26, 31, 37, 43]          # A1
    mps_spec = list(mps)
    mps_spec[59] = 60                        # A2

    def mn_flat(d):
        m = [0] * (4 * N_CTX)
        n = [0] * (4 * N_CTX)
        for (ctx, col), (mm, nn) in d.items():
            m[col * N_CTX + ctx] = mm
            n[col * N_CTX + ctx] = nn
        return m, n

    parts = [
        "/* GENERATED by tools/extract_tables.py from the data literals of the reference",
        " * (h264/rangeTabLPS.go:5-70, h264/stateTransxTab.go:9-74, h264/mn_vars.go:15-175,184-440).",
        " * Do not edit.  REF = the reference's numbers verbatim (typos A1..A4 of SURVEY.md kept);",
        " * SPEC = A1, A2, A3 corrected.  (m,n) arrays are indexed [col*1024 + ctxIdx], col = cabac_init_idc+1,",
        " * col 0 = NoCabacInitIdc / the I,SI column; undefined entries are 0 (Go zero value). */",
        "#ifndef %s" % guard,
        "#define %s" % guard,
        "#include <stdint.h>",
        "#define %sN_CTX_MAX %d" % (prefix.upper(), N_CTX),
    ]
    for tag, rtab, lp, mp, mn in (("ref", rt, lps, mps, mn_ref), ("spec", rt_spec, lps, mps_spec, mn_spec)):
        parts.append(c_array("%srange_tab_lps_%s" % (prefix, tag), "uint8_t", [v for r in rtab for v in r]))
        parts.append(c_array("%strans_idx_lps_%s" % (prefix, tag), "uint8_t", lp))
        parts.append(c_array("%strans_idx_mps_%s" % (prefix, tag), "uint8_t", mp))
        m, n = mn_flat(mn)
        parts.append(c_array("%smn_m_%s" % (prefix, tag), "int8_t", m, 32))
        parts.append(c_array("%smn_n_%s" % (prefix, tag), "int8_t", n, 32))
    parts.append("#endif")
    return "\n".join(parts) + "\n"


TARGETS = [
    ("oracle/ref_tables.h", "orc_", "ORACLE_REF_TABLES_H"),
    ("h264decode_b200/csrc/tables.inc", "h264b_", "H264B_TABLES_INC"),
    ("harness/harness_tables.h", "hz_", "HARNESS_TABLES_H"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    bad = 0
    for rel, prefix, guard in TARGETS:
        text = render(prefix, guard)
        path = os.path.join(ROOT, rel)
        if args.check:
            if not os.path.exists(path) or open(path).read() != text:
                print("MISMATCH", rel)
                bad = 1
        else:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            open(path, "w").write(text)
            print("wrote", rel)
    sys.exit(bad)


if __name__ == "__main__":
    main()
