#!/usr/bin/env python3
"""Executed instructions and stall samples per SOURCE LINE of one kernel: joins the SASS page of an .ncu-rep with the line
table of the built library (nvdisasm -g), instruction by instruction.
   python tools/ncu_lines.py rep.ncu-rep kernel-substring [library.so] [min-share-percent]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, filt):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name" and filt in rows[k][1]:
            hdr = rows[k + 1]
            body = []
            k += 2
            while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
                if len(rows[k]) == len(hdr):
                    body.append(rows[k])
                k += 1
            return hdr, body
        k += 1
    raise SystemExit("kernel not found")


def line_table(lib, filt):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.\S*%s\S*:$" % re.escape(filt), txt, re.M)
        if not m:
            continue
        lines = []
        cur = ("?", 0)
        stack = ""
        for ln in txt[m.end():].split("\n"):
            s = ln.strip()
            if s.startswith("//## File"):
                mm = re.match(r'//## File "([^"]+)", line (\d+)(.*)', s)
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                stack = mm.group(3)
                if OUTER:  # attribute to the outermost call site instead of the innermost line
                    chain = re.findall(r'inlined at "([^"]+)", line (\d+)', stack)
                    if chain:
                        cur = (os.path.basename(chain[-1][0]), int(chain[-1][1]))
                continue
            if s.startswith(".section") or s.startswith("//-----"):
                break
            mm = re.match(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", s)
            if mm:
                lines.append((cur, stack, mm.group(2)))
        return lines
    raise SystemExit("kernel not in library")


OUTER = bool(int(os.environ.get("NCU_LINES_OUTER", "0")))


def main():
    rep, filt = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else "h264decode_b200/libh264b200.so"
    min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    hdr, body = sass_rows(rep, filt)
    ix = {h: i for i, h in enumerate(hdr)}
    lt = line_table(lib, filt)
    if len(lt) != len(body):
        print("warning: %d instructions in the report, %d in the library (different build?)" % (len(body), len(lt)))
    agg = {}
    tot_e = tot_s = 0.0
    for r, (cur, stack, _) in zip(body, lt):
        e = float(r[ix["Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        a = agg.setdefault(cur, [0.0, 0.0, 0])
        a[0] += e
        a[1] += s
        a[2] += 1
        tot_e += e
        tot_s += s
    print("kernel %s: %.1f M instructions executed, %d samples" % (filt, tot_e / 1e6, tot_s))
    src_cache = {}
    for (f, l), (e, s, n) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if 100 * e / tot_e < min_share and 100 * s / tot_s < min_share:
            continue
        text = ""
        for root in ("h264decode_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().split("\n")
                if 0 < l <= len(src_cache[p]):
                    text = src_cache[p][l - 1].strip()[:90]
                break
        print("%-18s %5d  exec %5.1f%%  samples %5.1f%%  (%3d sass)  %s" % (f, l, 100 * e / tot_e, 100 * s / tot_s, n, text))


if __name__ == "__main__":
    main()
