#!/usr/bin/env python3
"""Per-SASS-instruction stall profile of one kernel from an .ncu-rep (ncu --set full --import-source on):
   python tools/ncu_sass.py rep.ncu-rep [kernel-substring] [min-executed]
prints address, instruction, executions, stall samples, samples per 1000 executions and the dominant stall reasons."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    min_exec = float(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            name = rows[k][1]
            hdr = rows[k + 1]
            k += 2
            body = []
            while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
                body.append(rows[k])
                k += 1
            if filt in name:
                report(name, hdr, body, min_exec)
                return
        else:
            k += 1


def report(name, hdr, body, min_exec):
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("kernel:", name)
    total = sum(float(r[ix["# Samples"]] or 0) for r in body if len(r) > ix["# Samples"])
    print("total samples:", total)
    for r in body:
        if len(r) <= ix["# Samples"]:
            continue
        ex = float(r[ix["Instructions Executed"]] or 0)
        if ex < min_exec:
            continue
        smp = float(r[ix["# Samples"]] or 0)
        reasons = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        rs = " ".join("%s:%d" % (c, v) for v, c in reasons if v > 0)
        print("%s  %-58s %10d %7d %7.2f  %s" % (r[ix["Address"]][-5:], r[ix["Source"]].strip()[:58], ex, smp,
                                              1000.0 * smp / ex if ex else 0, rs))


if __name__ == "__main__":
    main()
