#!/usr/bin/env python3
"""profiles/r2_constants.json: the ncu-derived numbers bench.py quotes, with where they came from.
   python tools/ncu_constants.py <cabac.ncu-rep> <bins per launch> [<launch-list.csv> <alg bytes per scan pass>]
cabac.ncu-rep: ncu --set full of one cabac_decode_kernel launch of `python bench.py` (configs[3]); launch-list.csv: the
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum pass of the same command."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def num(x):
    return float(str(x).replace(",", ""))


def main():
    rep, bins = sys.argv[1], float(sys.argv[2])
    rows = [r for r in raw_rows(rep) if "cabac_decode_kernel" in r.get("Kernel Name", "")]
    r = rows[-1]
    inst = num(r["smsp__inst_executed.sum"])
    k = {
        "source": os.path.relpath(rep, ROOT),
        "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
        "cabac_bins_per_launch": bins,
        "cabac_warp_inst_per_bin": inst / (bins / 32.0),
        "cabac_dram_bytes_per_launch": num(r["dram__bytes_read.sum"]) * (1e9 if "Gbyte" in str(r.get("dram__bytes_read.sum (unit)", "")) else 1)
        if False else None,
        "cabac_issue_active_pct_of_active": num(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "cabac_issue_utilisation_elapsed": inst / (num(r["sm__cycles_elapsed.avg"]) * 592.0),
        "cabac_smem_bank_conflict_frac": num(r["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]) /
        max(1.0, num(r["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"])),
        "cabac_registers": num(r["launch__registers_per_thread"]),
        "cabac_gpu_time_ms_under_ncu": None,
    }
    # units: the raw page gives dram bytes in the unit of row 2 of the csv; read them from the details page instead
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    hdr, units = rr[0], rr[1]
    unit = dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def bytes_of(name):
        return num(r[name]) * scale.get(unit.get(name, "byte"), 1.0)

    k["cabac_dram_bytes_per_launch"] = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    tu = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    k["cabac_gpu_time_ms_under_ncu"] = num(r["gpu__time_duration.sum"]) * tu.get(unit.get("gpu__time_duration.sum", "ms"), 1.0)
    if len(sys.argv) > 4:
        alg = float(sys.argv[4])
        tot = 0.0
        names = set()
        for row in csv.DictReader(open(sys.argv[3])):
            if row.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum") and any(
                    t in row.get("Kernel Name", "") for t in ("annexb_", "order_", "nal_permute", "scan_finalize", "nal_fixup")):
                tot += num(row["Metric Value"]) * scale.get(row.get("Metric Unit", "byte"), 1.0)
                names.add(row["Kernel Name"].split("(")[0])
        passes = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
        k["scan_dram_bytes_per_alg_byte"] = tot / passes / alg
        k["scan_source"] = os.path.relpath(sys.argv[3], ROOT)
    json.dump(k, open(os.path.join(ROOT, "profiles", "r2_constants.json"), "w"), indent=1)
    print(json.dumps(k, indent=1))


if __name__ == "__main__":
    main()
