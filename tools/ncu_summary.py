#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full --import-source on) into text: per-kernel key metrics, stall reasons, and
the hottest source lines.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/x.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEY = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "sm__cycles_elapsed.avg", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_xu.sum",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows:
        name = r[idx["Kernel Name"]]
        if filt not in name or name in seen or num(r[idx["sm__cycles_elapsed.avg"]]) is None:
            continue
        seen.add(name)
        print("=" * 100)
        print("kernel:", name)
        for k in KEY:
            if k in idx:
                print("  %-72s %s %s" % (k, r[idx[k]], units[idx[k]]))
        st = [(h, num(r[i])) for h, i in idx.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and num(r[i])]
        tot = sum(v for _, v in st) or 1
        print("  stall samples (pc sampling):")
        for h, v in sorted(st, key=lambda x: -x[1])[:10]:
            print("    %-40s %8.0f  %5.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v, 100 * v / tot))
    # source page: hottest source lines (per-file sections; rows with a line number carry the per-line aggregates)
    src = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] +
              (["-k", "regex:" + filt] if filt else []))
    agg = {}
    cur, hd = "", None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hd = r
        elif hd and r[0].isdigit():
            try:
                si, ii, ti = hd.index("# Samples"), hd.index("Instructions Executed"), hd.index("Thread Instructions Executed")
                bi = hd.index("stall_barrier")
                key = (cur, int(r[0]), r[1].strip())
                v = agg.setdefault(key, [0.0, 0.0, 0.0, 0.0])
                v[0] += num(r[si]) or 0
                v[1] += num(r[ii]) or 0
                v[2] += num(r[ti]) or 0
                v[3] += num(r[bi]) or 0
            except Exception:
                pass
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print("=" * 100)
    print("hottest source lines (%% of stall samples | %% of warp instructions | avg active threads | barrier samples):")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:45]:
        print("  %5.1f%% %5.1f%% %5.1f %6.0f  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / toti, v[2] / v[1] if v[1] else 0,
                                                  v[3], k[0], k[1], k[2][:110]))


if __name__ == "__main__":
    main()
