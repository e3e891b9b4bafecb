#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json: "CABAC bins/s + Annex-B NAL/RBSP GB/s").

A step = one pass of the whole front end over one synthetic Annex-B stream that is already resident in HBM:
    Annex-B split + RBSP emulation-prevention strip (K1/K2)  ->  slice NAL list  ->  CABAC engine (K3, contexts
    initialised in-kernel by the K4 rule).
Workload (config.workload): BASELINE configs[3] "1080p-shaped stream": 8 slices/frame, SPS+PPS every 250 frames, slice
payloads = encoder-generated CABAC data (shared op schedule of SURVEY.md §8d) ~50 KB each, emulation-prevention
escaped.  `--frames 10000` is the full ~4 GB config (default); every rank processes its own stream ("sharded by
stream", weak scaling, no collective on the data path).

value      = bins decoded by all ranks / max-over-ranks device time of the K timed steps (inputs resident in HBM)
e2e        = the same pass through the host-buffer entry point h264b_stream_decode: pinned host stream in,
             NAL index + packed bins + final engine states back in pinned host memory, copies inside the timed region
roofline   = Annex-B scan kernel group vs the measured HBM copy peak (MEASURED_PEAKS.json)
cpu_baseline = the oracle (C restatement of the Go reference; no Go toolchain exists) on the host cores, bounded sample

`--impl reference` times that CPU restatement alone (all host threads) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# DRAM bytes (read + write) per algorithmic byte of the Annex-B pass, from the ncu capture of this command line
# (profiles/r1_ncu_launch_list_bench.csv: dram__bytes_read.sum + dram__bytes_write.sum of the seven kernels of one pass,
# 7.885 GB against 7.866 GB algorithmic; annexb_copy_kernel alone 7.826 GB, profiles/r1_ncu_copy_kernel_summary.txt)
TRAFFIC_PER_ALG_BYTE = 1.0024
TRAFFIC_SOURCE = "ncu dram__bytes_read.sum + dram__bytes_write.sum per pass (profiles/r1_ncu_launch_list_bench.csv)"
MEAN_BINS = 455_000      # ~50 KB of CABAC data per slice at ~0.88 bit/bin
N_ACTIVE = 64
IN_FLIGHT = 3            # H264B_STREAM_JOBS_IN_FLIGHT
CABAC_WARP_INST_PER_OP = 38.9   # ncu: 44.26 G warp instructions / 1.137 G warp-ops (profiles/r1_ncu_cabac_final2_summary.txt)
N_CTX = 64
SLICES_PER_FRAME = 8
FRAMES_PER_PARAMS = 250


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_pass(sample, threads):
    """The oracle's pass over a bounded sample: literal readNalUnit/NewNalUnit over the sample stream, then
    initDecodingEngine + one primitive per op for each slice, `threads` slices at a time (one thread per slice,
    mirroring the reference's goroutine-per-stream).  Returns (bins, seconds, seconds_scan, seconds_cabac)."""
    from oracle import oracle as orc
    t0 = time.perf_counter()
    nal, rbsp = orc.read_nal_units_arrays(sample["stream"], literal=True)
    t1 = time.perf_counter()
    sl = np.flatnonzero((nal["type"] == 1) | (nal["type"] == 5))
    init = orc.ctx_init(sample["qp"][:len(sl)], sample["idc"][:len(sl)], N_CTX)
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    bins = [0] * len(sl)

    def work(idx):
        for i in idx:
            k = sl[i]
            data = rbsp[nal["rbsp_off"][k]:nal["rbsp_off"][k] + nal["rbsp_len"][k]]
            ops = np.concatenate([sample["ops"][:sample["n_ops"][i]], term])
            rc, _, fin, _ = orc.cabac_decode_slice(data, ops, init[i], orc.BYPASS_SPEC_OR)
            bins[i] = fin["n_bins"]

    t2 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(range(t, len(sl), threads),)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    t3 = time.perf_counter()
    return sum(bins), (t1 - t0) + (t3 - t2), t1 - t0, t3 - t2


def make_cpu_sample(n_slices, id_base=0):
    import harness as hz
    b = hz.build_stream_cabac(n_slices, MEAN_BINS, config=4, n_active=N_ACTIVE, n_ctx=N_CTX,
                              slices_per_frame=SLICES_PER_FRAME, frames_per_params=FRAMES_PER_PARAMS, id_base=id_base)
    return b


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    n_slices = args.cpu_slices or max(threads * 16, 64)
    sample = make_cpu_sample(n_slices)
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_reference_pass(sample, threads)
    tot_bins, tot_s = 0, 0.0
    for _ in range(args.steps):
        b, s, _, _ = cpu_reference_pass(sample, threads)
        tot_bins += b
        tot_s += s
    v = tot_bins / tot_s
    desc = "%d slices (~%d KB each, %.1f MB Annex-B) of the same generator, literal oracle" % (
        n_slices, MEAN_BINS * 0.88 / 8 / 1000, len(sample["stream"]) / 1e6)
    print(json.dumps({
        "impl": "reference", "metric": "cabac_bins_per_s", "value": v, "unit": "bins/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": v, "unit": "bins/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement (oracle/) of the pure-Go reference: no Go toolchain in this image, reference does not "
                "build as shipped",
    }))


def config_dict(args):
    return {"workload": "configs[3]: 1080p-shaped Annex-B stream, %d frames x %d slices/frame, ~50 KB CABAC slice "
                        "payloads (shared op schedule, 64 active contexts), SPS+PPS every %d frames; split + EPB strip "
                        "+ CABAC bins; one stream per GPU" % (args.frames, SLICES_PER_FRAME, FRAMES_PER_PARAMS),
            "frames": args.frames, "slices_per_stream": args.frames * SLICES_PER_FRAME, "mean_bins_per_slice": MEAN_BINS,
            "n_ctx": N_CTX, "bypass_form": os.environ.get("H264B_BENCH_BYPASS_FORM", "SPEC_OR"), "tables": "REF",
            "l2": "inputs (GBs per step) far exceed the 126 MB L2; no explicit flush needed",
            "parallelism": "stream-sharded, no collective"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args, rank, world, local_rank):
    import torch
    import harness as hz
    from h264decode_b200 import capi

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = capi.Context(local_rank)
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    # a dedicated (non-default) stream: torch's events and the library's launches must be on the same one, and a NULL
    # handle would mean "the context's own stream" to h264b_set_stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    if os.environ.get("H264B_BENCH_BYPASS_FORM") == "REF_SHIFT":  # diagnostic: the literal int64 engine on the same input
        flags = capi.CABAC_FINAL_TERMINATE

    # ---- synthetic input, generated on the GPU by the harness (outside every timed region)
    n_slices = args.frames * SLICES_PER_FRAME
    t_gen = time.perf_counter()
    g = hz.gpu_build_stream_cabac(torch, dev, n_slices, MEAN_BINS, config=4, n_active=N_ACTIVE, n_ctx=N_CTX,
                                  slices_per_frame=SLICES_PER_FRAME, frames_per_params=FRAMES_PER_PARAMS,
                                  id_base=rank * n_slices, want_bins=False)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    n = g["n"]
    d_stream = g["stream"]
    n_nals = g["n_nals"]
    nal_cap = n_nals + 16
    ops, n_ops, qp, idc = g["ops"], g["n_ops"], g["qp"], g["idc"]
    total_bins = int(n_ops.astype(np.int64).sum()) + n_slices

    # ---- device buffers of one step
    d_rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    d_nals = torch.empty(nal_cap * 32, dtype=torch.uint8, device=dev)
    d_sum = torch.zeros(64, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n_slices, dtype=torch.int64, device=dev)
    d_len = torch.empty(n_slices, dtype=torch.int32, device=dev)
    d_snal = torch.empty(n_slices, dtype=torch.int32, device=dev)
    d_ns = torch.zeros(4, dtype=torch.int32, device=dev)
    d_ops = torch.from_numpy(ops.view(np.int16)).to(dev)
    d_nops = torch.from_numpy(n_ops.view(np.int32)).to(dev)
    p = capi.Context.slice_qp(qp, idc)
    d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
    boff = np.zeros(n_slices + 1, dtype=np.uint64)
    boff[1:] = np.cumsum((n_ops.astype(np.uint64) + 1 + 31) // 32)
    d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)
    d_bins = torch.empty(int(boff[-1]), dtype=torch.int32, device=dev)
    d_fin = torch.empty(n_slices * 32, dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(events=None):
        if events:
            events[0].record(stream)
        ctx.annexb_scan_dev(d_stream.data_ptr(), n, d_rbsp.data_ptr(), d_nals.data_ptr(), None, nal_cap,
                            d_sum.data_ptr(), 0)
        if events:
            events[1].record(stream)
        ctx.slice_select_dev(d_nals.data_ptr(), d_sum.data_ptr(), nal_cap, 0, n_slices, d_off.data_ptr(),
                             d_len.data_ptr(), d_snal.data_ptr(), d_ns.data_ptr())
        ctx.cabac_decode_dev(bytes=d_rbsp.data_ptr(), total_bytes=n + 16, off=d_off.data_ptr(), len=d_len.data_ptr(),
                             n_slices=n_slices, n_ctx=N_CTX, ops=d_ops.data_ptr(), n_ops_max=len(ops),
                             n_ops=d_nops.data_ptr(), qp=d_qp.data_ptr(), init_states=None, bins=d_bins.data_ptr(),
                             bins_off=d_boff.data_ptr(), bins_stride_words=0, final=d_fin.data_ptr(),
                             final_states=None, flags=flags)
        if events:
            events[2].record(stream)

    # ---- warm-up (also: result sanity outside the timed region)
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    summ = np.frombuffer(d_sum.cpu().numpy().tobytes()[:48], dtype=np.uint64, count=5)
    fin = np.frombuffer(d_fin.cpu().numpy().tobytes(), dtype=capi.FINAL_DTYPE)
    ok = (int(summ[1]) == n_nals and int(d_ns.cpu()[0]) == n_slices and int(fin["n_bins"].astype(np.int64).sum())
          == total_bins and not (fin["flags"] & capi.F_OVERRUN).any()
          and np.array_equal(fin["n_bins"], n_ops + 1))
    # last bin of every slice is the terminate bin the encoder wrote (1), a cheap whole-workload self-check
    last_word = d_bins[torch.from_numpy((boff[1:] - 1).astype(np.int64)).to(dev)].cpu().numpy().view(np.uint32)
    ok = ok and bool(np.all((last_word >> (n_ops & 31).astype(np.uint32)) & 1 == 1))
    rbsp_bytes = int(summ[2])

    # ---- timed region: exactly K steps
    sampler = ClockSampler(local_rank)
    evs = [[ev(), ev(), ev()] for _ in range(args.steps)]
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    e_beg, e_end = ev(), ev()
    e_beg.record(stream)
    for k in range(args.steps):
        step(evs[k])
    e_end.record(stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    t_total_ms = e_beg.elapsed_time(e_end)
    t_scan_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    t_cabac_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))

    # ---- e2e through the host-buffer entry point
    del d_bins, d_rbsp
    torch.cuda.empty_cache()
    e2e_steps = max(IN_FLIGHT, min(args.steps, 12))
    e2e_error = None
    t_e2e, e2e_ok, h_stream = 0.0, True, None
    try:
        h_stream = ctx.host_alloc(n)
        ctx.d2h(h_stream, d_stream.data_ptr())
        ctx.sync()
        # warm-up: every job slot grows its pinned / device buffers
        tk = [_stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags) for _ in range(IN_FLIGHT)]
        for t in tk:
            r = _stream_wait_raw(ctx, capi, t)
        barrier()
        # timed: every step copies its stream in and its results out; consecutive steps overlap (three jobs in flight:
        # H2D of step k+1 | kernels of step k | D2H of step k-1), which is how an ingest loop drives the library
        t0 = time.perf_counter()
        pending = []
        for k in range(e2e_steps):
            if len(pending) == IN_FLIGHT:
                r = _stream_wait_raw(ctx, capi, pending.pop(0))
            pending.append(_stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags))
        while pending:
            r = _stream_wait_raw(ctx, capi, pending.pop(0))
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / e2e_steps
        e2e_ok = r["n_slices"] == n_slices and r["total_bins"] == total_bins
    except Exception as ex:  # e.g. not enough pinned host memory for every rank of a big box: report, do not die
        e2e_error = "%s: %s" % (type(ex).__name__, ex)
        if dist is not None:
            try:
                barrier()
            except Exception:
                pass
    d2h_bytes = int(boff[-1]) * 4 + n_slices * 32 + n_slices * 4 + n_nals * 32 + 48 + 4
    h2d_bytes = n + len(ops) * 2 + n_slices * (8 + 4) + (n_slices + 1) * 8

    # ---- reduce over ranks
    t_max_ms, t_e2e_max, bins_all, bytes_all = t_total_ms, t_e2e, total_bins, n
    if dist is not None:
        t = torch.tensor([t_total_ms, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_max_ms, t_e2e_max = float(t[0]), float(t[1])
        c = torch.tensor([total_bins, n, int(ok and e2e_ok), int(e2e_error is not None)], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        bins_all, bytes_all = int(c[0]), int(c[1])
        ok = int(c[2]) == world
        if int(c[3]) and e2e_error is None:
            e2e_error = "the end-to-end leg failed on %d other rank(s)" % int(c[3])
    else:
        ok = ok and e2e_ok

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = n + rbsp_bytes + 20 * n_nals          # N_in + N_rbsp + index
        achieved = alg_bytes / (t_scan_ms * 1e-3) / 1e9
        value = bins_all * args.steps / (t_max_ms * 1e-3)
        out = {
            "metric": "cabac_bins_per_s", "value": value, "unit": "bins/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic (harness GPU encoder, %.1f s)" % t_gen,
            "config": config_dict(args), "results_verified": bool(ok),
            "annexb_gbps": bytes_all / 1e9 / (t_scan_ms * 1e-3),
            "stage_ms": {"annexb_scan": t_scan_ms, "slice_select+cabac": t_cabac_ms},
            "stream_bytes_per_gpu": n, "bins_per_gpu": total_bins, "nals_per_gpu": n_nals,
            "roofline": {"bound": "hbm", "kernel": "annexb_copy_kernel + the six small launches of one Annex-B pass "
                                                   "(memset, dirty chunks, ordinal scan x2, permute, finalize, fixup), "
                                                   "timed together with CUDA events on the launching stream",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_PER_ALG_BYTE * alg_bytes if TRAFFIC_PER_ALG_BYTE else None,
                         "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src, "algorithmic_bytes": alg_bytes},
            "roofline_cabac": {"bound": "issue/latency (serial integer chain; not HBM, not tensor)",
                               "bins_per_s_per_gpu": total_bins / (t_cabac_ms * 1e-3),
                               "lanes": n_slices, "hbm_gbs_implied": total_bins * 0.235 / (t_cabac_ms * 1e-3) / 1e9,
                               # issue model: warp instructions per op from ncu (smsp__inst_executed.sum / warp-ops,
                               # profiles/r1_ncu_cabac_final2_summary.txt); one scheduler issues <= 1 per cycle and the
                               # ALU pipe most of these instructions use takes 2 cycles per warp instruction
                               "warp_inst_per_bin": CABAC_WARP_INST_PER_OP,
                               "issue_ipc_per_scheduler": (total_bins / 32.0) * CABAC_WARP_INST_PER_OP / (
                                   sm_count * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6 * t_cabac_ms * 1e-3),
                               "alu_pipe_ipc_peak": 0.5,
                               "equal_length_bins_per_s": 665e9,
                               "note": "bounded by its longest bundle: 1.92 x mean ops x ~157 cycles for a warp on its own "
                                       "(DESIGN.md section 4, K3; tools/cabac_balance_exp.py)"},
            "e2e": {"value": (bins_all / t_e2e_max) if e2e_error is None else None, "error": e2e_error,
                    "unit": "bins/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": t_e2e_max * 1e3, "steps": e2e_steps,
                    "api": "h264b_stream_submit / h264b_stream_wait, three jobs in flight (pinned host stream in; NAL index, "
                           "packed bins, final states out; copies of consecutive steps overlap kernels)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            # bounded sample: ~10-20 s of CPU work in all (all-core pass on `ns` slices, one-core pass on 1/8 of them)
            ns = args.cpu_slices or max(128 * cores, 512)
            sample = make_cpu_sample(ns)
            small = make_cpu_sample(max(ns // 8, 16)) if args.cpu_single else None
            b1, s1, s1_scan, s1_cabac = cpu_reference_pass(small, 1) if args.cpu_single else (None, None, None, None)
            bN, sN, sN_scan, sN_cabac = cpu_reference_pass(sample, cores)
            out["cpu_baseline"] = {
                "value": bN / sN, "unit": "bins/s", "cores": cores, "kind": "port",
                "sample": "%d slices of the same generator (%.1f MB Annex-B, %d bins), literal oracle: scan %.2f s "
                          "(1 thread), CABAC %.2f s (%d threads)" % (ns, len(sample["stream"]) / 1e6, bN, sN_scan,
                                                                     sN_cabac, cores),
                "scan_gbps_1core": len(sample["stream"]) / 1e9 / sN_scan,
                "cabac_bins_per_s_all_cores": bN / sN_cabac,
            }
            if b1:
                out["cpu_baseline"]["single_core_bins_per_s"] = b1 / s1
                out["cpu_baseline"]["cabac_bins_per_s_1core"] = b1 / s1_cabac
        print(json.dumps(out))
    if h_stream is not None:
        ctx.host_free(h_stream)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def _stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags):
    """h264b_stream_submit on buffers that stay alive in the caller"""
    import ctypes as C
    j = capi.StreamJob()
    j.stream = h_stream.ctypes.data
    j.n = len(h_stream)
    j.slice_data_offset = 0
    j.n_ctx = N_CTX
    j.ops = ops.ctypes.data
    j.n_ops_max = len(ops)
    j.n_ops = n_ops.ctypes.data
    j.qp = p.ctypes.data
    j.max_slices = len(p)
    j.flags = flags
    t = C.c_uint64()
    ctx._check(capi.lib().h264b_stream_submit(ctx.h, C.byref(j), C.byref(t)))
    return t.value


def _stream_wait_raw(ctx, capi, ticket):
    """h264b_stream_wait without copying the (multi-GB) results out of the library's pinned buffers again"""
    import ctypes as C
    r = capi.StreamResult()
    ctx._check(capi.lib().h264b_stream_wait(ctx.h, ticket, C.byref(r)))
    return {"n_slices": r.n_slices, "total_bins": r.total_bins, "n_nals": r.scan.n_nals}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=10000, help="frames per stream (10000 = the ~4 GB config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-single", action="store_true", default=True)
    ap.add_argument("--cpu-slices", type=int, default=0, help="slices in the bounded CPU sample (default 16 x cores)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
