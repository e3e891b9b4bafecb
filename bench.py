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

MEAN_BINS = 455_000      # ~50 KB of CABAC data per slice at ~0.88 bit/bin
N_ACTIVE = 64
IN_FLIGHT = 3            # H264B_STREAM_JOBS_IN_FLIGHT
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
def measured_constants():
    """Numbers this file quotes from ncu captures: profiles/r2_constants.json, written by tools/ncu_constants.py from
    the committed .ncu-rep summaries together with the commit they were measured at (no constants pasted here)."""
    p = os.path.join(ROOT, "profiles", "r2_constants.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


class StepBuffers:
    """Device buffers of one step in flight and the context (own CUDA stream) that runs it."""

    def __init__(self, torch, capi, dev, local_rank, n, nal_cap, n_slices, total_words):
        self.ctx = capi.Context(local_rank)
        self.stream = torch.cuda.Stream(device=dev)
        assert self.stream.cuda_stream != 0  # (a NULL handle would mean "the context's own stream" to h264b_set_stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.rbsp = torch.empty(n + 64, dtype=torch.uint8, device=dev)
        self.nals = torch.empty(nal_cap * 32, dtype=torch.uint8, device=dev)
        self.sum = torch.zeros(64, dtype=torch.uint8, device=dev)
        self.off = torch.empty(n_slices, dtype=torch.int64, device=dev)
        self.len = torch.empty(n_slices, dtype=torch.int32, device=dev)
        self.snal = torch.empty(n_slices, dtype=torch.int32, device=dev)
        self.ns = torch.zeros(4, dtype=torch.int32, device=dev)
        self.bins = torch.zeros(total_words, dtype=torch.int32, device=dev)
        self.fin = torch.empty(n_slices * 32, dtype=torch.uint8, device=dev)


def verify_slices(torch, capi, orc, buf, d_stream, which, ops, n_ops, qp, idc, boff, flags, threads):
    """Full comparison of the chosen slices with the oracle: the NAL unit's RBSP bytes (NewNalUnit on the unit's own
    stream bytes), every bin, the final (codIRange, codIOffset, bitsRead).  Returns the number of slices that agree."""
    from concurrent.futures import ThreadPoolExecutor
    nals = np.frombuffer(buf.nals.cpu().numpy().tobytes(), dtype=capi.NAL_DTYPE)
    snal = buf.snal.cpu().numpy()
    fin = np.frombuffer(buf.fin.cpu().numpy().tobytes(), dtype=capi.FINAL_DTYPE)
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    fo = orc.BYPASS_SPEC_OR if flags & capi.BYPASS_SPEC_OR else 0
    jobs = []
    for s in which:
        u = nals[snal[s]]
        unit = d_stream[int(u["start"]):int(u["start"]) + int(u["num_bytes"])].cpu().numpy()
        rb = buf.rbsp[int(u["rbsp_off"]):int(u["rbsp_off"]) + int(u["rbsp_len"])].cpu().numpy()
        bins = buf.bins[int(boff[s]):int(boff[s + 1])].cpu().numpy().view(np.uint32)
        jobs.append((int(s), unit, rb, bins))

    def check(job):
        s, unit, rb, bins = job
        st, _, orb = orc.new_nal_unit(unit)
        if st != orc.OK or orb != rb.tobytes():
            return False
        init = orc.ctx_init(qp[s:s + 1], idc[s:s + 1], N_CTX)[0]
        rc, obins, ofin, _ = orc.cabac_decode_slice(rb, np.concatenate([ops[:n_ops[s]], term]), init, fo)
        nb = int(n_ops[s]) + 1
        nw = (nb + 31) // 32
        got = bins[:nw].copy()
        if nb % 32:
            got[-1] &= np.uint32((1 << (nb % 32)) - 1)
        f = fin[s]
        return bool(rc == orc.OK and np.array_equal(got, obins[:nw]) and
                    (int(f["cod_i_range"]), int(f["cod_i_offset"]), int(f["bits_read"]), int(f["n_bins"])) ==
                    (ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"], ofin["n_bins"]))

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return int(sum(ex.map(check, jobs)))


def copy_probe(torch, dev, h2d_bytes, d2h_bytes, reps, barrier):
    """What this box can copy: plain pinned cudaMemcpyAsync of one step's bytes host -> device and device -> host at
    once (two streams), every rank at the same time.  The time per repetition is the floor of an end-to-end step."""
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode != "d2h":
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode != "h2d":
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        res[mode] = (time.perf_counter() - t0) / reps
    del h_in, h_out, d_in, d_out
    return res


def run_gpu(args, rank, world, local_rank):
    import torch
    import harness as hz
    from h264decode_b200 import capi
    from oracle import oracle as orc

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

    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
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
    d_ops = torch.from_numpy(ops.view(np.int16)).to(dev)
    d_nops = torch.from_numpy(n_ops.view(np.int32)).to(dev)
    p = capi.Context.slice_qp(qp, idc)
    d_qp = torch.from_numpy(p.view(np.int32).reshape(-1, 2).copy()).to(dev)
    boff = np.zeros(n_slices + 1, dtype=np.uint64)
    boff[1:] = np.cumsum((n_ops.astype(np.uint64) + 1 + 31) // 32)
    d_boff = torch.from_numpy(boff.view(np.int64)).to(dev)

    # ---- two steps' worth of device buffers: steps alternate between two contexts (two CUDA streams), so the tail of
    # one step's CABAC launch -- its longest slices, alone on their schedulers -- overlaps the next step's kernels
    bufs = [StepBuffers(torch, capi, dev, local_rank, n, nal_cap, n_slices, int(boff[-1])) for _ in range(2)]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(b, events=None):
        c, st = b.ctx, b.stream
        if events:
            events[0].record(st)
        c.annexb_scan_dev(d_stream.data_ptr(), n, b.rbsp.data_ptr(), b.nals.data_ptr(), None, nal_cap, b.sum.data_ptr(), 0)
        if events:
            events[1].record(st)
        c.slice_select_dev(b.nals.data_ptr(), b.sum.data_ptr(), nal_cap, 0, n_slices, b.off.data_ptr(), b.len.data_ptr(),
                           b.snal.data_ptr(), b.ns.data_ptr())
        c.cabac_decode_dev(bytes=b.rbsp.data_ptr(), total_bytes=n + 16, off=b.off.data_ptr(), len=b.len.data_ptr(),
                           n_slices=n_slices, n_ctx=N_CTX, ops=d_ops.data_ptr(), n_ops_max=len(ops),
                           n_ops=d_nops.data_ptr(), qp=d_qp.data_ptr(), init_states=None, bins=b.bins.data_ptr(),
                           bins_off=d_boff.data_ptr(), bins_stride_words=0, final=b.fin.data_ptr(), final_states=None,
                           flags=flags)
        if events:
            events[2].record(st)

    # ---- warm-up, then the results are checked outside the timed regions
    for k in range(max(args.warmup, 1)):
        step(bufs[k & 1])
    step(bufs[0])
    step(bufs[1])
    torch.cuda.synchronize()
    ok = True
    for b in bufs:   # whole-workload properties: counts, no overrun, and the terminate bin the encoder wrote last
        summ = np.frombuffer(b.sum.cpu().numpy().tobytes()[:48], dtype=np.uint64, count=5)
        fin = np.frombuffer(b.fin.cpu().numpy().tobytes(), dtype=capi.FINAL_DTYPE)
        ok = ok and (int(summ[1]) == n_nals and int(b.ns.cpu()[0]) == n_slices and
                     int(fin["n_bins"].astype(np.int64).sum()) == total_bins and not (fin["flags"] & capi.F_OVERRUN).any()
                     and np.array_equal(fin["n_bins"], n_ops + 1))
        last_word = b.bins[torch.from_numpy((boff[1:] - 1).astype(np.int64)).to(dev)].cpu().numpy().view(np.uint32)
        ok = ok and bool(np.all((last_word >> (n_ops & 31).astype(np.uint32)) & 1 == 1))
        rbsp_bytes = int(summ[2])
    # every bin, final engine state and RBSP byte of the 32 longest slices and of randomly chosen ones, against the oracle
    rng = np.random.default_rng(0x48323634 + rank)
    longest = np.argsort(n_ops)[::-1][:min(32, n_slices)]
    sample = rng.choice(n_slices, size=min(args.verify_slices, n_slices), replace=False)
    which = np.unique(np.concatenate([longest, sample]))
    t_ver = time.perf_counter()
    verified = verify_slices(torch, capi, orc, bufs[0], d_stream, which, ops, n_ops, qp, idc, boff, flags,
                             os.cpu_count() or 1) if len(which) else 0
    t_ver = time.perf_counter() - t_ver
    ok = ok and verified == len(which)
    ok = ok and bool(torch.equal(bufs[0].bins, bufs[1].bins) and torch.equal(bufs[0].fin, bufs[1].fin))

    # ---- timed region 1: exactly K steps, one after the other on one stream
    sampler = ClockSampler(local_rank)
    evs = [[ev(), ev(), ev()] for _ in range(args.steps)]
    launches0 = bufs[0].ctx.launch_count() + bufs[1].ctx.launch_count()
    barrier()
    sampler.start()
    e_beg, e_end = ev(), ev()
    e_beg.record(bufs[0].stream)
    for k in range(args.steps):
        step(bufs[0], evs[k])
    e_end.record(bufs[0].stream)
    torch.cuda.synchronize()
    barrier()
    t_serial_ms = e_beg.elapsed_time(e_end)
    launches = bufs[0].ctx.launch_count() + bufs[1].ctx.launch_count() - launches0
    t_scan_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    t_cabac_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    # ---- timed region 2: exactly K steps, two in flight (alternating contexts); timed on the device from the first
    # step's start to the end of whichever step ends last
    join = ev()
    barrier()
    o_beg, o_end = ev(), ev()
    o_beg.record(bufs[0].stream)
    bufs[1].stream.wait_event(o_beg)
    for k in range(args.steps):
        step(bufs[k & 1])
    join.record(bufs[1].stream)
    bufs[0].stream.wait_event(join)
    o_end.record(bufs[0].stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    t_overlap_ms = o_beg.elapsed_time(o_end)

    # ---- the EPB-dense stream of SURVEY.md 8(d) C1 (one emulation-prevention byte per ~200 bytes), 512 MiB: the
    # scan / strip kernels' other regime
    dense = None
    if rank == 0 and not args.no_dense:
        try:
            c1 = np.ascontiguousarray(hz.build_stream_c1(1 << 20), dtype=np.uint8)
            reps = 512
            d_c1 = torch.from_numpy(c1.copy()).to(dev).repeat(reps)
            nd = int(d_c1.numel())
            capd = nd // 64 + 4096
            db = StepBuffers(torch, capi, dev, local_rank, nd, capd, 1, 1)
            for _ in range(2):
                db.ctx.annexb_scan_dev(d_c1.data_ptr(), nd, db.rbsp.data_ptr(), db.nals.data_ptr(), None, capd, db.sum.data_ptr(), 0)
            d0, d1 = ev(), ev()
            d0.record(db.stream)
            for _ in range(3):
                db.ctx.annexb_scan_dev(d_c1.data_ptr(), nd, db.rbsp.data_ptr(), db.nals.data_ptr(), None, capd, db.sum.data_ptr(), 0)
            d1.record(db.stream)
            torch.cuda.synchronize()
            sd = np.frombuffer(db.sum.cpu().numpy().tobytes()[:48], dtype=np.uint64, count=5)
            dense = {"ms": d0.elapsed_time(d1) / 3, "bytes": nd, "n_nals": int(sd[1]), "rbsp_bytes": int(sd[2]), "n_epb": int(sd[4])}
            db.ctx.close()
            del db, d_c1
        except Exception as ex:
            dense = {"error": "%s: %s" % (type(ex).__name__, ex)}

    # ---- e2e through the host-buffer entry point
    ctx = bufs[0].ctx
    for b in bufs:
        del b.bins, b.rbsp
    torch.cuda.empty_cache()
    # (at least 24 steps: with three jobs in flight a short run is mostly pipeline fill and drain -- one H2D and one kernel
    #  stage before the first result, three D2H after the last submit; they stay inside the timed region and the value)
    e2e_steps = max(24, args.steps)
    done_at = []
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
                done_at.append(time.perf_counter())
            pending.append(_stream_submit_raw(ctx, capi, h_stream, ops, n_ops, p, flags))
        while pending:
            r = _stream_wait_raw(ctx, capi, pending.pop(0))
            done_at.append(time.perf_counter())
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
    if h_stream is not None:
        ctx.host_free(h_stream)
        h_stream = None
    probe = None
    if not args.no_probe:
        try:
            probe = copy_probe(torch, dev, h2d_bytes, d2h_bytes, 3, barrier)
        except Exception as ex:
            probe = {"error": "%s: %s" % (type(ex).__name__, ex)}
            if dist is not None:
                try:
                    barrier()
                except Exception:
                    pass

    # ---- reduce over ranks
    per_rank = None
    t_ser_max, t_ovl_max, t_e2e_max, bins_all, bytes_all = t_serial_ms, t_overlap_ms, t_e2e, total_bins, n
    probe_max = dict(probe) if probe and "error" not in probe else None
    if dist is not None:
        t = torch.tensor([t_serial_ms, t_overlap_ms, t_e2e] + ([probe[k] for k in ("h2d", "d2h", "both")] if probe_max else [0, 0, 0]),
                         dtype=torch.float64, device=dev)
        mine = torch.tensor([t_serial_ms / args.steps, t_overlap_ms / args.steps, t_cabac_ms, float(n_ops.max()),
                             float(total_bins)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[float(x) for x in r_] for r_ in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ser_max, t_ovl_max, t_e2e_max = float(t[0]), float(t[1]), float(t[2])
        if probe_max:
            probe_max = {"h2d": float(t[3]), "d2h": float(t[4]), "both": float(t[5])}
        c = torch.tensor([total_bins, n, int(ok and e2e_ok), int(e2e_error is not None), verified], dtype=torch.int64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        bins_all, bytes_all = int(c[0]), int(c[1])
        ok = int(c[2]) == world
        verified = int(c[4])
        if int(c[3]) and e2e_error is None:
            e2e_error = "the end-to-end leg failed on %d other rank(s)" % int(c[3])
    else:
        ok = ok and e2e_ok

    if rank == 0:
        peak, peak_src = measured_peaks()
        k = measured_constants()
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        alg_bytes = n + rbsp_bytes + 20 * n_nals          # N_in + N_rbsp + index
        achieved = alg_bytes / (t_scan_ms * 1e-3) / 1e9
        value = bins_all * args.steps / (t_ovl_max * 1e-3)
        serial = bins_all * args.steps / (t_ser_max * 1e-3)
        cabac_rate = total_bins / (t_cabac_ms * 1e-3)
        inst_per_bin = k.get("cabac_warp_inst_per_bin")
        issue_bound = (sm_count * 4 * 32 * sm_mhz * 1e6 / inst_per_bin) if inst_per_bin else None
        out = {
            "metric": "cabac_bins_per_s", "value": value, "unit": "bins/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_ovl_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u32 integer", "data": "synthetic (harness GPU encoder, %.1f s)" % t_gen,
            "config": config_dict(args), "results_verified": bool(ok),
            "value_mode": "K steps on resident input, two in flight (two contexts = two CUDA streams; the tail of one "
                          "step's CABAC launch overlaps the next step's kernels), device-timed first start -> last end",
            "serialized": {"value": serial, "ms_per_step": t_ser_max / args.steps,
                           "note": "the same K steps one after the other on one stream"},
            "verified_slices": {"compared_with_oracle": int(verified), "chosen": int(len(which)) * world,
                                "what": "RBSP bytes (NewNalUnit on the unit's stream bytes), every bin, final codIRange / "
                                        "codIOffset / bitsRead; the 32 longest slices + %d random ones per rank" % args.verify_slices,
                                "seconds": t_ver},
            "annexb_gbps": bytes_all / 1e9 / (t_scan_ms * 1e-3),
            "stage_ms": {"annexb_scan": t_scan_ms, "slice_select+sort+assign+cabac": t_cabac_ms},
            "stream_bytes_per_gpu": n, "bins_per_gpu": total_bins, "nals_per_gpu": n_nals,
            # the dominant kernel (98 % of a step): cabac_decode_kernel.  Serial integer work: neither HBM nor tensor
            # bound; its roofline is the warp schedulers' issue rate
            "roofline": {"bound": "issue (integer pipes; not hbm, not tensor)", "kernel": "cabac_decode_kernel",
                         "achieved": cabac_rate / 1e9, "peak": issue_bound / 1e9 if issue_bound else None, "unit": "Gbins/s",
                         "frac": cabac_rate / issue_bound if issue_bound else None,
                         "peak_source": "SMs x 4 schedulers x 32 lanes x %.0f MHz / %s warp instructions per bin (ncu "
                                        "smsp__inst_executed.sum / warp-bins, %s)" % (sm_mhz, inst_per_bin, k.get("source")),
                         "traffic": k.get("cabac_dram_bytes_per_launch"), "algorithmic_bytes": int(total_bins * 0.235),
                         "hbm_frac": total_bins * 0.235 / (t_cabac_ms * 1e-3) / 1e9 / peak,
                         "longest_slice_ops": int(n_ops.max()), "mean_slice_ops": float(n_ops.mean()),
                         "constants": k},
            "roofline_scan": {"bound": "hbm", "kernel": "annexb_copy_kernel + the small launches of one Annex-B pass, timed "
                                                        "together with CUDA events on the launching stream",
                              "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "traffic": k.get("scan_dram_bytes_per_alg_byte", 0) * alg_bytes or None,
                              "peak_source": peak_src, "algorithmic_bytes": alg_bytes},
            "e2e": {"value": (bins_all / t_e2e_max) if e2e_error is None else None, "error": e2e_error,
                    "unit": "bins/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": t_e2e_max * 1e3, "steps": e2e_steps,
                    # rank 0's median interval between two results coming home: the pipeline once it is full
                    "steady_ms_per_step": (float(np.median(np.diff(done_at))) * 1e3) if len(done_at) > 4 else None,
                    "api": "h264b_stream_submit / h264b_stream_wait, three jobs in flight (pinned host stream in; NAL index, "
                           "packed bins, final states out; copies of consecutive steps overlap kernels)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if dense and "error" not in dense:
            d_alg = dense["bytes"] + dense["rbsp_bytes"] + 20 * dense["n_nals"]
            out["roofline_scan_dense"] = {"bound": "hbm", "workload": "512 copies of the configs[0] stream: %d EPBs in %.0f MB"
                                          % (dense["n_epb"], dense["bytes"] / 1e6), "ms": dense["ms"],
                                          "achieved": d_alg / (dense["ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": d_alg / (dense["ms"] * 1e-3) / 1e9 / peak, "algorithmic_bytes": d_alg}
        elif dense:
            out["roofline_scan_dense"] = dense
        if probe_max:
            out["e2e"]["copy_probe"] = {
                "what": "plain pinned cudaMemcpyAsync of one step's bytes, host -> device and device -> host at once, "
                        "all %d ranks at the same time (max over ranks)" % world,
                "h2d_gbs": h2d_bytes / probe_max["h2d"] / 1e9, "d2h_gbs": d2h_bytes / probe_max["d2h"] / 1e9,
                "both_ms": probe_max["both"] * 1e3,
                "e2e_roofline_bins_per_s": bins_all / probe_max["both"],
                "e2e_frac_of_copy_floor": (probe_max["both"] / t_e2e_max) if e2e_error is None and t_e2e_max else None}
        elif probe:
            out["e2e"]["copy_probe"] = probe
        if per_rank:
            out["per_rank"] = {"columns": ["serialized ms/step", "overlapped ms/step", "cabac stage ms", "longest slice ops",
                                           "bins"], "rows": per_rank}
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
            # how to read value / cpu_baseline.value: 70 % of the CPU pass is its one-thread byte-at-a-time scan
            out["cpu_baseline"]["gpu_over_cpu"] = {
                "whole_pass_all_cores": value / (bN / sN), "cabac_only_all_cores": cabac_rate / (bN / sN_cabac),
                "cabac_only_one_core": (cabac_rate / (b1 / s1_cabac)) if b1 else None}
        print(json.dumps(out))
    for b in bufs:
        b.ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ configs[4]
def run_config4(args):
    """BASELINE configs[4], "multi-camera batch": 4096 concurrent synthetic streams x 16 slices, slice sizes
    1 KB * 2^(10 u^3) (1 KB .. 1 MB, skewed towards small), through h264b_scheduler on --gpus devices
    of ONE process: LPT by bytes over the devices, device jobs with the longest slices first, three jobs in flight per
    device.  Prints makespan, per-slice completion percentiles (tail latency) and the devices' busy-time imbalance."""
    import torch
    import harness as hz
    from h264decode_b200 import capi
    from oracle import oracle as orc

    n_dev = args.gpus
    n_streams, per = args.streams, 16
    rs = np.random.default_rng(4096)
    size_bytes = 1024.0 * 2.0 ** (10.0 * rs.random((n_streams, per)) ** 3)
    nb_all = np.maximum((size_bytes.reshape(-1) * 8 / 0.88).astype(np.int64), 32).astype(np.uint32)   # ~0.88 bit per bin
    dev = "cuda:0"
    torch.cuda.set_device(0)
    streams, n_ops, qp, idc, ops = [], [], [], [], None
    t_gen = time.perf_counter()
    part = max(1, min(n_streams, 256))   # streams generated at a time (the generator's scratch is slices x longest slice)
    pre_len = len(hz.SC + hz.SPS_NAL + hz.SC + hz.PPS_NAL)
    sc = np.frombuffer(hz.SC, np.uint8)
    for k0 in range(0, n_streams, part):
        k1 = min(k0 + part, n_streams)
        nb = nb_all[k0 * per:k1 * per]
        g = hz.gpu_build_stream_cabac(torch, dev, len(nb), 0, config=5, n_active=N_ACTIVE, n_ctx=N_CTX, slices_per_frame=per,
                                      frames_per_params=1, id_base=k0 * per, n_bins=nb)
        torch.cuda.synchronize()
        h = g["stream"][:g["n"]].cpu().numpy()
        sizes = (g["payload_lens"] + 5).reshape(-1, per).sum(1) + pre_len
        at = np.concatenate([[0], np.cumsum(sizes)])
        for i in range(k1 - k0):
            streams.append(np.concatenate([h[at[i]:at[i + 1]], sc]))
        n_ops.append(g["n_ops"])
        qp.append(g["qp"])
        idc.append(g["idc"])
        if ops is None or len(g["ops"]) > len(ops):
            ops = g["ops"]
        del g
        torch.cuda.empty_cache()
    n_ops, qp, idc = np.concatenate(n_ops), np.concatenate(qp), np.concatenate(idc)
    t_gen = time.perf_counter() - t_gen
    total_bytes = int(sum(len(x) for x in streams))
    flags = capi.BYPASS_SPEC_OR | capi.CABAC_FINAL_TERMINATE
    sch = capi.Scheduler(list(range(n_dev)))
    runs = []
    try:
        for rnd in range(args.warmup if args.warmup < 2 else 1):   # buffer growth
            sch.run(streams, [per] * n_streams, ops, n_ops, qp, idc, N_CTX, flags=flags, group_bytes=args.group_mb << 20)
        for rnd in range(args.steps):
            r = sch.run(streams, [per] * n_streams, ops, n_ops, qp, idc, N_CTX, flags=flags, group_bytes=args.group_mb << 20)
            runs.append(r)
    finally:
        sch.close()
    r = min(runs, key=lambda x: x["makespan_ms"])
    fin = r["final"]
    ok = bool(np.array_equal(fin["n_bins"], n_ops + 1)) and not (fin["flags"] & capi.F_OVERRUN).any()
    last = np.array([int(r["bins"][s][-1]) for s in range(len(n_ops))], dtype=np.uint64)
    ok = ok and bool(np.all((last >> (n_ops & 31).astype(np.uint64)) & 1 == 1))
    # a sample of slices against the oracle, bin by bin (the oracle strips the stream itself)
    rng = np.random.default_rng(44)
    check = np.unique(np.concatenate([rng.choice(len(n_ops), 64, replace=False), np.argsort(n_ops)[-4:]]))
    term = np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)
    verified = 0
    for s in check:
        st, k = divmod(int(s), per)
        onal, orbsp = orc.read_nal_units_arrays(streams[st])
        sl = np.flatnonzero((onal["type"] == 1) | (onal["type"] == 5))[k]
        data = orbsp[onal["rbsp_off"][sl]:onal["rbsp_off"][sl] + onal["rbsp_len"][sl]]
        init = orc.ctx_init(qp[s:s + 1], idc[s:s + 1], N_CTX)[0]
        rc, obins, ofin, _ = orc.cabac_decode_slice(data, np.concatenate([ops[:n_ops[s]], term]), init, orc.BYPASS_SPEC_OR)
        nbits = int(n_ops[s]) + 1
        nw = (nbits + 31) // 32
        got = np.array(r["bins"][s][:nw], dtype=np.uint32)
        if nbits % 32:
            got[-1] &= np.uint32((1 << (nbits % 32)) - 1)
        verified += int(rc == orc.OK and np.array_equal(got, obins[:nw]) and
                        (int(fin[s]["cod_i_range"]), int(fin[s]["cod_i_offset"]), int(fin[s]["bits_read"])) ==
                        (ofin["codIRange"], ofin["codIOffset"], ofin["bitsRead"]))
    ok = ok and verified == len(check)
    done = np.sort(r["slice_done_ms"])
    busy = r["device_busy_ms"]
    total_bins = int(n_ops.astype(np.int64).sum()) + len(n_ops)
    out = {
        "metric": "cabac_bins_per_s", "value": total_bins / (r["makespan_ms"] * 1e-3), "unit": "bins/s", "n_gpus": n_dev,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["makespan_ms"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8/u32 integer",
        "data": "synthetic (harness GPU encoder, %.1f s)" % t_gen,
        "config": {"workload": "configs[4]: multi-camera batch, %d streams x %d slices, slice sizes 1 KB * 2^(10 u^3); one "
                               "process, h264b_scheduler over %d device(s): LPT by bytes, one split + strip pass and five "
                               "CABAC launches by slice length (longest first, side by side) per device; host buffers in "
                               "and out" % (n_streams, per, n_dev),
                   "streams": n_streams, "slices": int(len(n_ops)), "stream_bytes": total_bytes, "n_ctx": N_CTX,
                   "longest_slice_bins": int(n_ops.max()), "mean_slice_bins": float(n_ops.mean())},
        "results_verified": bool(ok), "verified_slices": {"compared_with_oracle": int(verified), "chosen": int(len(check))},
        "makespan_ms": r["makespan_ms"], "makespan_ms_all_runs": [x["makespan_ms"] for x in runs],
        "slice_completion_ms": {"p50": float(done[len(done) // 2]), "p90": float(done[int(len(done) * 0.9)]),
                                "p99": float(done[int(len(done) * 0.99)]), "max": float(done[-1])},
        "device_busy_ms": [float(x) for x in busy], "device_busy_imbalance_max_over_mean": float(busy.max() / busy.mean()),
        "device_bytes": [int(x) for x in r["device_bytes"]], "device_jobs": [int(x) for x in r["device_jobs"]],
        "lpt_imbalance_bytes_max_over_mean": float(r["device_bytes"].max() / r["device_bytes"].mean()),
        # (a warp on its own: 104 cycles per bin at 1965 MHz, configs[1] in profiles/r2_configs_measured.txt)
        "longest_slice_floor_ms": float(n_ops.max()) * 104.0 / 1.965e6,
        "e2e": {"value": total_bins / (r["makespan_ms"] * 1e-3), "unit": "bins/s", "h2d_bytes_per_step": total_bytes,
                "d2h_bytes_per_step": int(r["bins_off"][-1]) * 4 + len(n_ops) * 32},
    }
    print(json.dumps(out))


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
    ap.add_argument("--config", type=int, default=3, choices=[3, 4], help="BASELINE.json configs[] index: 3 (default, the "
                    "headline workload) or 4 (multi-camera batch through the in-process multi-GPU scheduler)")
    ap.add_argument("--streams", type=int, default=4096, help="configs[4]: number of streams")
    ap.add_argument("--group-mb", type=int, default=16, help="configs[4]: stream bytes per device job")
    ap.add_argument("--verify-slices", type=int, default=256, help="random slices per rank compared bin by bin with the oracle")
    ap.add_argument("--no-probe", action="store_true", help="skip the pinned-copy probe (the end-to-end floor of this box)")
    ap.add_argument("--no-dense", action="store_true", help="skip the EPB-dense Annex-B stream (roofline_scan_dense)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config == 4:   # one process drives all --gpus devices (under torchrun: rank 0 alone)
        if rank == 0:
            run_config4(args)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
