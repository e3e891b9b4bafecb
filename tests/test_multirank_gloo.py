"""The N > 1 path on the CPU: world_size 2 over gloo.  Every rank computes the same LPT partition of a set of
independent streams, processes its share (here with the oracle standing in for the GPU: this test is about the host
logic -- partition, no data-path collective, MAX/SUM reduction of time and counters), and the reduced totals must equal
a single-process pass over all streams."""
import os

import numpy as np
import pytest

import harness as hz
from h264decode_b200 import sharding


def make_streams():
    """12 small multi-camera-style streams with skewed sizes (SURVEY.md §8d C5 shape, scaled down)"""
    out = []
    for i, n_slices in enumerate([1, 2, 9, 3, 1, 5, 2, 2, 7, 1, 4, 3]):
        b = hz.build_stream_cabac(n_slices, 600 + 150 * i, slices_per_frame=3, frames_per_params=2, id_base=1000 * i)
        out.append(b)
    return out


def process(b):
    """one rank's work on one stream: (NAL count, RBSP bytes, bins)"""
    from oracle import oracle as orc
    nal, rbsp = orc.read_nal_units_arrays(b["stream"])
    sl = np.flatnonzero((nal["type"] == 1) | (nal["type"] == 5))
    init = orc.ctx_init(b["qp"][:len(sl)], b["idc"][:len(sl)], b["n_ctx"])
    bins = 0
    for i, k in enumerate(sl):
        data = rbsp[nal["rbsp_off"][k]:nal["rbsp_off"][k] + nal["rbsp_len"][k]]
        ops = np.concatenate([b["ops"][:b["n_ops"][i]], np.array([orc.make_op(orc.OP_TERMINATE)], np.uint16)])
        rc, _, fin, _ = orc.cabac_decode_slice(data, ops, init[i], orc.BYPASS_SPEC_OR)
        assert rc == orc.OK
        bins += fin["n_bins"]
    return len(nal["start"]), len(rbsp), bins


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = make_streams()
    sizes = [len(b["stream"]) for b in streams]
    mine = sharding.my_share(sizes, rank, world)
    tot = np.zeros(3, np.int64)
    for i in mine:
        tot += np.array(process(streams[i]), np.int64)
    dist.barrier()
    seconds, counters = sharding.reduce_job(dist, "cpu", 0.5 + rank, list(tot) + [len(mine)])
    if rank == 0:
        q.put((seconds, counters, [int(i) for i in mine]))
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_partition_properties():
    rng = np.random.default_rng(1)
    for n_ranks in (1, 2, 3, 8):
        sizes = (1024 * 2 ** (10 * rng.random(200) ** 3)).astype(np.int64)  # 1 KB .. 1 MB, skewed (C5)
        parts = sharding.lpt_assign(sizes, n_ranks)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(sizes)))          # a partition: every stream exactly once
        assert sharding.imbalance(sizes, parts) < 1.0 + 1.0 * sizes.max() / (sizes.sum() / n_ranks)  # LPT bound
        again = sharding.lpt_assign(sizes, n_ranks)
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))  # deterministic on every rank
    assert [list(p) for p in sharding.lpt_assign([], 2)] == [[], []]


def test_two_ranks_gloo_match_single_process():
    import torch.multiprocessing as mp
    streams = make_streams()
    expect = np.zeros(3, np.int64)
    for b in streams:
        expect += np.array(process(b), np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    seconds, counters, mine0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert seconds == pytest.approx(1.5)                       # MAX over ranks
    assert counters[:3] == [int(x) for x in expect]            # SUM over ranks == single-process totals
    assert counters[3] == len(streams)
    sizes = [len(b["stream"]) for b in streams]
    assert mine0 == [int(i) for i in sharding.lpt_assign(sizes, 2)[0]]


# ------------------------------------------------------------------ one long stream cut into byte ranges (§8e)
def _nal_list(stream_bytes, base=0):
    """[(start in the whole stream, NumBytes, Type, RBSP bytes)] of one scan, by the oracle"""
    from oracle import oracle as orc
    nal, rbsp = orc.read_nal_units_arrays(stream_bytes)
    return [(int(nal["start"][k]) + base, int(nal["num_bytes"][k]), int(nal["type"][k]),
             bytes(rbsp[nal["rbsp_off"][k]:nal["rbsp_off"][k] + nal["rbsp_len"][k]])) for k in range(len(nal["start"]))]


def _random_stream(rng, n, p_zero, every):
    s = rng.integers(0, 256, n, dtype=np.uint8)
    s[rng.random(n) < p_zero] = 0
    for pos in rng.integers(0, max(1, n - 4), max(1, n // every)) if n >= 4 else []:
        s[pos:pos + 4] = [0, 0, 0, 1]
    return s


def test_byte_ranges_reproduce_the_whole_stream():
    rng = np.random.default_rng(11)
    cases = [np.zeros(0, np.uint8), np.array([0, 0, 0, 1], np.uint8), np.array([0, 0, 0, 1] * 6, np.uint8),
             rng.integers(2, 256, 3000, dtype=np.uint8)]                      # empty, bare start codes, no start code
    cases += [_random_stream(rng, int(rng.integers(1, 6000)), 0.3, 200) for _ in range(60)]
    cases += [_random_stream(rng, 4000, 0.9, 40) for _ in range(10)]          # zero runs around the cuts
    for s in cases:
        whole = _nal_list(s.tobytes())
        for n_shards in (1, 2, 3, 8):
            for inp in (s.tobytes(), s):
                ranges = sharding.cut_byte_ranges(inp, n_shards)
                assert len(ranges) == n_shards and ranges[0][0] == 0 and ranges[-1][1] == len(s)
                assert all(b <= e for b, e in ranges) and all(ranges[r][0] <= ranges[r + 1][0] for r in range(n_shards - 1))
                got = []
                for b, e in ranges:
                    got += _nal_list(s[b:e].tobytes(), b)
                assert got == whole


def _range_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = hz.build_stream_cabac(40, 500, slices_per_frame=4, frames_per_params=3, id_base=77)
    s = np.ascontiguousarray(b["stream"], np.uint8)
    lo, hi = sharding.cut_byte_ranges(s, world)[rank]
    mine = _nal_list(s[lo:hi].tobytes(), lo)
    dist.barrier()
    _, counters = sharding.reduce_job(dist, "cpu", 0.0, [len(mine), sum(len(x[3]) for x in mine),
                                                          sum(x[0] for x in mine)])
    if rank == 0:
        q.put((counters, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_byte_ranges_of_one_stream():
    import torch.multiprocessing as mp
    b = hz.build_stream_cabac(40, 500, slices_per_frame=4, frames_per_params=3, id_base=77)
    whole = _nal_list(np.ascontiguousarray(b["stream"], np.uint8).tobytes())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_range_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    counters, mine0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert counters == [len(whole), sum(len(x[3]) for x in whole), sum(x[0] for x in whole)]
    assert 0 < len(mine0) < len(whole) and mine0 == whole[:len(mine0)]      # rank 0 holds a proper prefix
