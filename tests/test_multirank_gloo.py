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
