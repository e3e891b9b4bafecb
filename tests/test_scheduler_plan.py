"""h264b_scheduler_plan: the host-side decisions of h264b_scheduler_run (which device a stream goes to, which pass of
that device, which launch class a slice) -- the same code decides in both.  Host only: runs without a GPU.  The rules are
restated here in numpy and checked against the library on BASELINE configs[4]-shaped batches."""
import numpy as np
import pytest

from h264decode_b200 import capi, sharding

SC = np.array([0, 0, 0, 1], np.uint8)


def fake_stream(rng, n_bytes, junk_front=0, open_end=False):
    """n_bytes from the first to the end of the last start code (what the scheduler counts), any payload"""
    body = rng.integers(4, 256, max(n_bytes - 8, 1), dtype=np.uint8)
    s = np.concatenate([SC, body[:max(n_bytes - 8, 0)], SC])
    if junk_front:
        s = np.concatenate([rng.integers(4, 256, junk_front, dtype=np.uint8), s])
    if open_end:
        s = np.concatenate([s, np.array([0x65, 9, 9], np.uint8)])
    return s


def make_batch(seed, n_streams, per=16, scale=1.0):
    rng = np.random.default_rng(seed)
    size = 1024.0 * 2.0 ** (10.0 * rng.random((n_streams, per)) ** 3) * scale   # configs[4]: 1 KB * 2^(10 u^3)
    n_ops = np.maximum((size * 8 / 0.88).astype(np.int64), 32).astype(np.uint32)
    nbytes = np.maximum(size.sum(1).astype(np.int64) // 64, 16)   # (the planner only sees extents; keep the test light)
    streams = [fake_stream(rng, int(nbytes[i]), junk_front=int(i % 5 == 1) * 7, open_end=(i % 7 == 2)) for i in range(n_streams)]
    return streams, nbytes, n_ops


def expect_plan(nbytes, n_ops, per, nd, sm_count, group_bytes):
    n_streams = len(nbytes)
    # devices: longest first onto the least loaded one (ties: lowest device, lowest stream) == sharding.lpt_assign
    parts = sharding.lpt_assign(np.asarray(nbytes, np.int64), nd)
    dev = np.full(n_streams, -1, np.int32)
    for d, p in enumerate(parts):
        dev[list(p)] = d
    ops2 = n_ops.reshape(n_streams, per).astype(np.int64)
    longest = ops2.max(1)
    pas = np.zeros(n_streams, np.uint32)
    cls = np.full(n_streams * per, 255, np.uint8)
    for d in range(nd):
        my = np.flatnonzero(dev == d)
        if not len(my):
            continue
        share = int(nbytes[my].sum())
        order = my[np.argsort(-longest[my], kind="stable")]
        cut = share > group_bytes and len(my) >= 8
        acc, raw = 0, {}
        for k in order:
            raw[k] = 0 if not cut else (0 if acc * 12 < share else (1 if acc * 2 < share else 2))
            acc += int(nbytes[k])
        used = sorted(set(raw.values()))
        for k in order:
            pas[k] = used.index(raw[k])
        top = int(longest[my].max())
        budget = sm_count * 4 // 3
        for p in used:
            st = np.sort([k for k in order if raw[k] == p])
            rows = np.concatenate([np.arange(k * per, (k + 1) * per) for k in st])
            o = n_ops[rows].astype(np.int64)
            perm = np.argsort(-o, kind="stable")
            so = o[perm]
            ptop = int(so[0])
            excl = (top * 53e-6 - 0.1 * share / 1e6) / 74e-6
            excl = max(excl, 0.3 * top)
            thr = [int(excl), ptop // 2, ptop // 8, ptop // 32, ptop // 128]
            thr[1] = min(thr[1], thr[0])
            n_excl = min(int((so > thr[0]).sum()), budget)
            budget -= n_excl
            c = np.full(len(so), 5, np.uint8)
            c[:n_excl] = 0
            rest = np.arange(len(so)) >= n_excl
            for ci in (4, 3, 2, 1):
                c[rest & (so > thr[ci])] = ci
            cls[rows[perm]] = c
    return dev, pas, cls


@pytest.mark.parametrize("nd", [1, 2, 8])
@pytest.mark.parametrize("seed,n_streams,scale,group", [(1, 512, 1.0, 1 << 20), (2, 64, 0.3, 4096), (3, 40, 1.0, 1 << 40), (4, 7, 1.0, 1)])
def test_plan_follows_the_stated_rules(nd, seed, n_streams, scale, group):
    per = 16
    streams, nbytes, n_ops = make_batch(seed, n_streams, per, scale)
    dev, pas, cls = capi.scheduler_plan(streams, [per] * n_streams, n_ops.reshape(-1), int(n_ops.max()), nd, 148, group)
    edev, epas, ecls = expect_plan(nbytes, n_ops.reshape(-1), per, nd, 148, group)
    assert np.array_equal(dev, edev)
    assert np.array_equal(pas, epas)
    assert np.array_equal(cls, ecls)


def test_plan_properties_on_a_configs4_sized_batch():
    """4096 streams x 16 slices over 8 devices: bytes balanced to 1e-3, every device's longest slices in its first pass and
    in class 0, class 0 within its budget, short slices in the last classes"""
    per, nd = 16, 8
    streams, nbytes, n_ops = make_batch(4096, 4096, per, 1.0)
    flat = n_ops.reshape(-1)
    dev, pas, cls = capi.scheduler_plan(streams, [per] * 4096, flat, int(flat.max()), nd, 148, 1 << 16)
    assert set(dev) == set(range(nd))
    by = np.array([nbytes[dev == d].sum() for d in range(nd)], np.float64)
    assert by.max() / by.mean() < 1.001
    ops2 = n_ops.reshape(4096, per)
    for d in range(nd):
        my = np.flatnonzero(dev == d)
        k = my[np.argmax(ops2[my].max(1))]            # the stream with the device's longest slice
        assert pas[k] == 0
        row = k * per + int(np.argmax(ops2[k]))
        assert cls[row] == 0
        rows = (my[:, None] * per + np.arange(per)).reshape(-1)
        assert (cls[rows] == 0).sum() <= 148 * 4 // 3
        assert set(pas[my]) == {0, 1, 2}
        # the first pass is small, the second ends at one half of the bytes (up to one stream)
        b0, b1 = nbytes[my][pas[my] == 0].sum(), nbytes[my][pas[my] <= 1].sum()
        assert b0 <= by[d] / 12 + nbytes[my].max() and b1 <= by[d] / 2 + nbytes[my].max()
        # classes are ordered by length inside a pass
        for p in range(3):
            r = rows[np.repeat(pas[my] == p, per)]
            for c in range(5):
                lo, hi = flat[r][cls[r] == c], flat[r][cls[r] == c + 1]
                if len(lo) and len(hi):
                    assert lo.min() >= hi.max()
    assert (cls != 255).all()


def test_plan_leaves_streams_without_a_nal_unit_alone():
    rng = np.random.default_rng(9)
    streams = [fake_stream(rng, 4000), np.array([1, 2, 3], np.uint8), np.zeros(0, np.uint8), fake_stream(rng, 900)]
    n_ops = np.array([100, 5000, 7, 7, 7, 300], np.uint32)
    dev, pas, cls = capi.scheduler_plan(streams, [2, 1, 1, 2], n_ops, 5000, 2, 148)
    assert list(dev) == [0, -1, -1, 1]
    assert list(cls[2:4]) == [255, 255] and (cls[[0, 1, 4, 5]] != 255).all()
