"""Host logic: the launch plan of the second-generation MP3 kernel (one SHARE of consecutive granules per warp,
mp3_kernel_v2.cu) covers every granule of every run exactly once, in order, with consistent state flags, balanced
shares, and cuts that prefer run boundaries.  No GPU needed."""
import ctypes

import numpy as np
import pytest

import symphonia_b200 as sb
from symphonia_b200._native import MP3_RUN_DTYPE

TILE_DTYPE = np.dtype([("first_frame", "<u4"), ("stream", "<u4"), ("first_gr", "<u2"), ("n_granules", "<u2"),
                       ("gpf", "u1"), ("n_ch", "u1"), ("flags", "u1"), ("pad", "u1")])
LOAD, STORE, CARRY_IN, CARRY_OUT = 1, 2, 4, 8


def _plan(runs, n_frames, n_streams, max_shares):
    lib = sb.lib()
    fn = lib.symgpu_debug_mp3_plan_v2
    fn.restype = ctypes.c_size_t
    fn.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                   ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_int)] * 3
    runs = np.ascontiguousarray(runs, dtype=MP3_RUN_DTYPE)
    n_shares, n_tiles, hdr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    n = fn(max_shares, n_streams, runs.ctypes.data, len(runs), n_frames, None, 0, n_shares, n_tiles, hdr)
    assert n > 0
    buf = np.zeros(n, dtype=TILE_DTYPE)
    fn(max_shares, n_streams, runs.ctypes.data, len(runs), n_frames, buf.ctypes.data, n, n_shares, n_tiles, hdr)
    first = buf[:hdr.value].view(np.uint32)[:n_shares.value + 1]
    tiles = buf[hdr.value:hdr.value + n_tiles.value]
    assert hdr.value + n_tiles.value == n
    return first, tiles


def _check(runs, n_frames, n_streams, max_shares):
    first, tiles = _plan(runs, n_frames, n_streams, max_shares)
    n_shares = len(first) - 1
    assert 1 <= n_shares <= max_shares
    assert first[0] == 0 and first[-1] == len(tiles) and (np.diff(first.astype(np.int64)) >= 0).all()
    run_of_frame = {int(r["first_frame"]): r for r in runs if r["n_frames"]}
    progress = {}
    share_of_tile = np.repeat(np.arange(n_shares), np.diff(first.astype(np.int64)))
    prev = None
    for i, t in enumerate(tiles):
        gpf = int(t["gpf"])
        key = max(k for k in run_of_frame if k <= t["first_frame"])
        run = run_of_frame[key]
        assert gpf == int(run["granules_per_frame"] or 2) and t["stream"] == run["stream"] and t["n_ch"] == (run["channels"] or 2)
        q0 = (int(t["first_frame"]) - key) * gpf + int(t["first_gr"])
        n = int(t["n_granules"])
        n_gran = int(run["n_frames"]) * gpf
        assert n >= 1 and q0 + n <= n_gran
        assert progress.get(key, 0) == q0, "granules of a run are covered in order, without gaps or repeats"
        progress[key] = q0 + n
        fl = int(t["flags"])
        assert not (fl & LOAD and fl & CARRY_IN) and not (fl & STORE and fl & CARRY_OUT)
        if fl & LOAD:
            assert q0 == 0
        if q0 == 0:
            assert fl & LOAD, "a segment that starts its run takes the stream state"
        assert bool(fl & STORE) == (q0 + n == n_gran), "exactly the segment that ends a run publishes the stream state"
        if not fl & (LOAD | CARRY_IN):
            assert q0 >= 2, "a halo recomputes two earlier granules of the run"
            assert i == first[share_of_tile[i]], "only the first segment of a share can start inside a run"
        if fl & CARRY_IN:
            pt, pkey, pend = prev
            assert share_of_tile[i] == share_of_tile[i - 1] and int(pt["flags"]) & CARRY_OUT and pkey == key and pend == q0
        if fl & CARRY_OUT:
            assert i + 1 < len(tiles) and share_of_tile[i + 1] == share_of_tile[i] and int(tiles[i + 1]["flags"]) & CARRY_IN
        prev = (t, key, q0 + n)
    for key, run in run_of_frame.items():
        assert progress.get(key, 0) == int(run["n_frames"]) * int(run["granules_per_frame"] or 2)
    per_share = np.bincount(share_of_tile, weights=tiles["n_granules"].astype(np.float64), minlength=n_shares)
    return first, tiles, per_share


def _runs(frames, gpf=None, ch=None, streams=None):
    runs = np.zeros(len(frames), dtype=MP3_RUN_DTYPE)
    runs["n_frames"] = frames
    runs["first_frame"] = np.concatenate([[0], np.cumsum(frames)[:-1]])
    runs["stream"] = np.arange(len(frames)) if streams is None else streams
    runs["granules_per_frame"] = 2 if gpf is None else gpf
    runs["channels"] = 2 if ch is None else ch
    return runs, int(np.sum(frames))


def test_bench_shape_is_balanced():
    runs, nf = _runs([128] * 64)
    first, tiles, per_share = _check(runs, nf, 64, 148 * 12)
    assert len(first) - 1 == 148 * 12
    ideal = 16384 / (148 * 12)
    assert per_share.min() >= 1 and per_share.max() <= ideal + 0.25 * ideal + 2
    halos = ((tiles["flags"] & (LOAD | CARRY_IN)) == 0).sum()
    assert halos <= 148 * 12 - 64 + 64  # at most one per share


def test_single_frame_streams_need_no_halo():
    runs, nf = _runs([1] * 8192)
    first, tiles, per_share = _check(runs, nf, 8192, 148 * 12)
    assert len(tiles) == 8192 and ((tiles["flags"] & 15) == (LOAD | STORE)).all()
    assert per_share.max() <= 12


def test_tiny_batches_use_few_shares():
    runs, nf = _runs([1])
    first, tiles, per_share = _check(runs, nf, 1, 148 * 12)
    assert len(first) - 1 == 1 and len(tiles) == 1 and tiles[0]["flags"] == (LOAD | STORE)
    runs, nf = _runs([10] * 4)
    first, tiles, per_share = _check(runs, nf, 4, 148 * 12)
    assert len(first) - 1 == 20  # 80 granules at >= 4 per share


@pytest.mark.parametrize("seed", range(16))
def test_random_run_sets(seed):
    rng = np.random.default_rng(seed)
    n_runs = int(rng.integers(1, 60))
    frames = rng.integers(0, 40, size=n_runs)
    if seed % 3 == 0:
        frames = rng.integers(0, 3, size=n_runs)
    if frames.sum() == 0:
        frames[0] = 1
    gpf = rng.integers(1, 3, size=n_runs)
    ch = rng.integers(1, 3, size=n_runs)
    runs, nf = _runs(frames, gpf, ch)
    for shares in (1, 2, 7, 148, 1776, 4000):
        _check(runs, nf, n_runs, shares)


def test_long_single_stream_is_cut_into_pieces_a_tile_can_count():
    runs, nf = _runs([5000])
    first, tiles, per_share = _check(runs, nf, 1, 148 * 12)
    assert len(first) - 1 == 1776
    assert ((tiles["flags"] & (LOAD | CARRY_IN)) == 0).sum() == 1775
    runs, nf = _runs([40000])  # 80000 granules on one warp: pieces of <= 32768 with the state kept in the warp
    first, tiles, per_share = _check(runs, nf, 1, 1)
    assert len(tiles) == 3 and [int(f) & 15 for f in tiles["flags"]] == [LOAD | CARRY_OUT, CARRY_IN | CARRY_OUT, CARRY_IN | STORE]
