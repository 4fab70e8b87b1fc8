"""Host logic: the launch plan of the MP3 kernel (chains of tiles per persistent CTA) covers every granule of
every run exactly once, in order, with consistent state hand-over flags.  No GPU needed."""
import ctypes

import numpy as np
import pytest

import symphonia_b200 as sb
from symphonia_b200._native import MP3_RUN_DTYPE

TILE_DTYPE = np.dtype([("first_frame", "<u4"), ("stream", "<u4"), ("first_gr", "<u2"), ("n_granules", "<u2"),
                       ("gpf", "u1"), ("n_ch", "u1"), ("flags", "u1"), ("pad", "u1")])
LOAD, STORE, CARRY_IN, CARRY_OUT, GROUP_END = 1, 2, 4, 8, 16
GROUP_TILES, GROUP_REGIONS = 8, 24
T, NW = 16, 16


def _plan(runs, n_frames, n_streams, grid):
    lib = sb.lib()
    fn = lib.symgpu_debug_mp3_plan
    fn.restype = ctypes.c_size_t
    fn.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                   ctypes.c_size_t] + [ctypes.POINTER(ctypes.c_int)] * 3
    runs = np.ascontiguousarray(runs, dtype=MP3_RUN_DTYPE)
    n_ctas, n_tiles, hdr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    n = fn(grid, n_streams, runs.ctypes.data, len(runs), n_frames, None, 0, n_ctas, n_tiles, hdr)
    assert n > 0
    buf = np.zeros(n, dtype=TILE_DTYPE)
    fn(grid, n_streams, runs.ctypes.data, len(runs), n_frames, buf.ctypes.data, n, n_ctas, n_tiles, hdr)
    first = buf[:hdr.value].view(np.uint32)[:n_ctas.value + 1]
    tiles = buf[hdr.value:hdr.value + n_tiles.value]
    assert hdr.value + n_tiles.value == n
    return first, tiles


def _check(runs, n_frames, n_streams, grid):
    first, tiles = _plan(runs, n_frames, n_streams, grid)
    assert first[0] == 0 and first[-1] == len(tiles) and (np.diff(first.astype(np.int64)) >= 0).all()
    assert len(first) - 1 <= grid
    # position of every run's granules
    run_of_frame = {}
    for r in runs:
        if r["n_frames"]:
            run_of_frame[int(r["first_frame"])] = r
    progress = {}  # run first_frame -> next granule expected
    chain_of_tile = np.repeat(np.arange(len(first) - 1), np.diff(first.astype(np.int64)))
    prev = None
    for i, t in enumerate(tiles):
        gpf = int(t["gpf"])
        g0 = int(t["first_frame"]) * gpf + int(t["first_gr"])
        # which run holds this granule
        key = max(k for k in run_of_frame if k <= t["first_frame"])
        run = run_of_frame[key]
        rg = int(run["granules_per_frame"] or 2)
        assert gpf == rg and t["stream"] == run["stream"] and t["n_ch"] == (run["channels"] or 2)
        q0 = g0 - key * gpf
        n = int(t["n_granules"])
        n_gran = int(run["n_frames"]) * gpf
        assert 1 <= n <= T and q0 + n <= n_gran
        assert progress.get(key, 0) == q0, "granules of a run are covered in order, without gaps or repeats"
        progress[key] = q0 + n
        fl = int(t["flags"])
        assert not (fl & LOAD and fl & CARRY_IN) and not (fl & STORE and fl & CARRY_OUT)
        assert bool(fl & LOAD) == (q0 == 0 and not fl & CARRY_IN) or fl & CARRY_IN
        if fl & LOAD:
            assert q0 == 0
        if fl & STORE:
            assert q0 + n == n_gran
        if q0 + n == n_gran:
            assert fl & STORE, "the tile that ends a run publishes the stream state"
        if not fl & (LOAD | CARRY_IN):
            assert q0 >= 2 and n + 2 <= NW, "a halo tile needs two earlier granules of its run and n + 2 warps"
        if fl & CARRY_IN:
            assert prev is not None and chain_of_tile[i] == chain_of_tile[i - 1]
            pt, pkey, pend = prev
            assert int(pt["flags"]) & CARRY_OUT and pkey == key and pend == q0
        if fl & CARRY_OUT:
            assert i + 1 < len(tiles) and chain_of_tile[i + 1] == chain_of_tile[i] and int(tiles[i + 1]["flags"]) & CARRY_IN
        prev = (t, key, q0 + n)
    # groups: what the CTA processes in one trip
    for c in range(len(first) - 1):
        jobs = regions = count = 0
        for i in range(int(first[c]), int(first[c + 1])):
            fl, n = int(tiles[i]["flags"]), int(tiles[i]["n_granules"])
            if count and fl & CARRY_IN:
                raise AssertionError("a tile that takes its state from the previous group must start a group")
            jobs += n + (0 if fl & (LOAD | CARRY_IN) else 2)
            regions += n + 1
            count += 1
            assert jobs <= NW and regions <= GROUP_REGIONS and count <= GROUP_TILES
            if fl & GROUP_END:
                jobs = regions = count = 0
            else:
                assert not fl & CARRY_OUT, "a tile that hands its state on must end its group"
                assert i + 1 < int(first[c + 1]), "the last tile of a chain ends a group"
    for key, run in run_of_frame.items():
        assert progress.get(key, 0) == int(run["n_frames"]) * int(run["granules_per_frame"] or 2)
    # load balance: no chain carries much more than its share
    per_chain = np.bincount(chain_of_tile, weights=tiles["n_granules"].astype(np.float64), minlength=len(first) - 1)
    total = per_chain.sum()
    assert per_chain.max() <= total / (len(first) - 1) + 2 * T + 2
    return first, tiles


def _runs(frames, gpf=None, ch=None, streams=None):
    runs = np.zeros(len(frames), dtype=MP3_RUN_DTYPE)
    runs["n_frames"] = frames
    runs["first_frame"] = np.concatenate([[0], np.cumsum(frames)[:-1]])
    runs["stream"] = np.arange(len(frames)) if streams is None else streams
    runs["granules_per_frame"] = 2 if gpf is None else gpf
    runs["channels"] = 2 if ch is None else ch
    return runs, int(np.sum(frames))


def test_bench_shape_has_one_halo_per_chain_at_most():
    runs, nf = _runs([128] * 64)
    first, tiles = _check(runs, nf, 64, 148)
    halo = ((tiles["flags"] & (LOAD | CARRY_IN)) == 0).sum()
    assert halo <= 148 and len(first) - 1 == 148
    assert tiles["n_granules"].mean() > 13.5


def test_single_frame_streams():
    runs, nf = _runs([1] * 8192)
    first, tiles = _check(runs, nf, 8192, 148)
    assert len(tiles) == 8192 and ((tiles["flags"] & 15) == (LOAD | STORE)).all()
    assert ((tiles["flags"] & GROUP_END) != 0).sum() <= 8192 // 8 + 2 * 148  # eight single-frame runs per trip


@pytest.mark.parametrize("seed", range(12))
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
    for grid in (1, 2, 7, 148, 296):
        _check(runs, nf, n_runs, grid)


def test_long_single_stream():
    runs, nf = _runs([5000])
    first, tiles = _check(runs, nf, 1, 148)
    assert len(first) - 1 == 148
    assert ((tiles["flags"] & (LOAD | CARRY_IN)) == 0).sum() == 147  # every chain but the first starts with a halo
