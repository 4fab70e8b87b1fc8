"""Known-answer checks for the AAC and Vorbis oracles.  The reference ships no vectors for these
stages, so they are pinned against the defining mathematics in f64:

  * AAC: a signal analysed with the standard's window sequences (forward MDCT in f64) must be
    reconstructed by Dsp::synth's IMDCT + window + overlap-add (TDAC), for every window sequence
    and both window shapes, to 1e-5;  TNS against a direct all-pole recursion in f64.
  * Vorbis: floor-1 rendering against an independent transcription of the Vorbis I specification's
    render_point / render_line; windows against the spec formula; synthesis via TDAC over a
    long/short block mix.
"""
import ctypes

import numpy as np

from tests._codec_helpers import (AAC_TNS, VORBIS_RUN, VORBIS_STREAM, VORBIS_UNIT, make_floor1_setup, mdct_forward)
from tests._oracle import ptr

ONLY_LONG, LONG_START, EIGHT_SHORT, LONG_STOP = 0, 1, 2, 3


def _aac_windows(oracle):
    get = oracle.oracle_aac_window
    get.restype = ctypes.POINTER(ctypes.c_float)
    w = {}
    for kbd in (0, 1):
        w[kbd, 0] = np.ctypeslib.as_array(get(kbd, 0), shape=(1024,)).astype(np.float64)
        w[kbd, 1] = np.ctypeslib.as_array(get(kbd, 1), shape=(128,)).astype(np.float64)
    return w


def test_aac_windows_are_power_complementary(oracle):
    w = _aac_windows(oracle)
    for key, v in w.items():
        assert np.abs(v ** 2 + v[::-1] ** 2 - 1.0).max() < 2e-6, key   # Princen-Bradley
    n = np.arange(1024)
    assert np.abs(w[0, 0] - np.sin((n + 0.5) * np.pi / 2048)).max() < 1e-6


def test_aac_filterbank_reconstructs_signal(oracle):
    w = _aac_windows(oracle)
    rng = np.random.default_rng(3)
    seqs = [ONLY_LONG, ONLY_LONG, LONG_START, EIGHT_SHORT, EIGHT_SHORT, LONG_STOP, ONLY_LONG, LONG_START, EIGHT_SHORT,
            LONG_STOP, ONLY_LONG]
    shapes = [0, 1, 1, 0, 1, 1, 0, 0, 0, 1, 1]
    n_f = len(seqs)
    sig = rng.standard_normal(1024 * (n_f + 1))
    delay = np.zeros(1024, dtype=np.float32)
    prev_shape = 0
    outs = []
    for f in range(n_f):
        blk = sig[1024 * f: 1024 * f + 2048]
        seq, shape = seqs[f], shapes[f]
        lw, sw, plw, psw = w[shape, 0], w[shape, 1], w[prev_shape, 0], w[prev_shape, 1]
        coef = np.zeros(1024)
        if seq == EIGHT_SHORT:
            for k in range(8):
                win = np.concatenate([psw if k == 0 else sw, sw[::-1]])
                coef[128 * k: 128 * k + 128] = mdct_forward(blk[448 + 128 * k: 448 + 128 * k + 256] * win, 128)
        else:
            left = plw.copy()
            right = lw[::-1].copy()
            if seq == LONG_START:
                right = np.concatenate([np.ones(448), sw[::-1], np.zeros(448)])
            if seq == LONG_STOP:
                left = np.concatenate([np.zeros(448), psw, np.ones(448)])
            coef = mdct_forward(blk * np.concatenate([left, right]), 1024)
        dst = np.zeros(1024, dtype=np.float32)
        c32 = coef.astype(np.float32)
        oracle.oracle_aac_synth(ptr(c32), ptr(delay), seq, shape, prev_shape, ptr(dst))
        outs.append(dst.copy())
        prev_shape = shape
    got = np.concatenate(outs[1:])                 # frame 0 has no left neighbour
    want = 0.25 * sig[1024: 1024 * n_f]            # (N / 2) * scale = 512 / 2048 = 64 / 256
    assert np.abs(got - want).max() < 1e-5 * 40


def test_aac_tns_matches_direct_recursion(oracle):
    rng = np.random.default_rng(5)
    for direction in (0, 1):
        c = rng.standard_normal(1024).astype(np.float32)
        f = np.zeros(1, dtype=AAC_TNS)
        f["start"], f["end"], f["order"], f["direction"] = 100, 400, 7, direction
        f["lpc"][0, :7] = (rng.standard_normal(7) * 0.2).astype(np.float32)
        got = c.copy()
        oracle.oracle_aac_tns(ptr(got), ptr(f), 1)
        ref = c.astype(np.float64)
        lpc = f["lpc"][0].astype(np.float64)
        idx = range(100, 400) if not direction else range(399, 99, -1)
        step = -1 if not direction else 1
        for m, i in enumerate(idx):
            for j in range(min(7, m)):
                ref[i] -= ref[i + step * (j + 1)] * lpc[j]
        assert np.abs(got - ref).max() < 1e-4
        assert (got[:100] == c[:100]).all() and (got[400:] == c[400:]).all()


# ---- Vorbis ---------------------------------------------------------------------------------------

def _spec_render_line(x0, y0, x1, y1, v, table):
    """Vorbis I specification, section 9.2.7 'render_line', transcribed independently."""
    dy, adx = y1 - y0, x1 - x0
    ady = abs(dy)
    base = int(dy / adx)  # truncation toward zero
    x, y, err = x0, y0, 0
    sy = base - 1 if dy < 0 else base + 1
    ady -= abs(base) * adx
    if x < len(v):
        v[x] = table[y]
    for x in range(x0 + 1, x1):
        err += ady
        if err >= adx:
            err -= adx
            y += sy
        else:
            y += base
        if x < len(v):
            v[x] = table[y]


def _spec_floor1(x_list, mult, floor_y, n, table):
    """Vorbis I specification, 7.2.4 'curve computation' step 1 and step 2."""
    rng_tab = {1: 256, 2: 128, 3: 86, 4: 64}
    rng_ = rng_tab[mult]
    cnt = len(x_list)
    final_y = [0] * cnt
    flag = [False] * cnt
    flag[0] = flag[1] = True
    final_y[0], final_y[1] = floor_y[0], floor_y[1]
    from tests._codec_helpers import find_neighbors
    for i in range(2, cnt):
        lo, hi = find_neighbors(x_list, i)
        dy = final_y[hi] - final_y[lo]
        adx = x_list[hi] - x_list[lo]
        off = (abs(dy) * (x_list[i] - x_list[lo])) // adx
        pred = final_y[lo] - off if dy < 0 else final_y[lo] + off
        val = floor_y[i]
        highroom, lowroom = rng_ - pred, pred
        room = highroom * 2 if highroom < lowroom else lowroom * 2
        if val:
            flag[lo] = flag[hi] = flag[i] = True
            if val >= room:
                final_y[i] = val - lowroom + pred if highroom > lowroom else pred - val + highroom - 1
            else:
                final_y[i] = pred - (val + 1) // 2 if val % 2 else pred + val // 2
        else:
            flag[i] = False
            final_y[i] = pred
    order = sorted(range(cnt), key=lambda k: x_list[k])
    v = [0.0] * n
    hx, hy, lx = 0, 0, 0
    ly = min(max(final_y[order[0]] * mult, 0), 255)
    for i in order[1:]:
        if flag[i]:
            hy = min(max(final_y[i] * mult, 0), 255)
            hx = x_list[i]
            _spec_render_line(lx, ly, hx, hy, v, table)
            lx, ly = hx, hy
    if hx < n:
        _spec_render_line(hx, hy, n, hy, v, table)
    return np.array(v, dtype=np.float32)


def _random_floor(rng, n2, mult):
    n_posts = int(rng.integers(8, 40))
    xs = [0, n2] + sorted(rng.choice(np.arange(1, n2), size=n_posts - 2, replace=False).tolist(), key=lambda _: rng.random())
    rng_ = {1: 256, 2: 128, 3: 86, 4: 64}[mult]
    ys = [int(rng.integers(0, rng_)), int(rng.integers(0, rng_))] + [int(v) for v in rng.integers(0, 24, size=n_posts - 2)]
    for k in range(2, n_posts):
        if rng.random() < 0.3:
            ys[k] = 0
    return xs, ys


def test_vorbis_floor1_matches_specification(oracle):
    oracle.oracle_vorbis_inverse_db.restype = ctypes.c_float
    table = [np.float32(oracle.oracle_vorbis_inverse_db(i)) for i in range(256)]
    assert table[255] == 1.0 and abs(table[0] - 1.0649863e-07) < 1e-13
    rng = np.random.default_rng(11)
    for trial in range(60):
        n2 = int(rng.choice([128, 1024]))
        mult = int(rng.integers(1, 5))
        xs, ys = _random_floor(rng, n2, mult)
        setup = make_floor1_setup(xs, mult)
        fy = np.zeros(65, dtype=np.uint16)
        fy[:len(ys)] = ys
        got = np.zeros(n2, dtype=np.float32)
        oracle.oracle_vorbis_floor1(ptr(np.array(setup)), ptr(fy), n2, ptr(got))
        want = _spec_floor1(xs, mult, ys, n2, table)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), trial


def test_vorbis_window_formula(oracle):
    oracle.oracle_vorbis_window.restype = ctypes.POINTER(ctypes.c_float)
    for bs in (64, 256, 2048, 8192):
        w = np.ctypeslib.as_array(oracle.oracle_vorbis_window(bs), shape=(bs // 2,))
        i = np.arange(bs // 2)
        want = np.sin(np.pi / 2 * np.sin(np.pi / 2 * (i + 0.5) / (bs // 2)) ** 2)
        assert np.abs(w - want).max() < 1e-7
        assert np.abs(w.astype(np.float64) ** 2 + w[::-1].astype(np.float64) ** 2 - 1).max() < 1e-6


def test_vorbis_synthesis_reconstructs_signal(oracle):
    """Long/short block mix through floor*residue -> IMDCT -> window -> overlap-add (dsp.rs:68-126)."""
    from tests._oracle import VorbisState
    oracle.oracle_vorbis_window.restype = ctypes.POINTER(ctypes.c_float)
    oracle.oracle_vorbis_inverse_db.restype = ctypes.c_float
    bs0, bs1 = 256, 2048
    w0 = np.ctypeslib.as_array(oracle.oracle_vorbis_window(bs0), shape=(bs0 // 2,)).astype(np.float64)
    w1 = np.ctypeslib.as_array(oracle.oracle_vorbis_window(bs1), shape=(bs1 // 2,)).astype(np.float64)
    flags = [1, 1, 0, 0, 0, 1, 0, 1, 1]
    rng = np.random.default_rng(9)
    # block centres advance by (prev_n + n) / 4 (lib.rs:300-305)
    centres, c = [], bs1
    for k, f in enumerate(flags):
        n = bs1 if f else bs0
        if k:
            pn = bs1 if flags[k - 1] else bs0
            c += (pn + n) // 4
        centres.append(c)
    sig = rng.standard_normal(centres[-1] + bs1)
    streams = np.zeros(1, dtype=VORBIS_STREAM)
    streams["bs0_exp"], streams["bs1_exp"], streams["channels"], streams["coupled"] = 8, 11, 1, 0
    setup = make_floor1_setup([0, 1024], 1)
    Y = 200
    cval = float(oracle.oracle_vorbis_inverse_db(Y))
    units = np.zeros(len(flags), dtype=VORBIS_UNIT)
    slot = bs1 // 2
    residue = np.zeros((len(flags), 2, slot), dtype=np.float32)
    fy = np.zeros((len(flags), 2, 65), dtype=np.uint16)
    fy[:, :, 0] = fy[:, :, 1] = Y
    for k, f in enumerate(flags):
        n = bs1 if f else bs0
        pf = flags[k - 1] if k else f
        nf = flags[k + 1] if k + 1 < len(flags) else f
        lw = (w1 if (f and pf) else w0)
        rw = (w1 if (f and nf) else w0)
        win = np.zeros(n)
        lo = n // 4 - len(lw) // 2
        win[lo: lo + len(lw)] = lw
        win[lo + len(lw): n // 2] = 1.0
        ro = 3 * n // 4 - len(rw) // 2
        win[n // 2: ro] = 1.0
        win[ro: ro + len(rw)] = rw[::-1]
        blk = sig[centres[k] - n // 2: centres[k] + n // 2] * win
        X = mdct_forward(blk, n // 2) / (n // 4)   # TDAC gain of an unscaled N-line IMDCT is N / 2
        residue[k, 0, : n // 2] = (X / cval).astype(np.float32)
        units[k]["block_flag"], units[k]["prev_block_flag"] = f, pf
        units[k]["floor"] = [0, 0xFFFF]
        units[k]["do_not_decode"] = [0, 1]
    runs = np.zeros(1, dtype=VORBIS_RUN)
    runs["n_packets"] = len(flags)
    st = (VorbisState * 1)()
    pcm = np.zeros((len(flags), 2, slot), dtype=np.float32)
    oracle.oracle_vorbis_batch(ctypes.byref(st), ptr(streams), ptr(np.array([setup])), ptr(units), ptr(fy), ptr(residue),
                               ptr(runs), 1, slot, ptr(pcm), 1)
    for k in range(1, len(flags)):
        n, pn = (bs1 if flags[k] else bs0), (bs1 if flags[k - 1] else bs0)
        length = (pn + n) // 4
        # output of packet k covers [centre(k-1), centre(k))
        want = sig[centres[k - 1]: centres[k]]
        assert len(want) == length
        assert np.abs(pcm[k, 0, :length] - want).max() < 2e-4, k
