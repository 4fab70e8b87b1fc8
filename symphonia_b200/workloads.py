"""Seeded synthetic workloads: batches shaped like what the CPU entropy stage of each
reference decoder hands to the synthesis stage.  No audio files exist on the build or GPU box, so
every test and benchmark input comes from here; `data` in bench.py's JSON says "synthetic".
"""
import numpy as np

from . import _native
from ._native import (F_INTENSITY, F_MID_SIDE, F_MIXED, F_MPEG1, F_MUTE, F_PREFLAG, F_SCALEFAC_SCALE,
                      F_SFC_LSB, MP3_END, MP3_GC_DTYPE, MP3_LONG, MP3_RUN_DTYPE, MP3_SHORT, MP3_START)

SEED_BASE = 0x5EED0000


def mp3_batch(n_streams=64, frames_per_stream=128, seed=SEED_BASE + 1, sample_rate_idx=0, channels=2,
              joint=True, block_switching=True, pow43=None, first_frame_types=None):
    """MPEG Layer III batch: S streams x F consecutive frames, stream-major.

    Returns (units [S*F,2,2] MP3_GC_DTYPE, spectra [S*F,2,2,576] f32, runs [S] MP3_RUN_DTYPE).
    Spectrum values are sign * POW43[q] exactly as read_huffman_samples emits them
    (requantize.rs:128,:144) and exactly +0.0 from `rzero` on (requantize.rs:234).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    S, F = int(n_streams), int(frames_per_stream)
    mpeg1 = sample_rate_idx < 3
    gpf = 2 if mpeg1 else 1
    n = S * F
    if pow43 is None:
        pow43 = _native.mp3_pow43()
    units = np.zeros((n, 2, 2), dtype=MP3_GC_DTYPE)
    spectra = np.zeros((n, 2, 2, 576), dtype=np.float32)

    # ---- block types: Markov chain per stream over granules, both channels share the type -----
    G = F * gpf
    bt = np.zeros((S, G), dtype=np.uint8)
    cur = np.zeros(S, dtype=np.uint8)
    for g in range(G):
        if block_switching:
            u = rng.random(S)
            nxt = cur.copy()
            nxt[(cur == MP3_LONG) & (u < 0.1)] = MP3_START
            nxt[cur == MP3_START] = MP3_SHORT
            nxt[(cur == MP3_SHORT) & (u >= 0.5)] = MP3_END
            nxt[cur == MP3_END] = MP3_LONG
            cur = nxt
        bt[:, g] = cur
    mixed = (rng.random((S, G)) < 0.2) & (bt == MP3_SHORT)

    # ---- frame-level joint stereo mode: none .3, MS .5, MS+IS .15, IS .05 ---------------------
    u = rng.random((S, F))
    ms = ((u >= 0.3) & (u < 0.95)) if (joint and channels == 2) else np.zeros((S, F), bool)
    is_ = (u >= 0.8) if (joint and channels == 2) else np.zeros((S, F), bool)

    lam = np.linspace(40.0, 0.5, 576)  # mean quantised magnitude per line
    for gr in range(gpf):
        for ch in range(2):
            uu = units[:, gr, ch].reshape(S, F)
            if ch >= channels:
                units["flags"][:, gr, ch] = F_MUTE
                continue
            g_idx = np.arange(F) * gpf + gr
            btg = bt[:, g_idx]
            uu["block_type"] = btg
            uu["sample_rate_idx"] = sample_rate_idx
            uu["global_gain"] = rng.integers(120, 201, size=(S, F))
            flags = np.zeros((S, F), dtype=np.uint8)
            flags |= np.where(mixed[:, g_idx], F_MIXED, 0).astype(np.uint8)
            flags |= np.where(rng.random((S, F)) < 0.2, F_SCALEFAC_SCALE, 0).astype(np.uint8)
            flags |= np.where((rng.random((S, F)) < 0.2) & (btg != MP3_SHORT), F_PREFLAG, 0).astype(np.uint8)
            flags |= np.where(rng.random((S, F)) < 0.5, F_SFC_LSB, 0).astype(np.uint8)
            flags |= np.where(ms, F_MID_SIDE, 0).astype(np.uint8)
            flags |= np.where(is_, F_INTENSITY, 0).astype(np.uint8)
            if mpeg1:
                flags |= F_MPEG1
            uu["flags"] = flags
            uu["subblock_gain"] = rng.integers(0, 8, size=(S, F, 3))
            sf = rng.integers(0, 16, size=(S, F, 39)).astype(np.uint8)
            long_like = (btg != MP3_SHORT)[..., None]
            idx = np.arange(39)[None, None, :]
            sf = np.where(long_like & (idx >= 21), 0, sf)
            sf = np.where(~long_like & (idx >= 36), 0, sf)
            uu["scalefacs"] = sf
            # rzero: even, U{288..576}; with intensity stereo channel 1 ends early so that the top
            # bands really are intensity coded (stereo.rs:238-258).
            rz = 2 * rng.integers(144, 289, size=(S, F))
            if ch == 1:
                rz = np.where(is_, 2 * rng.integers(40, 200, size=(S, F)), rz)
            uu["rzero"] = rz
            q = np.minimum(8206, np.floor(rng.exponential(lam[None, None, :], size=(S, F, 576)))).astype(np.int64)
            sign = np.where(rng.random((S, F, 576)) < 0.5, -1.0, 1.0).astype(np.float32)
            val = sign * pow43[q]
            val = np.where(np.arange(576)[None, None, :] < rz[..., None], val, np.float32(0.0))
            val = np.where(q == 0, np.float32(0.0), val)  # x == 0 -> buf[i] = 0.0 (+0), requantize.rs:131-133
            spectra[:, gr, ch] = val.reshape(n, 576).astype(np.float32)
            units[:, gr, ch] = uu.reshape(n)
    if gpf == 1:
        units["flags"][:, 1, :] = F_MUTE
    runs = np.zeros(S, dtype=MP3_RUN_DTYPE)
    runs["stream"] = np.arange(S)
    runs["first_frame"] = np.arange(S) * F
    runs["n_frames"] = F
    runs["granules_per_frame"] = gpf
    runs["channels"] = channels
    return units, spectra, runs


def mpa12_batch(n_streams=64, frames_per_stream=128, layer=2, seed=SEED_BASE + 4, channels=2):
    """MPEG Layer I (12 time slots per frame) / Layer II (36) batch: dequantised, scaled sub-band samples as the layer
    decoders hand them to the synthesis bank (layer1/mod.rs:150-180, layer2/mod.rs:330-372): amplitudes falling with
    the sub-band index, the top sub-bands unallocated (exact zeros).
    Returns (subbands [S*F,2,32,n_slots] f32, runs [S] MPA12_RUN_DTYPE)."""
    from ._native import MPA12_RUN_DTYPE
    rng = np.random.Generator(np.random.PCG64(seed))
    S, F = int(n_streams), int(frames_per_stream)
    n_slots = 12 if layer == 1 else 36
    env = (2.0 ** (-np.arange(32) / 4.0))[None, None, :, None]
    x = (rng.standard_normal((S * F, 2, 32, n_slots)) * env).astype(np.float32)
    sblimit = rng.integers(8, 33, size=(S * F, 1, 1, 1))
    x = np.where(np.arange(32)[None, None, :, None] < sblimit, x, np.float32(0.0)).astype(np.float32)
    if channels == 1:
        x[:, 1] = 0.0
    runs = np.zeros(S, dtype=MPA12_RUN_DTYPE)
    runs["stream"] = np.arange(S)
    runs["first_frame"] = np.arange(S) * F
    runs["n_frames"] = F
    runs["channels"] = channels
    return x, runs


def flac_batch(n_frames=64, block_size=4096, seed=SEED_BASE + 5, bps=16, channels=2, return_pcm=False):
    """FLAC frames as the Rice stage leaves them: per sub-frame the warm-up samples followed by the residuals of an
    ENCODER written from the format definition (residual = sample - (sum(c[j] * s[i-1-j]) >> shift), exact Python /
    int64 arithmetic), so that restoring them must give back the PCM generated here.
    Returns (frames, subframes, samples [int32]) and, with return_pcm, the expected planes scaled to 32 bits."""
    from ._native import (FLAC_CONSTANT, FLAC_FIXED, FLAC_FRAME_DTYPE, FLAC_INDEPENDENT, FLAC_LEFT_SIDE, FLAC_LPC, FLAC_MID_SIDE,
                          FLAC_RIGHT_SIDE, FLAC_SUBFRAME_DTYPE, FLAC_VERBATIM)
    rng = np.random.Generator(np.random.PCG64(seed))
    frames = np.zeros(n_frames, dtype=FLAC_FRAME_DTYPE)
    subs = np.zeros(n_frames * channels, dtype=FLAC_SUBFRAME_DTYPE)
    samples = np.zeros(n_frames * channels * block_size, dtype=np.int32)
    expect = np.zeros_like(samples)
    fixed = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}
    t = np.arange(block_size)
    for f in range(n_frames):
        n = block_size if f % 7 else max(channels and 40, block_size // 3 + 1)  # some short blocks
        # 32-bit planes: mid + side = 2 * left must still fit the decoder's i32 arithmetic (decoder.rs:32-82 wraps), so one bit less
        amp = (1 << (bps - (3 if bps == 32 else 2))) - 1
        # a smooth signal plus noise, per channel, in `bps` bits
        pcm = [np.round(amp * 0.6 * np.sin(2 * np.pi * (rng.uniform(20, 2000) / 44100.0) * t[:n] + rng.uniform(0, 6))
                        + rng.normal(0, amp * 0.02, n)).astype(np.int64) for _ in range(channels)]
        assignment = int(rng.integers(0, 4)) if channels == 2 else FLAC_INDEPENDENT
        frames[f] = (f * channels, channels, assignment, bps, 0, (0, 0))
        if assignment == FLAC_LEFT_SIDE:
            planes = [pcm[0], pcm[0] - pcm[1]]
        elif assignment == FLAC_MID_SIDE:
            planes = [(pcm[0] + pcm[1]) >> 1, pcm[0] - pcm[1]]
        elif assignment == FLAC_RIGHT_SIDE:
            planes = [pcm[0] - pcm[1], pcm[1]]
        else:
            planes = pcm
        for c in range(channels):
            k = f * channels + c
            off = k * block_size
            x = planes[c].copy()
            kind = int(rng.choice([FLAC_LPC, FLAC_LPC, FLAC_LPC, FLAC_FIXED, FLAC_VERBATIM, FLAC_CONSTANT]))
            wasted = int(rng.choice([0, 0, 0, 1, 3]))
            if wasted:  # the encoder drops low zero bits: make them zero in the signal this plane carries
                x = (x >> wasted) << wasted
            if kind == FLAC_CONSTANT:
                x[:] = x[0]
            # what the other channel / the listener must end up with follows from the planes actually coded
            planes[c] = x
            y = x >> wasted
            sub = subs[k]
            sub["offset"], sub["n"], sub["type"], sub["wasted"] = off, n, kind, wasted
            if kind == FLAC_CONSTANT:
                samples[off] = y[0]
            elif kind == FLAC_VERBATIM:
                samples[off:off + n] = y
            else:
                if kind == FLAC_FIXED:
                    order = int(rng.integers(0, 5))
                    coeffs, shift = fixed[order], 0
                else:
                    order = int(rng.integers(1, 33)) if rng.random() < 0.3 else int(rng.integers(1, 13))
                    order = min(order, n)
                    prec, shift = int(rng.integers(5, 16)), int(rng.integers(0, 15))
                    coeffs = [int(v) for v in rng.integers(-(1 << (prec - 1)), 1 << (prec - 1), size=order)]
                    # keep the recursion stable-ish: scale so that sum |c| >> shift stays around 1
                    while sum(abs(v) for v in coeffs) >> shift > 2 and shift < 15:
                        shift += 1
                    sub["coeffs"][:order] = coeffs
                sub["order"], sub["shift"] = order, shift
                res = y.copy()
                for i in range(order, n):
                    pred = sum(coeffs[j] * int(y[i - 1 - j]) for j in range(order)) >> shift
                    res[i] = y[i] - pred
                samples[off:off + n] = ((res + (1 << 31)) % (1 << 32) - (1 << 31)).astype(np.int32)  # residuals as i32
        # expected output planes after decorrelation of the planes actually coded, scaled to 32 bits
        if assignment == FLAC_LEFT_SIDE:
            out = [planes[0], planes[0] - planes[1]]
        elif assignment == FLAC_MID_SIDE:
            mid = (planes[0] << 1) | (planes[1] & 1)
            out = [(mid + planes[1]) >> 1, (mid - planes[1]) >> 1]
        elif assignment == FLAC_RIGHT_SIDE:
            out = [planes[0] + planes[1], planes[1]]
        else:
            out = planes
        for c in range(channels):
            off = (f * channels + c) * block_size
            v = (out[c] << (32 - bps)) if bps < 32 else out[c]
            expect[off:off + n] = ((v + (1 << 31)) % (1 << 32) - (1 << 31)).astype(np.int32)
    return (frames, subs, samples, expect) if return_pcm else (frames, subs, samples)


def mp3_quantize(spectra, pow43=None):
    """Inverse of read_huffman_samples' table lookup: the int16 q with sign(q) * POW43[|q|] == spectra, exactly."""
    if pow43 is None:
        pow43 = _native.mp3_pow43()
    mag = np.abs(spectra)
    idx = np.searchsorted(pow43, mag)
    idx = np.minimum(idx, len(pow43) - 1)
    if not (pow43[idx] == mag).all():
        raise ValueError("spectra hold values that are not sign * POW43[q]")
    return np.where(np.signbit(spectra), -idx, idx).astype(np.int16)


def mp3_audio_seconds(n_frames, sample_rate_idx=0):
    rate = [44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000][sample_rate_idx]
    per_frame = 1152 if sample_rate_idx < 3 else 576
    return n_frames * per_frame / rate


MP3_ALGO_BYTES_PER_FRAME = 9216 + 256 + 9216  # spectra + descriptors + PCM


# =================================================================================================
# AAC-LC
# =================================================================================================
from ._native import (AAC_EIGHT_SHORT, AAC_LONG_START, AAC_LONG_STOP, AAC_ONLY_LONG, AAC_RUN_DTYPE, AAC_TNS_DTYPE,  # noqa: E402
                      AAC_UNIT_DTYPE, VORBIS_FLOOR1_DTYPE, VORBIS_RUN_DTYPE, VORBIS_STREAM_DTYPE, VORBIS_UNIT_DTYPE)


def _tns_lpc(rng, order, coef_res=True):
    """TNS LPC coefficients by the quantised-sine construction of aac/ics/tns.rs:66-101."""
    bits = 4 if coef_res else 3
    fac_base = 8.0 if coef_res else 4.0
    iqfac = np.float32((fac_base - 0.5) / (np.pi / 2))
    iqfac_m = np.float32((fac_base + 0.5) / (np.pi / 2))
    c = rng.integers(-(1 << (bits - 1)), 1 << (bits - 1), size=order).astype(np.float32)
    tmp = np.sin(np.where(c >= 0, c / iqfac, c / iqfac_m).astype(np.float32)).astype(np.float32)
    coef = np.zeros(21, dtype=np.float32)
    b = np.zeros(21, dtype=np.float32)
    for m in range(1, order + 1):
        for i in range(1, m):
            b[i] = coef[i - 1] + tmp[m - 1] * coef[m - i - 1]
        coef[: m - 1] = b[1:m]
        coef[m - 1] = tmp[m - 1]
    return coef[:20]


def aac_batch(n_streams=64, frames_per_stream=128, seed=SEED_BASE + 2, channels=2, tns_prob=0.2, block_switching=True):
    """AAC-LC batch: (units [S*F,2], tns [T], coeffs [S*F,2,1024], runs [S])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    S, F = int(n_streams), int(frames_per_stream)
    n = S * F
    units = np.zeros((S, F, 2), dtype=AAC_UNIT_DTYPE)
    k = np.arange(1024)
    env = (2.0 ** (-k / 128.0) * 2000.0) * (k < 672)
    coeffs = (rng.standard_normal((S, F, 2, 1024)) * env).astype(np.float32)
    coeffs[..., 672:] = 0.0
    tns = []
    for s in range(S):
        for ch in range(2):
            seq, prev_shape = AAC_ONLY_LONG, 0
            for f in range(F):
                if block_switching:
                    u = rng.random()
                    if seq == AAC_ONLY_LONG:
                        nxt = AAC_LONG_START if u < 0.1 else AAC_ONLY_LONG
                    elif seq == AAC_LONG_START:
                        nxt = AAC_EIGHT_SHORT
                    elif seq == AAC_EIGHT_SHORT:
                        nxt = AAC_EIGHT_SHORT if u < 0.5 else AAC_LONG_STOP
                    else:
                        nxt = AAC_ONLY_LONG
                    if f:
                        seq = nxt
                shape = int(rng.random() < 0.5)
                uu = units[s, f, ch]
                uu["window_sequence"], uu["window_shape"], uu["prev_window_shape"] = seq, shape, prev_shape
                prev_shape = shape
                if ch < channels and rng.random() < tns_prob:
                    first = len(tns)
                    if seq == AAC_EIGHT_SHORT:
                        for w in range(8):
                            if rng.random() < 0.5:
                                lo = int(rng.integers(0, 60))
                                hi = int(rng.integers(lo + 4, 112))
                                order = int(rng.integers(1, 8))
                                tns.append((w * 128 + lo, w * 128 + hi, order, int(rng.random() < 0.5), _tns_lpc(rng, order, False)))
                    else:
                        top = int(rng.integers(300, 672))
                        for _ in range(int(rng.integers(1, 4))):
                            bottom = max(0, top - int(rng.integers(40, 300)))
                            order = int(rng.integers(1, 13))
                            if top > bottom:
                                tns.append((bottom, top, order, int(rng.random() < 0.5), _tns_lpc(rng, order)))
                            top = bottom
                    uu["n_tns"], uu["tns_first"] = len(tns) - first, first
    if channels == 1:
        coeffs[:, :, 1] = 0.0
    tns_arr = np.zeros(len(tns), dtype=AAC_TNS_DTYPE)
    for i, (a, b, o, d, lpc) in enumerate(tns):
        tns_arr[i]["start"], tns_arr[i]["end"], tns_arr[i]["order"], tns_arr[i]["direction"] = a, b, o, d
        tns_arr[i]["lpc"] = lpc
    runs = np.zeros(S, dtype=AAC_RUN_DTYPE)
    runs["stream"] = np.arange(S)
    runs["first_frame"] = np.arange(S) * F
    runs["n_frames"] = F
    runs["channels"] = channels
    return units.reshape(n, 2), tns_arr, coeffs.reshape(n, 2, 1024), runs


AAC_ALGO_BYTES_PER_FRAME = 8192 + 32 + 8192  # SURVEY.md §8d (two 16-byte units)


# =================================================================================================
# Vorbis
# =================================================================================================
def find_neighbors(x_list, i):
    """low_neighbor / high_neighbor of the Vorbis I spec 9.2.4-9.2.5 (floor.rs:748-773)."""
    bound = x_list[i]
    low, high = 0, 0xFFFFFFFF
    res = [0, 0]
    for k in range(i):
        xv = x_list[k]
        if low < xv < bound:
            low, res[0] = xv, k
        if bound < xv < high:
            high, res[1] = xv, k
    return res


def make_floor1_setup(x_list, multiplier):
    """What Floor1::read_setup precomputes (floor.rs:546-560) for a given X list."""
    s = np.zeros((), dtype=VORBIS_FLOOR1_DTYPE)
    n = len(x_list)
    s["multiplier"] = multiplier
    s["n_posts"] = n
    s["x_list"][:n] = x_list
    for i in range(n):
        lo, hi = find_neighbors(list(x_list), i)
        s["low"][i], s["high"][i] = lo, hi
    s["sort_order"][:n] = sorted(range(n), key=lambda k: x_list[k])
    return s


def vorbis_batch(n_streams=64, packets_per_stream=128, seed=SEED_BASE + 3, bs_exp=(8, 11), channels=2, coupled=True,
                 unused_prob=0.05, posts=None):
    """Vorbis batch with a long/short block mix.

    Returns dict(streams, floors, units [P], floor_y [P,2,65], residue [P,2,slot], runs, slot, out_len [P])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    S, F = int(n_streams), int(packets_per_stream)
    bs0, bs1 = 1 << bs_exp[0], 1 << bs_exp[1]
    slot = bs1 // 2
    streams = np.zeros(S, dtype=VORBIS_STREAM_DTYPE)
    streams["bs0_exp"], streams["bs1_exp"], streams["channels"] = bs_exp[0], bs_exp[1], channels
    streams["coupled"] = 1 if (coupled and channels == 2) else 0
    # four floor configurations per block size: 20-40 posts, multiplier 1..4 (2 most common)
    floors, by_size = [], {0: [], 1: []}
    for flag, n2 in ((0, bs0 // 2), (1, bs1 // 2)):
        for _ in range(4):
            n_posts = int(min(rng.integers(20, 41), n2 // 2))
            if posts is not None:  # e.g. 65: the largest setup floor.rs:455-560 admits (post 64 needs its own step-2 flag)
                n_posts = int(min(posts, n2 // 2))
            inner = rng.choice(np.arange(1, n2), size=n_posts - 2, replace=False).tolist()
            rng.shuffle(inner)
            mult = int(rng.choice([1, 2, 2, 2, 3, 4]))
            by_size[flag].append(len(floors))
            floors.append(make_floor1_setup([0, n2] + inner, mult))
    floors = np.array(floors, dtype=VORBIS_FLOOR1_DTYPE)
    P = S * F
    units = np.zeros((S, F), dtype=VORBIS_UNIT_DTYPE)
    floor_y = np.zeros((S, F, 2, 65), dtype=np.uint16)
    residue = np.zeros((S, F, 2, slot), dtype=np.float32)
    out_len = np.zeros((S, F), dtype=np.int32)
    flag = (rng.random(S) < 0.8).astype(np.uint8)
    for f in range(F):
        if f:
            u = rng.random(S)
            flag = np.where(flag == 1, (u < 0.92).astype(np.uint8), (u >= 0.75).astype(np.uint8))
        prev = units["block_flag"][:, f - 1] if f else flag
        units["block_flag"][:, f] = flag
        units["prev_block_flag"][:, f] = prev
        out_len[:, f] = (np.where(prev == 1, bs1, bs0) + np.where(flag == 1, bs1, bs0)) // 4
        for s in range(S):
            n2 = (bs1 if flag[s] else bs0) // 2
            unused = [False, False]
            for ch in range(channels):
                fi = by_size[int(flag[s])][int(rng.integers(0, 4))]
                setup = floors[fi]
                npost = int(setup["n_posts"])
                rng_ = {1: 256, 2: 128, 3: 86, 4: 64}[int(setup["multiplier"])]
                y = rng.integers(0, 24, size=npost)
                y[rng.random(npost) < 0.3] = 0
                y[0], y[1] = rng.integers(0, rng_, size=2)
                floor_y[s, f, ch, :npost] = y
                unused[ch] = rng.random() < unused_prob
                units[s, f]["floor"][ch] = 0xFFFF if unused[ch] else fi
                r = rng.standard_normal(n2).astype(np.float32) * 4.0
                r[int(0.8 * n2):] = 0.0
                residue[s, f, ch, :n2] = r
            if channels == 1:
                units[s, f]["floor"][1] = 0xFFFF
                units[s, f]["do_not_decode"] = [int(unused[0]), 1]
            else:
                dnd = list(unused)
                if streams["coupled"][s] and dnd[0] != dnd[1]:  # non-zero vector propagate, lib.rs:215-225
                    dnd = [False, False]
                units[s, f]["do_not_decode"] = [int(dnd[0]), int(dnd[1])]
    runs = np.zeros(S, dtype=VORBIS_RUN_DTYPE)
    runs["stream"] = np.arange(S)
    runs["first_packet"] = np.arange(S) * F
    runs["n_packets"] = F
    return dict(streams=streams, floors=floors, units=units.reshape(P), floor_y=floor_y.reshape(P, 2, 65),
                residue=residue.reshape(P, 2, slot), runs=runs, slot=slot, out_len=out_len.reshape(P))


def vorbis_mc_batch(n_streams=4, packets_per_stream=24, seed=SEED_BASE + 9, bs_exp=(8, 11), channels=6,
                    couplings=((0, 2), (3, 4), (1, 0)), unused_prob=0.05):
    """Multichannel Vorbis batch (symgpu_vorbis_mc_*): `channels` planes per packet and a list of coupling steps
    (magnitude channel, angle channel) applied in order -- the default is a 5.1 layout in which channel 0 takes part in two steps,
    so the order of the steps matters.  Same floors / residues / block mix as vorbis_batch.

    Returns dict(streams, floors, units [P], floor_y [P,C,65], residue [P,C,slot], runs, slot, out_len [P], channels)."""
    from ._native import VORBIS_STREAM_MC_DTYPE, VORBIS_UNIT_MC_DTYPE
    rng = np.random.Generator(np.random.PCG64(seed))
    S, F, C = int(n_streams), int(packets_per_stream), int(channels)
    bs0, bs1 = 1 << bs_exp[0], 1 << bs_exp[1]
    slot = bs1 // 2
    streams = np.zeros(S, dtype=VORBIS_STREAM_MC_DTYPE)
    streams["bs0_exp"], streams["bs1_exp"], streams["channels"], streams["n_couplings"] = bs_exp[0], bs_exp[1], C, len(couplings)
    for k, (m, a) in enumerate(couplings):
        streams["magnitude_ch"][:, k], streams["angle_ch"][:, k] = m, a
    floors, by_size = [], {0: [], 1: []}
    for flag, n2 in ((0, bs0 // 2), (1, bs1 // 2)):
        for _ in range(4):
            n_posts = int(min(rng.integers(20, 41), n2 // 2))
            inner = rng.choice(np.arange(1, n2), size=n_posts - 2, replace=False).tolist()
            rng.shuffle(inner)
            by_size[flag].append(len(floors))
            floors.append(make_floor1_setup([0, n2] + inner, int(rng.choice([1, 2, 2, 2, 3, 4]))))
    floors = np.array(floors, dtype=VORBIS_FLOOR1_DTYPE)
    P = S * F
    units = np.zeros((S, F), dtype=VORBIS_UNIT_MC_DTYPE)
    units["floor"] = 0xFFFF
    units["do_not_decode"] = 1
    floor_y = np.zeros((S, F, C, 65), dtype=np.uint16)
    residue = np.zeros((S, F, C, slot), dtype=np.float32)
    out_len = np.zeros((S, F), dtype=np.int32)
    flag = (rng.random(S) < 0.8).astype(np.uint8)
    for f in range(F):
        if f:
            u = rng.random(S)
            flag = np.where(flag == 1, (u < 0.9).astype(np.uint8), (u >= 0.7).astype(np.uint8))
        prev = units["block_flag"][:, f - 1] if f else flag
        units["block_flag"][:, f] = flag
        units["prev_block_flag"][:, f] = prev
        out_len[:, f] = (np.where(prev == 1, bs1, bs0) + np.where(flag == 1, bs1, bs0)) // 4
        for s in range(S):
            n2 = (bs1 if flag[s] else bs0) // 2
            dnd = []
            for ch in range(C):
                fi = by_size[int(flag[s])][int(rng.integers(0, 4))]
                setup = floors[fi]
                npost = int(setup["n_posts"])
                rng_ = {1: 256, 2: 128, 3: 86, 4: 64}[int(setup["multiplier"])]
                y = rng.integers(0, 24, size=npost)
                y[rng.random(npost) < 0.3] = 0
                y[0], y[1] = rng.integers(0, rng_, size=2)
                floor_y[s, f, ch, :npost] = y
                unused = rng.random() < unused_prob
                units[s, f]["floor"][ch] = 0xFFFF if unused else fi
                dnd.append(bool(unused))
                r = rng.standard_normal(n2).astype(np.float32) * 4.0
                r[int(0.8 * n2):] = 0.0
                residue[s, f, ch, :n2] = r
            for m, a in couplings:  # non-zero vector propagate (lib.rs:215-225), in mapping order
                if dnd[m] != dnd[a]:
                    dnd[m] = dnd[a] = False
            units[s, f]["do_not_decode"][:C] = [int(d) for d in dnd]
    runs = np.zeros(S, dtype=VORBIS_RUN_DTYPE)
    runs["stream"] = np.arange(S)
    runs["first_packet"] = np.arange(S) * F
    runs["n_packets"] = F
    return dict(streams=streams, floors=floors, units=units.reshape(P), floor_y=floor_y.reshape(P, C, 65),
                residue=residue.reshape(P, C, slot), runs=runs, slot=slot, out_len=out_len.reshape(P), channels=C)
