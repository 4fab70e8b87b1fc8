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


def mp3_audio_seconds(n_frames, sample_rate_idx=0):
    rate = [44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000][sample_rate_idx]
    per_frame = 1152 if sample_rate_idx < 3 else 576
    return n_frames * per_frame / rate


MP3_ALGO_BYTES_PER_FRAME = 9216 + 256 + 9216  # spectra + descriptors + PCM
