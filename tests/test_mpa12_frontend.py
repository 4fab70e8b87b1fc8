"""Layer I / II sample decoders (`symgpu_mpa12_fe_*`, SURVEY §8f N1 for the Layer I / II path) against
oracle/mpa12_frontend_oracle.py (the reference's sequence, numpy f32 arithmetic, the reference's literal constants) and
against an independent bitstream writer.  Floating point, but every sample is a short chain of single IEEE operations
on exact inputs, so the bar is bit equality.  CPU only."""
import json
import os

import numpy as np
import pytest

import symphonia_b200 as sb
from oracle import mpa12_frontend_oracle as mo
from oracle.mp3_frontend_oracle import DecodeError
from symphonia_b200 import frontend, packetizer
from tests import _mpa12_bitstream as bw
from tests import _streams as st

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpa12_constants.json")))


def test_closed_form_constants_equal_the_reference_literals():
    scale, c, d = frontend.mpa12_constants()
    assert scale.view(np.uint32).tolist() == GOLD["scalefactors"]
    assert c.view(np.uint32).tolist() == [q["c"] for q in GOLD["quant_classes"]]
    assert d.view(np.uint32).tolist() == [q["d"] for q in GOLD["quant_classes"]]
    # the two D values the standard's 11 printed decimals push off the power of two
    assert float(d[15]) != 2.0 ** -14 and float(d[16]) != 2.0 ** -15 and float(d[14]) == 2.0 ** -13
    assert float(scale[3]) == 1.0 and float(scale[0]) == 2.0 and float(scale[63]) == 0.0


def test_hand_computed_samples():
    # Layer I, mono, 32 kbit/s at 32 kHz (48-byte frame): sub-band 0 gets 4-bit samples (allocation field 3), scale factor
    # index 3 (= 1.0).  Codes 1111, 1000, 0000 -> offset binary -> 7, 0, -8 -> (x + 1) * (16 / 15) / 8
    w = bw.BitWriterMsb()
    w.put(3, 4)
    for _ in range(31):
        w.put(0, 4)
    w.put(3, 6)
    for v in (0b1111, 0b1000, 0b0000) + (0b1000,) * 9:
        w.put(v, 4)
    frame = st.mpa_word(version="1", layer=1, bitrate_idx=1, rate_idx=2, mode=3).to_bytes(4, "big") + w.bytes()
    frame += bytes(48 - len(frame))
    out, info = frontend.mpa12_decode(frame, 1)
    f = np.float32(16) / np.float32(15) * (np.float32(1) / np.float32(8))
    assert out[0, 0, :3].tolist() == [float(f * np.float32(8)), float(f * np.float32(1)), float(f * np.float32(-7))]
    assert abs(out[0, 0, 0] - 16 / 15) < 1e-6 and not out[1].any() and not out[0, 1:].any()
    h, want = mo.Mpa12Frontend(1).decode(frame)
    assert (out.reshape(2, 384).view(np.uint32) == want.view(np.uint32)).all()
    # Layer II, mono 64 kbit/s at 48 kHz: 64 kbit/s per channel -> Table 3-B.2a; sub-band 0 allocation index 1 = 3 levels,
    # grouped: codeword 0 + 3*1 + 9*2 = 21 -> levels (0, 1, 2) -> the 3-level quantiser's -2/3, 0, +2/3; scfsi 2, scale factor 1.0
    w = bw.BitWriterMsb()
    w.put(1, 4)
    for sb in range(1, 27):
        w.put(0, 4 if sb < 11 else 3 if sb < 23 else 2)
    w.put(2, 2), w.put(3, 6)
    for _ in range(12):
        w.put(21, 5)
    frame = st.mpa_word(version="1", layer=2, bitrate_idx=4, rate_idx=1, mode=3).to_bytes(4, "big") + w.bytes()
    frame += bytes(192 - len(frame))
    out, info = frontend.mpa12_decode(frame, 2)
    third = np.float32(4) / np.float32(3)
    assert out[0, 0, :3].tolist() == [float(third * np.float32(-0.5)), 0.0, float(third * np.float32(0.5))]
    assert (out[0, 0].reshape(12, 3) == out[0, 0, :3]).all() and not out[0, 1:].any()
    h, want = mo.Mpa12Frontend(2).decode(frame)
    assert (out.reshape(2, 1152).view(np.uint32) == want.view(np.uint32)).all()


def _both(frames, layer):
    ofe = mo.Mpa12Frontend(layer)
    outcomes = []
    for k, f in enumerate(frames):
        try:
            h, want = ofe.decode(f)
        except DecodeError:
            want = None
        try:
            got, info = frontend.mpa12_decode(f, layer)
            assert want is not None, f"frame {k}: accepted, the oracle refuses it"
            n_slots = 12 if layer == 1 else 36
            same = got.reshape(2, 32 * n_slots).view(np.uint32) == want.view(np.uint32)
            assert same.all(), (k, np.argwhere(~same)[:3])
            assert (int(info["channels"]), int(info["sample_rate"])) == (h["n_channels"], h["sample_rate"])
        except sb.SymgpuError:
            assert want is None, f"frame {k}: refused, the oracle accepts it"
        outcomes.append(want is not None)
    return outcomes


@pytest.mark.parametrize("version,bitrate_idx,rate_idx,mode,protected", [
    ("1", 9, 0, 0, False), ("1", 14, 1, 1, True), ("1", 2, 2, 3, False), ("2", 5, 0, 1, False), ("2.5", 3, 2, 3, True), ("1", 7, 0, 2, False)])
def test_layer1_streams(version, bitrate_idx, rate_idx, mode, protected):
    rng = np.random.default_rng(10 * bitrate_idx + mode)
    frames, truths = [], []
    for k in range(25):
        f, t = bw.gen_layer1_frame(rng, version, bitrate_idx, rate_idx, mode, mode_ext=k % 4, protected=protected)
        frames.append(f), truths.append(t)
    assert all(_both(frames, 1))
    # what the writer packed is what comes out: x = (code - 2^(nb-1) + 1) * 2^nb / (2^nb - 1) / 2^(nb-1) * scale
    scale = 2.0 ** (1 - np.arange(64) / 3.0)
    scale[63] = 0
    for f, t in zip(frames, truths):
        got, _ = frontend.mpa12_decode(f, 1)
        for ch in range(t["n_ch"]):
            for sbn in range(32):
                src = ch if sbn < t["bound"] else 0
                nb = t["alloc"][src][sbn] + 1 if t["alloc"][src][sbn] else 0
                if not nb:
                    assert not got[ch, sbn].any()
                    continue
                ideal = (t["raw"][src, sbn] - 2.0 ** (nb - 1) + 1) * (2.0 ** nb / (2.0 ** nb - 1)) / 2.0 ** (nb - 1) * scale[t["sf"][ch][sbn]]
                assert np.allclose(got[ch, sbn], ideal, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("version,bitrate_idx,rate_idx,mode,protected", [
    ("1", 8, 0, 0, False),    # 64 kbit/s per channel: table a
    ("1", 14, 0, 1, True),    # 192 per channel at 44.1 kHz: table b, joint stereo
    ("1", 12, 1, 0, False),   # 48 kHz: table a above 80 kbit/s as well
    ("1", 9, 0, 0, False),    # exactly 80 kbit/s per channel at 44.1 kHz: still table a (b starts above 80)
    ("1", 2, 0, 3, False),    # 48 kbit/s mono: table c
    ("1", 1, 2, 3, False),    # 32 kbit/s mono at 32 kHz: table d
    ("1", 6, 2, 1, False),    # 96 kbit/s stereo = 48 per channel at 32 kHz: table d, joint stereo bound above sblimit
    ("2", 10, 0, 1, False), ("2.5", 4, 1, 3, True), ("2", 14, 2, 2, False)])
def test_layer2_streams(version, bitrate_idx, rate_idx, mode, protected):
    rng = np.random.default_rng(100 + 10 * bitrate_idx + mode)
    frames, truths = [], []
    for k in range(25):
        f, t = bw.gen_layer2_frame(rng, version, bitrate_idx, rate_idx, mode, mode_ext=k % 4, protected=protected)
        frames.append(f), truths.append(t)
    assert all(_both(frames, 2))
    scale = 2.0 ** (1 - np.arange(64) / 3.0)
    scale[63] = 0
    for f, t in zip(frames, truths):
        got, _ = frontend.mpa12_decode(f, 2)
        for ch in range(t["n_ch"]):
            for sbn in range(32):
                src = ch if sbn < t["bound"] else 0
                a = t["alloc"][src][sbn] if sbn < t["sblimit"] else 0
                if not a:
                    assert not got[ch, sbn].any()
                    continue
                levels = t["table"][sbn][1][a]
                b = int(np.ceil(np.log2(levels + 1)))
                dd = 0.5 if levels in (3, 5, 9) else 2.0 ** -(b - 1)
                s = (t["raw"][src, sbn] - 2.0 ** (b - 1)) / 2.0 ** (b - 1)
                ideal = (2.0 ** b / levels) * (s + dd) * scale[t["sf"][ch, np.arange(36) // 12, sbn]]
                assert np.allclose(got[ch, sbn], ideal, rtol=2e-6, atol=1e-10), (ch, sbn, levels)


def test_refusals_and_streams():
    rng = np.random.default_rng(5)
    f1, _ = bw.gen_layer1_frame(rng, "1", 9, 0, 0)
    f2, _ = bw.gen_layer2_frame(rng, "1", 8, 0, 0)
    bad_alloc = bytearray(f1)
    bad_alloc[4] |= 0xF0   # first allocation field = 15
    seq1 = [f1, bytes(bad_alloc), f1[:-1], f2, f1, b"", f1[:40]]
    assert _both(seq1, 1) == [True, False, False, False, True, False, False]
    assert _both([f2, f1, f2[:100], f2], 2) == [True, False, False, True]
    # streams through the packetiser: junk in front, a Layer I packet inside a Layer II stream, another sample rate
    frames = [bw.gen_layer2_frame(rng, "1", 8, 0, 0)[0] for _ in range(12)]
    alien = bw.gen_layer2_frame(rng, "1", 8, 1, 0)[0]
    noise = rng.integers(0, 255, 100, dtype=np.uint8).tobytes()
    data = noise + b"".join(frames[:5]) + alien + b"".join(frames[5:])
    track, packets = packetizer.mpa_index(data)
    assert len(packets) == 13 and int(track["layer"]) == 2
    sub, frame_of, info = frontend.mpa12_decode_packets(data, packets, 2)
    assert frame_of.tolist() == [k for k in range(13) if k != 5] and sub.shape == (12, 2, 32, 36)
    for row, f in zip(sub, frames):
        assert (row == frontend.mpa12_decode(f, 2)[0]).all()
