"""Pins the output-stage oracle (oracle/oracle_conv.cpp) to the reference's own assertions and to the
language rules its conversions rely on.

Reference vectors: symphonia-core/src/audio/conv.rs:713-715 (u8), :885-887 (i8, same rule), :928-930 (i16),
:971-973 (i24), :1014-1016 (i32), :1100-1102 (f32): from_sample(+1.0) == MAX, (0.0) == MID, (-1.0) == MIN.
"""
import numpy as np

from symphonia_b200._native import FMT_F32, FMT_S16, FMT_S24, FMT_S32, FMT_U8, PCM_SPAN_DTYPE
from tests import _oracle


def test_reference_min_mid_max_vectors(oracle):
    assert (oracle.oracle_conv_s16(1.0), oracle.oracle_conv_s16(0.0), oracle.oracle_conv_s16(-1.0)) == (32767, 0, -32768)
    assert (oracle.oracle_conv_s24(1.0), oracle.oracle_conv_s24(0.0), oracle.oracle_conv_s24(-1.0)) == (8388607, 0, -8388608)
    assert (oracle.oracle_conv_s32(1.0), oracle.oracle_conv_s32(0.0), oracle.oracle_conv_s32(-1.0)) == (2147483647, 0, -2147483648)
    assert (oracle.oracle_conv_u8(1.0), oracle.oracle_conv_u8(0.0), oracle.oracle_conv_u8(-1.0)) == (255, 128, 0)


def test_clamp_truncation_and_nan(oracle):
    # clamp_f32 (util.rs:258-266) then a truncating, saturating cast; NaN survives the clamp and casts to 0.
    assert oracle.oracle_conv_s16(7.5) == 32767 and oracle.oracle_conv_s16(-3.0) == -32768
    assert oracle.oracle_conv_s16(float("inf")) == 32767 and oracle.oracle_conv_s16(float("-inf")) == -32768
    assert oracle.oracle_conv_s16(float("nan")) == 0 and oracle.oracle_conv_s32(float("nan")) == 0
    assert oracle.oracle_conv_u8(float("nan")) == 0  # (NaN + 1) * 128 is NaN -> 0, not MID
    # toward zero, both signs
    assert oracle.oracle_conv_s16(0.99999 / 32768.0) == 0 and oracle.oracle_conv_s16(-0.99999 / 32768.0) == 0
    assert oracle.oracle_conv_s16(1.5 / 32768.0) == 1 and oracle.oracle_conv_s16(-1.5 / 32768.0) == -1
    assert oracle.oracle_conv_s24(np.float32(0.5)) == 4194304 and oracle.oracle_conv_s32(np.float32(-0.5)) == -1073741824
    # the largest f32 below 1.0 keeps all 24 bits in every integer format
    below_one = float(np.nextafter(np.float32(1.0), np.float32(0.0)))
    assert oracle.oracle_conv_s24(below_one) == 8388607 and oracle.oracle_conv_s32(below_one) == 2147483520


def test_numpy_cross_check(oracle):
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.normal(0, 0.6, 4000), rng.uniform(-1, 1, 4000) * 2.0 ** rng.integers(-30, 1, 4000)])
    x = x.astype(np.float32)
    c = np.clip(x, -1.0, 1.0)
    want16 = np.clip(np.trunc(c.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int64)
    want32 = np.clip(np.trunc(c.astype(np.float64) * 2147483648.0), -2 ** 31, 2 ** 31 - 1).astype(np.int64)
    got16 = np.array([oracle.oracle_conv_s16(float(v)) for v in x])
    got32 = np.array([oracle.oracle_conv_s32(float(v)) for v in x])
    np.testing.assert_array_equal(got16, want16)  # f32 * 2^15 is exact, so f64 arithmetic agrees
    np.testing.assert_array_equal(got32, want32)


def test_trim_and_interleave_follow_audio_buffer_trim(oracle):
    # trim(start, end): truncate(frames - end) first, then shift(start) (buf.rs:426-433).
    pcm = np.arange(2 * 3 * 10, dtype=np.float32).reshape(2, 3, 10) / 64.0  # 2 packets, 3 planes, 10 frames
    spans = np.zeros(2, dtype=PCM_SPAN_DTYPE)
    spans[0] = (0, 10, 10, 3, 2, 0)       # keeps frames 3..7  -> 5 frames
    spans[1] = (30, 10, 10, 9, 4, 5)      # end trim leaves 6, start 9 >= 6 -> nothing
    out = _oracle.pcm_pack(oracle, pcm, spans, 3, FMT_F32, 6)
    np.testing.assert_array_equal(out[:5], pcm[0, :, 3:8].T)
    assert (out[5] == 0).all()
    spans[1] = (30, 10, 10, 1, 20, 5)     # end trim saturates to 0 frames
    out2 = _oracle.pcm_pack(oracle, pcm, spans, 3, FMT_F32, 6)
    np.testing.assert_array_equal(out, out2)
    # uniform packets: back to back, untrimmed
    out3 = _oracle.pcm_pack(oracle, pcm, None, 3, FMT_S16, 20, plane_stride=10, frames=10, n_spans=2)
    want = np.concatenate([pcm[0].T, pcm[1].T])
    np.testing.assert_array_equal(out3, np.trunc(np.clip(want, -1, 1) * 32768.0).clip(-32768, 32767).astype(np.int16))
    for fmt in (FMT_S24, FMT_S32, FMT_U8):
        assert _oracle.pcm_pack(oracle, pcm, spans, 3, fmt, 6).shape == (6, 3)
