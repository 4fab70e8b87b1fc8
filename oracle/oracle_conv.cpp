// ORACLE (test infrastructure, NOT product code): CPU restatement of the reference's output stage --
// gapless trim, interleave and f32 -> integer sample conversion.
//
//   AudioBuffer::truncate / shift / trim       symphonia-core/src/audio/buf.rs:404-433
//   copy_to_slice_interleaved                  symphonia-core/src/audio/buf.rs:469-476
//   FromSample<f32> for u8, i16, i24, i32, f32 symphonia-core/src/audio/conv.rs:592-607
//   clamp_f32, clamp_i24                       symphonia-core/src/util.rs:230-237, :258-266
//   Rust `as` from float to int: truncates toward zero, saturates, NaN -> 0 (language semantics).
//
// PARITY PINNING: tests/test_oracle_kat_conv.py replays the reference's own assertions
// (conv.rs:713-715, :885-887, :928-930, :971-973, :1014-1016, :1100-1102: +1.0 -> MAX, 0 -> MID,
// -1.0 -> MIN for each target type).
#include <cmath>
#include <cstdint>
#include <cstring>

#include "oracle.h"

namespace {

inline float clamp_f32(float v) {
    float c = v;
    c = c > 1.0f ? 1.0f : c;
    c = c < -1.0f ? -1.0f : c;
    return c;
}

// Rust `x as iN` / `x as uN` for a float x.
template <typename F>
inline int64_t cast_sat(F x, int64_t lo, int64_t hi) {
    if (x != x) return 0;
    if (x <= (F)lo) return lo;
    if (x >= (F)hi) return hi;
    return (int64_t)x; // C++ truncates toward zero; in range here
}

inline int32_t clamp_i24(int32_t v) {
    if (((uint32_t)v + 0x00800000u) & ~0x00ffffffu) return 0x007fffff ^ (v >> 31);
    return v;
}

} // namespace

extern "C" {

int16_t oracle_conv_s16(float s) { return (int16_t)cast_sat<float>(clamp_f32(s) * 32768.0f, -32768, 32767); }
int32_t oracle_conv_s24(float s) {
    return clamp_i24((int32_t)cast_sat<float>(clamp_f32(s) * 8388608.0f, INT32_MIN, INT32_MAX));
}
int32_t oracle_conv_s32(float s) {
    return (int32_t)cast_sat<double>((double)clamp_f32(s) * 2147483648.0, INT32_MIN, INT32_MAX);
}
uint8_t oracle_conv_u8(float s) { return (uint8_t)cast_sat<float>((clamp_f32(s) + 1.0f) * 128.0f, 0, 255); }

// Same contract as symgpu_pcm_pack_host (include/symgpu.h).
int oracle_pcm_pack(const float* pcm, const symgpu_pcm_span* spans, uint32_t n_spans, uint32_t channels,
                    uint32_t plane_stride, uint32_t frames, int format, void* out) {
    for (uint32_t p = 0; p < n_spans; ++p) {
        symgpu_pcm_span sp;
        if (spans) {
            sp = spans[p];
        } else {
            sp.src = (uint64_t)p * channels * plane_stride;
            sp.plane_stride = plane_stride;
            sp.frames = frames;
            sp.trim_start = sp.trim_end = 0;
            sp.dst_frame = (uint64_t)p * frames;
        }
        // trim(): first truncate the end, then shift the start (buf.rs:426-433).
        uint32_t n = sp.frames > sp.trim_end ? sp.frames - sp.trim_end : 0;
        uint32_t first = 0;
        if (sp.trim_start >= n) {
            n = 0;
        } else {
            first = sp.trim_start;
            n -= sp.trim_start;
        }
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t c = 0; c < channels; ++c) {
                const float s = pcm[sp.src + (uint64_t)c * sp.plane_stride + first + i];
                const uint64_t o = (sp.dst_frame + i) * channels + c;
                switch (format) {
                case SYMGPU_FMT_F32: static_cast<float*>(out)[o] = s; break;
                case SYMGPU_FMT_S16: static_cast<int16_t*>(out)[o] = oracle_conv_s16(s); break;
                case SYMGPU_FMT_S24: static_cast<int32_t*>(out)[o] = oracle_conv_s24(s); break;
                case SYMGPU_FMT_S32: static_cast<int32_t*>(out)[o] = oracle_conv_s32(s); break;
                case SYMGPU_FMT_U8: static_cast<uint8_t*>(out)[o] = oracle_conv_u8(s); break;
                default: return 1;
                }
            }
    }
    return 0;
}
}
