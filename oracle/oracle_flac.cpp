// ORACLE (test infrastructure, NOT product code): CPU restatement of what FlacDecoder::decode_inner does after
// the Rice residuals are decoded.
//
//   fixed_predict                symphonia-bundle-flac/src/decoder.rs:663-707
//   lpc_predict::<N>             decoder.rs:713-752 (dispatched by order at :483-501)
//   samples_shl                  decoder.rs:387-394
//   decorrelate_left_side / mid_side / right_side   decoder.rs:32-82
//   final `sample << (32 - bps)` decoder.rs:232-241
//
// All of it is integer arithmetic; `+=`, `-` and `<<` on i32 are restated with wrapping semantics (a release
// build of the reference wraps; a debug build panics on streams that overflow, which valid streams do not).
//
// PARITY PINNING: the reference has no vectors for these functions (its only FLAC unit test covers the Rice
// mapping).  tests/test_oracle_kat_flac.py pins the restatement with the property the format exists for: an
// encoder written from the FLAC definition (residual = sample - prediction, exact integers) followed by this
// restoration returns the original samples bit for bit, for every sub-frame type, order, shift and assignment.
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t wshl(int32_t a, unsigned s) { return (int32_t)((uint32_t)a << (s & 31)); }

// lpc_predict with N coefficients stored most-recent-first (c[j] multiplies buf[i-1-j]); the reference zero-pads
// the coefficient array to N in {4, 6, 8, 10, 12, 32}, which changes nothing in integer arithmetic.
void lpc(int32_t* buf, uint32_t n, int order, const int32_t* c, unsigned shift) {
    for (uint32_t i = (uint32_t)order; i < n; ++i) {
        int64_t predicted = 0;
        for (int j = 0; j < order; ++j) predicted += (int64_t)c[j] * (int64_t)buf[i - 1 - j];
        buf[i] = wadd(buf[i], (int32_t)(predicted >> shift));
    }
}

void fixed(int32_t* buf, uint32_t n, int order) {
    static const int32_t kFixed[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
    // decoder.rs:672-705: i64 products summed with wrapping arithmetic, truncated to i32, added with wrapping_add
    for (uint32_t i = (uint32_t)order; i < n; ++i) {
        uint64_t acc = 0;
        for (int j = 0; j < order; ++j) acc += (uint64_t)((int64_t)kFixed[order][j] * (int64_t)buf[i - 1 - j]);
        buf[i] = wadd(buf[i], (int32_t)(int64_t)acc);
    }
}

} // namespace

extern "C" int oracle_flac_restore(const symgpu_flac_frame* frames, uint32_t n_frames, const symgpu_flac_subframe* subframes,
                                   uint32_t n_subframes, int32_t* samples, size_t n_samples) {
    for (uint32_t f = 0; f < n_frames; ++f) {
        const symgpu_flac_frame& fr = frames[f];
        if (fr.channels < 1 || fr.channels > 8 || (uint64_t)fr.first_subframe + fr.channels > n_subframes) return 1;
        if (fr.assignment != SYMGPU_FLAC_INDEPENDENT && fr.channels != 2) return 1;
        for (int c = 0; c < fr.channels; ++c) {
            const symgpu_flac_subframe& sf = subframes[fr.first_subframe + c];
            if (sf.offset + sf.n > n_samples || sf.n == 0) return 1;
            int32_t* buf = samples + sf.offset;
            switch (sf.type) {
            case SYMGPU_FLAC_CONSTANT:
                for (uint32_t i = 1; i < sf.n; ++i) buf[i] = buf[0]; // decode_constant, decoder.rs:396-404
                break;
            case SYMGPU_FLAC_VERBATIM: break;
            case SYMGPU_FLAC_FIXED:
                if (sf.order > 4 || sf.order > sf.n) return 1;
                fixed(buf, sf.n, sf.order);
                break;
            case SYMGPU_FLAC_LPC:
                if (sf.order < 1 || sf.order > 32 || sf.order > sf.n || sf.shift > 15) return 1;
                lpc(buf, sf.n, sf.order, sf.coeffs, sf.shift);
                break;
            default: return 1;
            }
            if (sf.wasted) // samples_shl, decoder.rs:387-394
                for (uint32_t i = 0; i < sf.n; ++i) buf[i] = wshl(buf[i], sf.wasted);
        }
        const uint32_t n = subframes[fr.first_subframe].n;
        int32_t* a = samples + subframes[fr.first_subframe].offset;
        int32_t* b = fr.channels >= 2 ? samples + subframes[fr.first_subframe + 1].offset : nullptr;
        if (fr.assignment != SYMGPU_FLAC_INDEPENDENT && subframes[fr.first_subframe + 1].n != n) return 1;
        switch (fr.assignment) {
        case SYMGPU_FLAC_LEFT_SIDE: // plane 0 = left, plane 1 = side -> right = left - side (decoder.rs:32-36)
            for (uint32_t i = 0; i < n; ++i) b[i] = wsub(a[i], b[i]);
            break;
        case SYMGPU_FLAC_MID_SIDE: // decoder.rs:38-76
            for (uint32_t i = 0; i < n; ++i) {
                const int32_t mid = wshl(a[i], 1) | (b[i] & 1);
                const int32_t side = b[i];
                a[i] = wadd(mid, side) >> 1;
                b[i] = wsub(mid, side) >> 1;
            }
            break;
        case SYMGPU_FLAC_RIGHT_SIDE: // plane 0 = side, plane 1 = right -> left = side + right (decoder.rs:78-82)
            for (uint32_t i = 0; i < n; ++i) a[i] = wadd(a[i], b[i]);
            break;
        default: break;
        }
        if (fr.bits_per_sample < 32) { // decoder.rs:237-240
            const unsigned sh = 32u - fr.bits_per_sample;
            for (int c = 0; c < fr.channels; ++c) {
                const symgpu_flac_subframe& sf = subframes[fr.first_subframe + c];
                for (uint32_t i = 0; i < sf.n; ++i) samples[sf.offset + i] = wshl(samples[sf.offset + i], sh);
            }
        }
    }
    return 0;
}
