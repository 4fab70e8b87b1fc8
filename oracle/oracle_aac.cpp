// ORACLE (test infrastructure, NOT product code): CPU restatement of the AAC-LC synthesis stage of
// pdeljanov/Symphonia @ ee35874:
//   Tns::synth             symphonia-codec-aac/src/aac/ics/tns.rs:149-199 (filter loops :183-196)
//   generate_window        symphonia-codec-aac/src/aac/window.rs:28-63
//   Dsp::new / Dsp::synth  symphonia-codec-aac/src/aac/dsp.rs:34-158
//   Ics::synth_channel     symphonia-codec-aac/src/aac/ics/mod.rs:449-468 (pulse stays with the parser)
// on top of oracle_mdct.cpp (symphonia-core Imdct, no_simd FFT).
//
// PARITY PINNING: the reference holds no known-answer vector for the AAC filterbank; its only
// pins are the shared IMDCT/FFT vectors (replayed in tests/test_oracle_kat_mdct.py).  On top of
// those, tests/test_oracle_kat_aac.py checks this file against the textbook definition in f64
// (TDAC: windowed IMDCT + overlap-add reconstructs a windowed-MDCT'd signal) at 1e-5.
// Bit-level agreement with the Rust binary is by construction.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

struct AacTables {
    float sine_long[1024], sine_short[128], kbd_long[1024], kbd_short[128];
    AacTables() {
        sine(sine_long, 1024);
        sine(sine_short, 128);
        kbd(kbd_long, 1024, 4.0f);
        kbd(kbd_short, 128, 6.0f);
    }
    // window.rs:29-36 (half = true, scale = 1.0).  All f32 arithmetic, f32 sinf.
    static void sine(float* dst, int size) {
        const float PI = 3.14159265358979323846264338327950288f;
        const float param = PI / (float)(2 * size);
        for (int n = 0; n < size; ++n) dst[n] = std::sin(((float)n + 0.5f) * param) * 1.0f;
    }
    static double bessel_i0(double inval) { // window.rs:56-63
        double val = 1.0;
        for (int n = 63; n >= 1; --n) {
            val *= inval / (double)(n * n);
            val += 1.0;
        }
        return val;
    }
    static void kbd(float* dst, int size, float alpha) { // window.rs:37-52 (half = true)
        const float PI = 3.14159265358979323846264338327950288f;
        const float dlen = (float)size;
        const double alpha2 = (double)((alpha * PI / dlen) * (alpha * PI / dlen));
        std::vector<double> kb(size);
        double sum = 0.0;
        for (int n = 0; n < size; ++n) {
            sum += bessel_i0((double)(n * (size - n)) * alpha2);
            kb[n] = sum;
        }
        sum += 1.0;
        for (int n = 0; n < size; ++n) dst[n] = (float)std::sqrt(kb[n] / sum);
    }
};
const AacTables& W() {
    static const AacTables t;
    return t;
}

} // namespace

extern "C" {

const float* oracle_aac_window(int kbd, int is_short) {
    return kbd ? (is_short ? W().kbd_short : W().kbd_long) : (is_short ? W().sine_short : W().sine_long);
}

// The filter loops of Tns::synth (tns.rs:181-196) over already-resolved line ranges.
void oracle_aac_tns(float* coeffs, const symgpu_aac_tns* filters, uint32_t n_filters) {
    for (uint32_t f = 0; f < n_filters; ++f) {
        const symgpu_aac_tns& t = filters[f];
        const int order = t.order;
        if (order == 0) continue;
        if (!t.direction) {
            int m = 0;
            for (int i = t.start; i < (int)t.end; ++i, ++m)
                for (int j = 0; j < std::min(order, m); ++j) coeffs[i] -= coeffs[i - j - 1] * t.lpc[j];
        } else {
            int m = 0;
            for (int i = (int)t.end - 1; i >= (int)t.start; --i, ++m)
                for (int j = 0; j < std::min(order, m); ++j) coeffs[i] -= coeffs[i + j + 1] * t.lpc[j];
        }
    }
}

// Dsp::synth (dsp.rs:57-158).
void oracle_aac_synth(const float* coeffs, float* delay, int seq, int window_shape, int prev_window_shape, float* dst) {
    const float* long_win = window_shape ? W().kbd_long : W().sine_long;
    const float* short_win = window_shape ? W().kbd_short : W().sine_short;
    const float* prev_long_win = prev_window_shape ? W().kbd_long : W().sine_long;
    const float* prev_short_win = prev_window_shape ? W().kbd_short : W().sine_short;
    const int P0 = 512 - 64, P1 = 512 + 64;
    float pcm_long[2048];
    float pcm_short[1152];
    if (seq != SYMGPU_AAC_EIGHT_SHORT) {
        oracle::imdct_for(1024, 1.0 / 2048.0).run(coeffs, pcm_long);
    } else {
        const oracle::Imdct& im = oracle::imdct_for(128, 1.0 / 256.0);
        for (int w = 0; w < 8; ++w) im.run(coeffs + 128 * w, pcm_long + 256 * w);
        for (int i = 0; i < 1152; ++i) pcm_short[i] = 0.0f;
        for (int w = 0; w < 8; ++w) {
            const float* src = pcm_long + 256 * w;
            if (w > 0) {
                for (int i = 0; i < 128; ++i) {
                    pcm_short[w * 128 + i] += src[i] * short_win[i];
                    pcm_short[w * 128 + i + 128] += src[i + 128] * short_win[127 - i];
                }
            } else {
                for (int i = 0; i < 128; ++i) {
                    pcm_short[i] = src[i] * prev_short_win[i];
                    pcm_short[i + 128] = src[i + 128] * short_win[127 - i];
                }
            }
        }
    }
    switch (seq) { // output, dsp.rs:104-129
        case SYMGPU_AAC_ONLY_LONG:
        case SYMGPU_AAC_LONG_START:
            for (int i = 0; i < 1024; ++i) dst[i] = delay[i] + (pcm_long[i] * prev_long_win[i]);
            break;
        case SYMGPU_AAC_EIGHT_SHORT:
            std::memcpy(dst, delay, sizeof(float) * P0);
            for (int i = P0; i < 1024; ++i) dst[i] = delay[i] + pcm_short[i - P0];
            break;
        default: // LONG_STOP
            std::memcpy(dst, delay, sizeof(float) * P0);
            for (int i = P0; i < P1; ++i) dst[i] = delay[i] + pcm_long[i] * prev_short_win[i - P0];
            for (int i = P1; i < 1024; ++i) dst[i] = delay[i] + pcm_long[i];
            break;
    }
    switch (seq) { // new delay, dsp.rs:131-157
        case SYMGPU_AAC_ONLY_LONG:
        case SYMGPU_AAC_LONG_STOP:
            for (int i = 0; i < 1024; ++i) delay[i] = pcm_long[i + 1024] * long_win[1023 - i];
            break;
        case SYMGPU_AAC_EIGHT_SHORT:
            for (int i = 0; i < P1; ++i) delay[i] = pcm_short[i + 512 + 64];
            for (int i = P1; i < 1024; ++i) delay[i] = 0.0f;
            break;
        default: // LONG_START
            std::memcpy(delay, pcm_long + 1024, sizeof(float) * P0);
            for (int i = P0; i < P1; ++i) delay[i] = pcm_long[i + 1024] * short_win[127 - (i - P0)];
            for (int i = P1; i < 1024; ++i) delay[i] = 0.0f;
            break;
    }
}

// Same contract as symgpu_aac_synth_host.  states[stream] holds the delay of channel 0 then 1.
int oracle_aac_batch(oracle_aac_state* states, const symgpu_aac_unit* units, const symgpu_aac_tns* tns,
                     const float* coeffs, const symgpu_aac_run* runs, uint32_t n_runs, float* pcm, int n_threads) {
    auto work = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; ++r) {
            const int n_ch = runs[r].channels ? runs[r].channels : 2;
            for (uint32_t f = runs[r].first_frame; f < runs[r].first_frame + runs[r].n_frames; ++f)
                for (int ch = 0; ch < n_ch; ++ch) {
                    const symgpu_aac_unit& u = units[2 * (size_t)f + ch];
                    float c[1024];
                    std::memcpy(c, coeffs + (2 * (size_t)f + ch) * 1024, sizeof c);
                    if (u.n_tns) oracle_aac_tns(c, tns + u.tns_first, u.n_tns);
                    oracle_aac_synth(c, states[2 * (size_t)runs[r].stream + ch].delay, u.window_sequence, u.window_shape,
                                     u.prev_window_shape, pcm + (2 * (size_t)f + ch) * 1024);
                }
        }
    };
    if (n_threads <= 1) {
        work(0, n_runs);
        return 0;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back(work, (uint32_t)((uint64_t)n_runs * t / n_threads), (uint32_t)((uint64_t)n_runs * (t + 1) / n_threads));
    for (auto& th : pool) th.join();
    return 0;
}
}
