// ORACLE (test infrastructure, NOT product code): CPU restatement of symphonia-core's power-of-two
// IMDCT and its in-tree radix-2 FFT (the `no_simd` build: symphonia-core default features).
//
//   Imdct::new_scaled / Imdct::imdct   symphonia-core/src/dsp/mdct.rs:35-146
//   Fft::new / fft_inplace / transform symphonia-core/src/dsp/fft/no_simd.rs:74-118, :221-281
//   fft32/16/8/4/2                      symphonia-core/src/dsp/fft/no_simd.rs:289-454
//   Complex<f32> Mul (num-complex 0.4, not vendored): (a+bi)(c+di) = (ac - bd) + (ad + bc)i, each
//   product and sum rounded separately.
//
// PARITY PINNING: tests/test_oracle_kat_mdct.py replays the reference's own vectors -- the 32-point
// ramp of mdct.rs:177-201 and the 64-point complex TEST_VECTOR of fft/mod.rs:88-153 against naive
// f64 DFT/IMDCT at the reference's 1e-5.  For the `opt-simd` build (rustfft) parity is UNPINNED:
// that dependency is not vendored and picks algorithms at run time.
//
// Structure is our own: the reference hard-codes fft4..fft32 and then merges breadth-first; here
// every level is one generic decimation-in-time pass with a per-level butterfly rule, which performs
// exactly the same floating-point operations on the same operands.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "oracle.h"

namespace oracle {

struct Cpx {
    float re, im;
};
static inline Cpx add(Cpx a, Cpx b) { return {a.re + b.re, a.im + b.im}; }
static inline Cpx sub(Cpx a, Cpx b) { return {a.re - b.re, a.im - b.im}; }
static inline Cpx mul(Cpx a, Cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

static const float kC = 0.707106781186547524400844362104849039f; // f32::consts::FRAC_1_SQRT_2
static const double kPi = 3.14159265358979323846264338327950288;

// Twiddle of level `size` (butterflies span `size` points, half = size/2), index k in [0, half):
// (cos(pi k / half), -sin(pi k / half)) computed in f64 and cast (no_simd.rs:16-36 for size >= 64;
// for 16 and 32 the reference's 20-digit literals, which are the same values:
// tools/verify_constants_vs_reference.py checks them bit for bit).
static Cpx twiddle(int size, int k) {
    const double ang = kPi / (double)(size / 2) * (double)k;
    return {(float)std::cos(ang), (float)(-std::sin(ang))};
}

// One butterfly of level `size`: p' = p + q, o' = p - q, q = o "times" W_size^k with the reference's
// special cases for size <= 32.
static inline void butterfly(int size, int k, Cpx& p, Cpx& o) {
    Cpx q;
    const int half = size / 2;
    if (size <= 32 && k == 0) {
        q = o; // x1[0] is used as is (no_simd.rs:308, :375, :418, :442)
    } else if (size <= 32 && size >= 4 && 2 * k == half) {
        q = {o.im, -o.re}; // k = N/4 (:316, :379, :420, :440)
    } else if (size <= 32 && size >= 8 && 4 * k == half) {
        const float a = kC * o.re, b = kC * o.im; // k = N/8 (:302-303, :312)
        q = {a + b, b - a};
    } else if (size <= 32 && size >= 8 && 4 * k == 3 * half) {
        const float a = -kC * o.re, b = -kC * o.im; // k = 3N/8 (:304-305, :320)
        q = {a - b, a + b};
    } else if (size <= 32) {
        q = mul(twiddle(size, k), o); // literal on the left (:309 ...)
    } else {
        q = mul(o, twiddle(size, k)); // merge(): o[0] * w[0] (:227)
    }
    const Cpx pp = p;
    p = add(pp, q);
    o = sub(pp, q);
}

static void fft_inplace(Cpx* x, int n) {
    // Bit reversal (no_simd.rs:83-85, :101-107).
    int bits = 0;
    while ((1 << bits) < n) ++bits;
    for (int i = 0; i < n; ++i) {
        int j = 0;
        for (int b = 0; b < bits; ++b)
            if (i & (1 << b)) j |= 1 << (bits - 1 - b);
        if (i < j) std::swap(x[i], x[j]);
    }
    // Levels 2, 4, ..., n.  Twiddles for size >= 64 are cached per size.
    static std::mutex mu;
    static std::map<int, std::vector<Cpx>> cache;
    for (int size = 2; size <= n; size <<= 1) {
        const int half = size >> 1;
        const Cpx* tw = nullptr;
        if (size > 32) {
            std::lock_guard<std::mutex> lock(mu);
            auto& v = cache[size];
            if (v.empty())
                for (int k = 0; k < half; ++k) v.push_back(twiddle(size, k));
            tw = v.data();
        }
        for (int base = 0; base < n; base += size)
            for (int k = 0; k < half; ++k) {
                if (tw) {
                    const Cpx q = mul(x[base + half + k], tw[k]);
                    const Cpx p = x[base + k];
                    x[base + k] = add(p, q);
                    x[base + half + k] = sub(p, q);
                } else {
                    butterfly(size, k, x[base + k], x[base + half + k]);
                }
            }
    }
}

Imdct::Imdct(int n_, double scale) : n(n_) {
    // mdct.rs:35-60
    const int n2 = n / 2;
    tw_re.resize(n2);
    tw_im.resize(n2);
    const double alpha = 1.0 / 8.0 + (std::signbit(scale) ? (double)n2 : 0.0);
    const double pi_n = kPi / (double)n;
    const double sqrt_scale = std::sqrt(std::fabs(scale));
    for (int k = 0; k < n2; ++k) {
        const double theta = pi_n * (alpha + (double)k);
        tw_re[k] = (float)(sqrt_scale * std::cos(theta));
        tw_im[k] = (float)(sqrt_scale * std::sin(theta));
    }
}

void Imdct::run(const float* spec, float* out) const {
    // mdct.rs:67-146
    const int n2 = n >> 1, n4 = n >> 2;
    std::vector<Cpx> z(n2);
    for (int i = 0; i < n2; ++i) {
        const float even = spec[i * 2];
        const float odd = -spec[n - 1 - i * 2];
        z[i].re = odd * tw_im[i] - even * tw_re[i];
        z[i].im = odd * tw_re[i] + even * tw_im[i];
    }
    fft_inplace(z.data(), n2);
    float* vec0 = out;
    float* vec1 = out + n2;
    float* vec2 = out + 2 * n2;
    float* vec3 = out + 3 * n2;
    for (int i = 0; i < n4; ++i) {
        const Cpx w = {tw_re[i], tw_im[i]};
        const Cpx val = mul(w, Cpx{z[i].re, -z[i].im});
        const int fi = 2 * i, ri = n2 - 1 - 2 * i;
        vec0[ri] = -val.im;
        vec1[fi] = val.im;
        vec2[ri] = val.re;
        vec3[fi] = val.re;
    }
    for (int i = 0; i < n2 - n4; ++i) {
        const Cpx w = {tw_re[n4 + i], tw_im[n4 + i]};
        const Cpx val = mul(w, Cpx{z[n4 + i].re, -z[n4 + i].im});
        const int fi = 2 * i, ri = n2 - 1 - 2 * i;
        vec0[fi] = -val.re;
        vec1[ri] = val.re;
        vec2[fi] = val.im;
        vec3[ri] = val.im;
    }
}

const Imdct& imdct_for(int n, double scale) {
    static std::mutex mu;
    static std::map<std::pair<int, double>, std::unique_ptr<Imdct>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto& p = cache[{n, scale}];
    if (!p) p.reset(new Imdct(n, scale));
    return *p;
}

} // namespace oracle

extern "C" {

// Forward complex FFT of n interleaved (re, im) points, in place (Fft::fft_inplace).
void oracle_fft_inplace(float* x, int n) { oracle::fft_inplace(reinterpret_cast<oracle::Cpx*>(x), n); }

// N-point IMDCT with scaling: spec[n] -> out[2n] (Imdct::new_scaled(n, scale).imdct).
void oracle_imdct(const float* spec, float* out, int n, double scale) { oracle::imdct_for(n, scale).run(spec, out); }

// Imdct.twiddle[k] of Imdct::new_scaled(n, scale) (mdct.rs:42-54).
void oracle_imdct_twiddle(int n, double scale, int k, float* re_im) {
    const oracle::Imdct& im = oracle::imdct_for(n, scale);
    re_im[0] = im.tw_re[k];
    re_im[1] = im.tw_im[k];
}

// Twiddle of the level-`size` butterfly, for the constant check against the reference's literals.
void oracle_fft_twiddle(int size, int k, float* re_im) {
    const oracle::Cpx w = oracle::twiddle(size, k);
    re_im[0] = w.re;
    re_im[1] = w.im;
}
}
