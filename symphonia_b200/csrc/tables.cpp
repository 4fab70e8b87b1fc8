// Host-side table builder (product code; independent of oracle/).
//
// ISO constant data is stored here in "width" form (band widths, window numerators); the oracle
// keeps its own copy in "edge" form, and tests/test_tables.py requires the two to agree bit for
// bit, which guards against a typo in either.
#include "tables.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace symgpu {
namespace {

// ISO/IEC 11172-3 Table B.8 / 13818-3 Table B.2: widths of the 22 long scale-factor bands, in the
// sample-rate order 44.1k 48k 32k | 22.05k 24k 16k | 11.025k 12k 8k (layer3/common.rs:9-55).
const uint8_t kLongWidths[9][22] = {
    {4, 4, 4, 4, 4, 4, 6, 6, 8, 8, 10, 12, 16, 20, 24, 28, 34, 42, 50, 54, 76, 158},
    {4, 4, 4, 4, 4, 4, 6, 6, 6, 8, 10, 12, 16, 18, 22, 28, 34, 40, 46, 54, 54, 192},
    {4, 4, 4, 4, 4, 4, 6, 6, 8, 10, 12, 16, 20, 24, 30, 38, 46, 56, 68, 84, 102, 26},
    {6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 38, 46, 52, 60, 68, 58, 54},
    {6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 18, 22, 26, 32, 38, 46, 54, 62, 70, 76, 36},
    {6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 38, 46, 52, 60, 68, 58, 54},
    {6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 38, 46, 52, 60, 68, 58, 54},
    {6, 6, 6, 6, 6, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32, 38, 46, 52, 60, 68, 58, 54},
    {12, 12, 12, 12, 12, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 76, 90, 2, 2, 2, 2, 2},
};
// Width of ONE window of each of the 13 short bands (layer3/common.rs:60-107).
const uint8_t kShortWidths[9][13] = {
    {4, 4, 4, 4, 6, 8, 10, 12, 14, 18, 22, 30, 56}, {4, 4, 4, 4, 6, 6, 10, 12, 14, 16, 20, 26, 66},
    {4, 4, 4, 4, 6, 8, 12, 16, 20, 26, 34, 42, 12}, {4, 4, 4, 6, 6, 8, 10, 14, 18, 26, 32, 42, 18},
    {4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 32, 44, 12}, {4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 30, 40, 18},
    {4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 30, 40, 18}, {4, 4, 4, 6, 8, 10, 12, 14, 18, 24, 30, 40, 18},
    {8, 8, 8, 12, 16, 20, 24, 28, 36, 2, 2, 2, 26},
};
// Mixed blocks (layer3/common.rs:109-172): `n_long` long-band edges taken from the long table,
// then short bands starting with short band `first_short`.  The reference's 8 kHz row is its own
// "educated guess" (common.rs:159-167) and is reproduced as such: long edges 0,12,24, then the
// literal edges 36,40,44,48 followed by the regular short bands from band 2.
struct MixedRule { uint8_t n_long_edges; uint8_t first_short; };
const MixedRule kMixedRule[9] = {{8, 3}, {8, 3}, {8, 3}, {6, 3}, {6, 3}, {6, 3}, {6, 3}, {6, 3}, {3, 2}};
const uint8_t kPreEmphasis[22] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 2, 0};

// Numerators of the synthesis window, D[i] = n/65536 (ISO 11172-3 Table B.3).  The table has the
// structure D[i] = -D[512-i] for i not a multiple of 64 (odd symmetry around the centre tap); it
// is nevertheless stored in full so that no reconstruction rule can be wrong.
const int32_t kWindowNum[512] = {
#include "synth_window_num.inc"
};

void build_edges(Mp3Tables& t) {
    std::memset(t.edges, 0, sizeof t.edges);
    for (int sr = 0; sr < 9; ++sr) {
        uint16_t* L = t.edges[sr][kKindLong];
        L[0] = 0;
        for (int b = 0; b < 22; ++b) L[b + 1] = (uint16_t)(L[b] + kLongWidths[sr][b]);
        t.n_edges[sr][kKindLong] = 23;
        uint16_t* S = t.edges[sr][kKindShort];
        S[0] = 0;
        for (int b = 0; b < 13; ++b)
            for (int w = 0; w < 3; ++w) S[3 * b + w + 1] = (uint16_t)(S[3 * b + w] + kShortWidths[sr][b]);
        t.n_edges[sr][kKindShort] = 40;
        uint16_t* M = t.edges[sr][kKindMixed];
        const MixedRule r = kMixedRule[sr];
        int n = 0;
        for (int i = 0; i < r.n_long_edges; ++i) M[n++] = L[i];
        if (sr == 8) {
            const uint16_t guess[4] = {36, 40, 44, 48};
            for (uint16_t g : guess) M[n++] = g;
            for (int e = 3 * r.first_short + 1; e < 40; ++e) M[n++] = S[e];
        } else {
            for (int e = 3 * r.first_short; e < 40; ++e) M[n++] = S[e];
        }
        t.n_edges[sr][kKindMixed] = (uint8_t)n;
        t.mixed_switch[sr] = r.n_long_edges;
    }
    std::memset(t.pre_emphasis, 0, sizeof t.pre_emphasis);
    std::memcpy(t.pre_emphasis, kPreEmphasis, 22);
}

void build_line_maps(Mp3Tables& t) {
    for (int sr = 0; sr < 9; ++sr) {
        for (int kind = 0; kind < 3; ++kind) {
            const uint16_t* e = t.edges[sr][kind];
            const int n = t.n_edges[sr][kind];
            int iv = 0;
            for (int line = 0; line < 576; ++line) {
                while (iv + 2 < n && line >= e[iv + 1]) ++iv;
                t.iv_of_line[sr][kind][line] = (uint8_t)iv;
            }
        }
        // Reorder map (hybrid_synthesis.rs:153-215): within each short band the three windows
        // [w0..][w1..][w2..] are interleaved sample by sample.
        for (int m = 0; m < 2; ++m) {
            const int kind = m ? kKindMixed : kKindShort;
            const int sw = m ? t.mixed_switch[sr] : 0;
            const uint16_t* e = t.edges[sr][kind] + sw;
            const int n = t.n_edges[sr][kind] - sw;
            uint16_t* src = t.reorder_src[sr][m];
            for (int line = 0; line < 576; ++line) src[line] = (uint16_t)line;
            int i = e[0];
            t.reorder_start[sr][m] = e[0];
            for (int q = 0; q + 3 < n; q += 3) {
                const int len = e[q + 1] - e[q];
                if (e[q + 2] - e[q + 1] != len || e[q + 3] - e[q + 2] != len || i != e[q]) {
                    std::fprintf(stderr, "symgpu: short-band table is not 3 equal windows (sr=%d)\n", sr);
                    std::abort();
                }
                for (int k = 0; k < len; ++k)
                    for (int w = 0; w < 3; ++w) src[i++] = (uint16_t)(e[q + w] + k);
            }
            const uint8_t* ivm = t.iv_of_line[sr][kind];
            for (int line = 0; line < 576; ++line)
                t.short_map[sr][m][line] = (uint32_t)src[line] | ((uint32_t)ivm[src[line]] << 10) | ((uint32_t)ivm[line] << 16);
        }
    }
}

void build_float_tables(Mp3Tables& t) {
    const double PI = 3.14159265358979323846264338327950288;
    for (int i = 0; i < 512; ++i) {
        // The reference's literal is the 9-decimal rounding of n/65536 (synthesis.rs:13-142).
        char txt[32];
        std::snprintf(txt, sizeof txt, "%.9f", (double)kWindowNum[i] / 65536.0);
        t.synth_d[i] = std::strtof(txt, nullptr);
    }
    std::memset(t.imdct_win, 0, sizeof t.imdct_win);
    auto s36 = [&](int i) { return (float)std::sin(PI / 36.0 * ((double)i + 0.5)); };
    auto s12 = [&](int i) { return (float)std::sin(PI / 12.0 * ((double)i + 0.5)); };
    for (int i = 0; i < 36; ++i) t.imdct_win[0][i] = s36(i);           // hybrid_synthesis.rs:60-62
    for (int i = 0; i < 18; ++i) t.imdct_win[1][i] = s36(i);           // :65-73
    for (int i = 18; i < 24; ++i) t.imdct_win[1][i] = 1.0f;
    for (int i = 24; i < 30; ++i) t.imdct_win[1][i] = s12(i - 18);
    for (int i = 0; i < 12; ++i) t.imdct_win[2][i] = s12(i);           // :76-78
    for (int i = 6; i < 12; ++i) t.imdct_win[3][i] = s12(i - 6);       // :81-89
    for (int i = 12; i < 18; ++i) t.imdct_win[3][i] = 1.0f;
    for (int i = 18; i < 36; ++i) t.imdct_win[3][i] = s36(i);
    for (int i = 0; i < 6; ++i)                                        // :105-119
        for (int k = 0; k < 6; ++k)
            t.half_cos12[i][k] = (float)std::cos(PI / 24.0 * (double)((2 * (i + 3) + 7) * (2 * k + 1)));
    const double c[8] = {-0.6, -0.535, -0.33, -0.185, -0.095, -0.041, -0.0142, -0.0037}; // :136-149
    for (int i = 0; i < 8; ++i) {
        const double root = std::sqrt(1.0 + c[i] * c[i]);
        t.cs[i] = (float)(1.0 / root);
        t.ca[i] = (float)(c[i] / root);
    }
    for (int p = 0; p < 7; ++p) {                                      // stereo.rs:105-121
        const double ratio = std::tan(PI / 12.0 * (double)p);
        t.is_mpeg1[p][0] = (float)(ratio / (1.0 + ratio));
        t.is_mpeg1[p][1] = (float)(1.0 / (1.0 + ratio));
    }
    t.is_mpeg1[6][0] = 1.0f;
    t.is_mpeg1[6][1] = 0.0f;
    const double sqrt2 = 1.41421356237309504880168872420969808;
    const double i0[2] = {1.0 / std::sqrt(sqrt2), 0.707106781186547524400844362104849039};
    for (int s = 0; s < 2; ++s)                                        // stereo.rs:59-81
        for (int p = 0; p < 32; ++p) {
            const bool odd = p & 1;
            const float v = (float)std::pow(i0[s], (double)(odd ? p + 1 : p) / 2.0);
            t.is_mpeg2[s][p][0] = odd ? v : 1.0f;
            t.is_mpeg2[s][p][1] = odd ? 1.0f : v;
        }
    for (int m = 0; m < 18; ++m) t.dct_iv_scale[m] = (float)(2.0 * std::cos(PI * (2 * m + 1) / 72.0));
    for (int m = 0; m < 9; ++m) t.sdct18_scale[m] = (float)(2.0 * std::cos(PI * (2 * m + 1) / 36.0));
    t.sdct18_scale[4] = 1.41421356237309504880168872420969808f;
    const double ang[3] = {8.0 * PI / 9.0, 4.0 * PI / 9.0, 2.0 * PI / 9.0};
    t.sdct9_d[0] = (float)(-std::sqrt(3.0));
    for (int k = 0; k < 3; ++k) {
        t.sdct9_d[1 + k] = (float)(-2.0 * std::cos(ang[k]));
        t.sdct9_d[4 + k] = (float)(-2.0 * std::sin(ang[k]));
    }
    auto lee = [&](float* out, int half) {   // 1 / (2 cos(pi (2i+1) / (4 half)))
        for (int i = 0; i < half; ++i) out[i] = (float)(1.0 / (2.0 * std::cos(PI * (2 * i + 1) / (4.0 * half))));
    };
    lee(t.lee16, 16);
    lee(t.lee8, 8);
    lee(t.lee4, 4);
    lee(t.lee2, 2);
    t.lee1 = 0.707106781186547524400844362104849039f;
    for (int k = 0; k < kPow2qLen; ++k) t.pow2q[k] = (float)std::pow(2.0, 0.25 * (double)(k + kPow2qMin));
    for (int i = 0; i < 8208; ++i) t.pow43[i] = std::pow((float)i, 4.0f / 3.0f);
}

} // namespace

const Mp3Tables& mp3_tables_host() {
    static Mp3Tables* tab = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        tab = new Mp3Tables();
        std::memset(tab, 0, sizeof *tab);
        build_edges(*tab);
        build_line_maps(*tab);
        build_float_tables(*tab);
    });
    return *tab;
}

} // namespace symgpu

// =============================================================================================
// IMDCT codecs (AAC-LC, Vorbis)
// =============================================================================================
namespace symgpu {
namespace {

// floor1_inverse_dB_table of the Vorbis I specification as f32 bit patterns (the values the
// reference's 8-digit literals parse to, floor.rs:21-86).
const uint32_t kInverseDbBits[256] = {
    0x33e4b43e, 0x33f39109, 0x3401b28b, 0x340a203c, 0x34131a23, 0x341ca960, 0x3426d7a7, 0x3431af4b,
    0x343d3b50, 0x34498770, 0x3456a023, 0x346492b8, 0x34736d55, 0x34819f88, 0x348a0bfc, 0x34930493,
    0x349c9269, 0x34a6bf32, 0x34b1953f, 0x34bd1f93, 0x34c969e4, 0x34d680ad, 0x34e47136, 0x34f349a6,
    0x35018c88, 0x3509f7c0, 0x3512ef06, 0x351c7b76, 0x3526a6c0, 0x35317b37, 0x353d03da, 0x35494c5e,
    0x3556613b, 0x35644fb9, 0x357325fc, 0x3581798a, 0x3589e386, 0x3592d97c, 0x359c6485, 0x35a68e52,
    0x35b16133, 0x35bce825, 0x35c92edc, 0x35d641ce, 0x35e42e41, 0x35f30257, 0x3601668f, 0x3609cf4f,
    0x3612c3f5, 0x361c4d98, 0x362675e8, 0x36314732, 0x363ccc74, 0x3649115e, 0x36562265, 0x36640cce,
    0x3672deb8, 0x36815397, 0x3689bb1c, 0x3692ae72, 0x369c36af, 0x36a65d81, 0x36b12d35, 0x36bcb0c7,
    0x36c8f3e4, 0x36d60301, 0x36e3eb60, 0x36f2bb1e, 0x370140a2, 0x3709a6eb, 0x371298f1, 0x371c1fc9,
    0x3726451e, 0x3731133d, 0x373c951e, 0x3748d66f, 0x3755e3a2, 0x3763c9f7, 0x37729789, 0x37812daf,
    0x378992be, 0x37928374, 0x379c08e6, 0x37a62cbe, 0x37b0f947, 0x37bc7979, 0x37c8b8fe, 0x37d5c447,
    0x37e3a892, 0x37f273f8, 0x38011ac0, 0x38097e93, 0x38126df9, 0x381bf206, 0x38261462, 0x3830df56,
    0x383c5dd8, 0x38489b92, 0x3855a4f2, 0x38638733, 0x3872506e, 0x388107d3, 0x38896a6b, 0x38925882,
    0x389bdb2a, 0x38a5fc09, 0x38b0c568, 0x38bc423b, 0x38c87e29, 0x38d585a0, 0x38e365d9, 0x38f22ce8,
    0x3900f4e9, 0x39095646, 0x3912430e, 0x391bc451, 0x3925e3b5, 0x3930ab7f, 0x393c26a2, 0x394860c5,
    0x39556653, 0x39634483, 0x39720968, 0x3980e201, 0x39894224, 0x39922d9d, 0x399bad7b, 0x39a5cb63,
    0x39b09199, 0x39bc0b0d, 0x39c84366, 0x39d5470b, 0x39e32332, 0x39f1e5ed, 0x3a00cf1d, 0x3a092e05,
    0x3a121830, 0x3a1b96a9, 0x3a25b315, 0x3a3077b7, 0x3a3bef7c, 0x3a48260a, 0x3a5527c7, 0x3a6301e6,
    0x3a71c278, 0x3a80bc3b, 0x3a8919e9, 0x3a9202c6, 0x3a9b7fdb, 0x3aa59acb, 0x3ab05dd8, 0x3abbd3ef,
    0x3ac808b3, 0x3ad50888, 0x3ae2e09f, 0x3af19f07, 0x3b00a95c, 0x3b0905d0, 0x3b11ed5e, 0x3b1b690f,
    0x3b258284, 0x3b3043fd, 0x3b3bb867, 0x3b47eb61, 0x3b54e94d, 0x3b62bf5d, 0x3b717b9c, 0x3b80967f,
    0x3b88f1ba, 0x3b91d7f9, 0x3b9b5247, 0x3ba56a41, 0x3bb02a27, 0x3bbb9ce2, 0x3bc7ce12, 0x3bd4ca17,
    0x3be29e20, 0x3bf15835, 0x3c0083a6, 0x3c08dda7, 0x3c11c298, 0x3c1b3b82, 0x3c255201, 0x3c301054,
    0x3c3b8161, 0x3c47b0c8, 0x3c54aae5, 0x3c627ce8, 0x3c7134d4, 0x3c8070cf, 0x3c88c996, 0x3c91ad3a,
    0x3c9b24c0, 0x3ca539c5, 0x3caff685, 0x3cbb65e5, 0x3cc79382, 0x3cd48bb9, 0x3ce25bb4, 0x3cf11179,
    0x3d005dfb, 0x3d08b589, 0x3d1197df, 0x3d1b0e02, 0x3d25218d, 0x3d2fdcb9, 0x3d3b4a6d, 0x3d477640,
    0x3d546c91, 0x3d623a85, 0x3d70ee22, 0x3d804b2a, 0x3d88a17f, 0x3d918288, 0x3d9af748, 0x3da50958,
    0x3dafc2f2, 0x3dbb2ef8, 0x3dc75903, 0x3dd44d6d, 0x3de2195c, 0x3df0cad1, 0x3e00385b, 0x3e088d77,
    0x3e116d33, 0x3e1ae090, 0x3e24f127, 0x3e2fa92e, 0x3e3b1387, 0x3e473bca, 0x3e542e4d, 0x3e61f837,
    0x3e70a784, 0x3e80258f, 0x3e887973, 0x3e9157e2, 0x3e9ac9dc, 0x3ea4d8f9, 0x3eaf8f6d, 0x3ebaf81b,
    0x3ec71e95, 0x3ed40f33, 0x3ee1d717, 0x3ef0843d, 0x3f0012c6, 0x3f086572, 0x3f114293, 0x3f1ab32b,
    0x3f24c0ce, 0x3f2f75b1, 0x3f3adcb2, 0x3f470165, 0x3f53f01d, 0x3f61b5fb, 0x3f7060fb, 0x3f800000
};

void imdct_twiddles(Cplx* out, int n, double scale) { // mdct.rs:42-54
    const double PI = 3.14159265358979323846264338327950288;
    const int n2 = n / 2;
    const double alpha = 1.0 / 8.0 + (std::signbit(scale) ? (double)n2 : 0.0);
    const double pi_n = PI / (double)n;
    const double root = std::sqrt(std::fabs(scale));
    for (int k = 0; k < n2; ++k) {
        const double theta = pi_n * (alpha + (double)k);
        out[k].re = (float)(root * std::cos(theta));
        out[k].im = (float)(root * std::sin(theta));
    }
}

void fft_twiddles(Cplx* out, int size) { // no_simd.rs:16-36: (cos(pi k / half), -sin(pi k / half))
    const double PI = 3.14159265358979323846264338327950288;
    const int half = size / 2;
    const double theta = PI / (double)half;
    for (int k = 0; k < half; ++k) {
        const double angle = theta * (double)k;
        out[k].re = (float)std::cos(angle);
        out[k].im = (float)(-std::sin(angle));
    }
}

double bessel_i0(double x) { // aac/window.rs:56-63
    double val = 1.0;
    for (int n = 63; n >= 1; --n) {
        val *= x / (double)(n * n);
        val += 1.0;
    }
    return val;
}

void aac_sine_window(float* dst, int size) { // aac/window.rs:29-36: f32 arithmetic and f32 sin
    const float pi = 3.14159265358979323846264338327950288f;
    const float param = pi / (float)(2 * size);
    for (int n = 0; n < size; ++n) dst[n] = std::sin(((float)n + 0.5f) * param) * 1.0f;
}

void aac_kbd_window(float* dst, int size, float alpha) { // aac/window.rs:37-52
    const float pi = 3.14159265358979323846264338327950288f;
    const float dlen = (float)size;
    const float a = alpha * pi / dlen;
    const double alpha2 = (double)(a * a);
    std::vector<double> kb((size_t)size);
    double sum = 0.0;
    for (int n = 0; n < size; ++n) {
        sum += bessel_i0((double)(n * (size - n)) * alpha2);
        kb[(size_t)n] = sum;
    }
    sum += 1.0;
    for (int n = 0; n < size; ++n) dst[n] = (float)std::sqrt(kb[(size_t)n] / sum);
}

void vorbis_window(float* dst, int bs) { // codec-vorbis/src/window.rs:11-24
    const double half_pi = 1.57079632679489661923132169163975144;
    const int len = bs / 2;
    for (int i = 0; i < len; ++i) {
        const double frac = half_pi * (((double)i + 0.5) / (double)len);
        const double s = std::sin(frac);
        dst[i] = (float)std::sin(half_pi * (s * s));
    }
}

} // namespace

const CodecTables& codec_tables_host() {
    static CodecTables* tab = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        tab = new CodecTables();
        std::memset(tab, 0, sizeof *tab);
        fft_twiddles(tab->fft_lit16, 16);
        fft_twiddles(tab->fft_lit32, 32);
        for (int size = 64; size <= 2048; size <<= 1) fft_twiddles(tab->fft_merge + (size / 2 - 32), size);
        imdct_twiddles(tab->aac_tw_long, 1024, 1.0 / 2048.0);
        imdct_twiddles(tab->aac_tw_short, 128, 1.0 / 256.0);
        for (int n2 = 16; n2 <= 2048; n2 <<= 1) imdct_twiddles(tab->vorbis_tw + (n2 - 16), 2 * n2, 1.0);
        aac_sine_window(tab->aac_sine_long, 1024);
        aac_sine_window(tab->aac_sine_short, 128);
        aac_kbd_window(tab->aac_kbd_long, 1024, 4.0f);
        aac_kbd_window(tab->aac_kbd_short, 128, 6.0f);
        for (int bs = 64; bs <= 8192; bs <<= 1) vorbis_window(tab->vorbis_win + (bs / 2 - 32), bs);
        std::memcpy(tab->vorbis_inverse_db, kInverseDbBits, sizeof kInverseDbBits);
    });
    return *tab;
}

} // namespace symgpu
