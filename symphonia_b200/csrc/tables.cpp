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
