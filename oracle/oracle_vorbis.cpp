// ORACLE (test infrastructure, NOT product code): CPU restatement of the Vorbis synthesis stage of
// pdeljanov/Symphonia @ ee35874:
//   Floor1::synthesis_step1 / _step2   symphonia-codec-vorbis/src/floor.rs:568-653
//   render_point / render_line         floor.rs:776-825   (integer; FLOOR1_INVERSE_DB_TABLE :21-86)
//   inverse coupling, dot product      symphonia-codec-vorbis/src/lib.rs:252-292
//   DspChannel::synth / overlap_add    symphonia-codec-vorbis/src/dsp.rs:68-145
//   generate_win_curve                 symphonia-codec-vorbis/src/window.rs:11-24
// on top of oracle_mdct.cpp.  floor0 (LSP) is out of scope.
//
// PARITY PINNING: the reference has no known-answer vector for this stage; the
// integer floor renderer is checked against an independent line-by-line restatement of the Vorbis I
// specification's render_line in tests/test_oracle_kat_vorbis.py, the IMDCT through the shared
// vectors, and the window/overlap-add against the spec formula in f64.  Bit-level agreement with
// the Rust binary is by construction.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

// floor1_inverse_dB_table of the Vorbis I specification, as the f32 bit patterns the reference's
// 8-digit literals parse to (floor.rs:21-86).
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
float inverse_db(int y) {
    float f;
    std::memcpy(&f, &kInverseDbBits[y], 4);
    return f;
}

int get_range(int multiplier) { // floor.rs:737-745
    static const int r[4] = {256, 128, 86, 64};
    return r[multiplier - 1];
}

int render_point(uint32_t x0, int y0, uint32_t x1, int y1, uint32_t x) { // floor.rs:776-782
    const int dy = y1 - y0;
    const uint32_t adx = x1 - x0;
    const uint32_t err = (uint32_t)std::abs(dy) * (x - x0);
    const uint32_t off = err / adx;
    return dy < 0 ? y0 - (int)off : y0 + (int)off;
}

void render_line(uint32_t x0, int y0, uint32_t x1, int y1, size_t n, float* v) { // floor.rs:785-825
    if ((size_t)x0 >= n) return;
    const int dy = y1 - y0;
    const int adx = (int)(x1 - x0);
    const int base = dy / adx;
    int y = y0;
    const int sy = dy < 0 ? base - 1 : base + 1;
    const int ady = std::abs(dy) - std::abs(base) * adx;
    v[x0] = inverse_db(y);
    int err = 0;
    const size_t x_begin = (size_t)x0 + 1;
    const size_t x_end = std::min(n, (size_t)x1);
    if (x_begin > x_end) return;
    for (size_t x = x_begin; x < x_end; ++x) {
        err += ady;
        if (err >= adx) {
            err -= adx;
            y += sy;
        } else {
            y += base;
        }
        v[x] = inverse_db(y);
    }
}

const float* window_for(int bs) { // window.rs:11-24, f64 then cast
    static std::mutex mu;
    static std::map<int, std::vector<float>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto& w = cache[bs];
    if (w.empty()) {
        const double FRAC_PI_2 = 1.57079632679489661923132169163975144;
        const int len = bs / 2;
        w.resize(len);
        for (int i = 0; i < len; ++i) {
            const double frac = FRAC_PI_2 * (((double)i + 0.5) / (double)len);
            const double s = std::sin(frac);
            w[i] = (float)std::sin(FRAC_PI_2 * (s * s));
        }
    }
    return w.data();
}

void overlap_add(float* out, const float* left, const float* right, const float* win, int len) { // dsp.rs:135-145
    for (int k = 0; k < len; ++k) out[k] = left[k] * win[len - 1 - k] + right[k] * win[k];
}

// DspChannel::synth (dsp.rs:68-126); `overlap` has bs1/2 valid entries.
void channel_synth(const float* spec, float* overlap, int bs0, int bs1, bool block_flag, bool prev_block_flag, float* buf) {
    const int bs = block_flag ? bs1 : bs0;
    std::vector<float> imdct(bs);
    oracle::imdct_for(bs / 2, 1.0).run(spec, imdct.data());
    const float* win = (block_flag && prev_block_flag) ? window_for(bs1) : window_for(bs0);
    if (prev_block_flag == block_flag) {
        overlap_add(buf, overlap, imdct.data(), win, bs / 2);
    } else if (prev_block_flag && !block_flag) {
        const int start = (bs1 - bs0) / 4, end = start + bs0 / 2;
        std::memcpy(buf, overlap, sizeof(float) * start);
        overlap_add(buf + start, overlap + start, imdct.data(), win, end - start);
    } else {
        const int start = (bs1 - bs0) / 4, end = start + bs0 / 2;
        overlap_add(buf, overlap, imdct.data() + start, win, bs0 / 2);
        std::memcpy(buf + bs0 / 2, imdct.data() + end, sizeof(float) * (bs1 / 2 - end));
    }
    std::memcpy(overlap, imdct.data() + bs / 2, sizeof(float) * (bs / 2));
}

} // namespace

extern "C" {

float oracle_vorbis_inverse_db(int i) { return inverse_db(i); }
const float* oracle_vorbis_window(int bs) { return window_for(bs); }

// Floor1::synthesis (floor.rs:728-733): step 1 (amplitude unwrap) then step 2 (curve render) for a
// half-block of n lines.
void oracle_vorbis_floor1(const symgpu_vorbis_floor1* s, const uint16_t* floor_y, uint32_t n, float* floor) {
    const int count = s->n_posts;
    int final_y[65];
    bool step2[65];
    const int range = get_range(s->multiplier);
    step2[0] = step2[1] = true;
    final_y[0] = floor_y[0];
    final_y[1] = floor_y[1];
    for (int i = 2; i < count; ++i) { // floor.rs:578-624
        const int lo = s->low[i], hi = s->high[i];
        const int predicted = render_point(s->x_list[lo], final_y[lo], s->x_list[hi], final_y[hi], s->x_list[i]);
        const int val = floor_y[i];
        const int highroom = range - predicted, lowroom = predicted;
        if (val != 0) {
            const int room = 2 * (highroom < lowroom ? highroom : lowroom);
            step2[lo] = step2[hi] = step2[i] = true;
            if (val >= room)
                final_y[i] = highroom > lowroom ? val - lowroom + predicted : predicted - val + highroom - 1;
            else
                final_y[i] = (val & 1) ? predicted - ((val + 1) / 2) : predicted + (val / 2);
        } else {
            step2[i] = false;
            final_y[i] = predicted;
        }
    }
    const int mult = s->multiplier; // floor.rs:627-653
    uint32_t hx = 0, lx = 0;
    int hy = 0;
    int ly = std::min(std::max(final_y[s->sort_order[0]] * mult, 0), 255);
    for (int k = 1; k < count; ++k) {
        const int i = s->sort_order[k];
        if (step2[i]) {
            hy = std::min(std::max(final_y[i] * mult, 0), 255);
            hx = s->x_list[i];
            render_line(lx, ly, hx, hy, n, floor);
            lx = hx;
            ly = hy;
        }
    }
    if (hx < n) render_line(hx, hy, n, hy, n, floor);
}

// Same contract as symgpu_vorbis_synth_host (include/symgpu.h).
int oracle_vorbis_batch(oracle_vorbis_state* states, const symgpu_vorbis_stream* streams,
                        const symgpu_vorbis_floor1* floors, const symgpu_vorbis_unit* units, const uint16_t* floor_y,
                        const float* residue, const symgpu_vorbis_run* runs, uint32_t n_runs, uint32_t slot, float* pcm,
                        int n_threads) {
    auto work = [&](uint32_t rlo, uint32_t rhi) {
        std::vector<float> fl[2], res[2];
        for (uint32_t r = rlo; r < rhi; ++r) {
            const symgpu_vorbis_stream& cfg = streams[runs[r].stream];
            oracle_vorbis_state& st = states[runs[r].stream];
            const int bs0 = 1 << cfg.bs0_exp, bs1 = 1 << cfg.bs1_exp;
            for (uint32_t p = runs[r].first_packet; p < runs[r].first_packet + runs[r].n_packets; ++p) {
                const symgpu_vorbis_unit& u = units[p];
                const int n = u.block_flag ? bs1 : bs0, n2 = n >> 1;
                for (int ch = 0; ch < cfg.channels; ++ch) {
                    fl[ch].assign(n2, 0.0f);
                    res[ch].assign(residue + ((size_t)p * 2 + ch) * slot, residue + ((size_t)p * 2 + ch) * slot + n2);
                    // lib.rs:189-208: a channel whose floor is unused gets a zero floor vector (encoded by
                    // the caller as floor index 0xffff); every other floor is rendered now.
                    if (u.floor[ch] != 0xffff)
                        oracle_vorbis_floor1(&floors[u.floor[ch]], floor_y + ((size_t)p * 2 + ch) * 65, (uint32_t)n2,
                                             fl[ch].data());
                }
                if (cfg.coupled && cfg.channels == 2) { // lib.rs:252-278
                    for (int i = 0; i < n2; ++i) {
                        const float m = res[0][i], a = res[1][i];
                        float nm, na;
                        if (m > 0.0f) {
                            if (a > 0.0f) { nm = m; na = m - a; } else { nm = m + a; na = m; }
                        } else {
                            if (a > 0.0f) { nm = m; na = m + a; } else { nm = m - a; na = m; }
                        }
                        res[0][i] = nm;
                        res[1][i] = na;
                    }
                }
                for (int ch = 0; ch < cfg.channels; ++ch) { // lib.rs:282-292
                    if (u.do_not_decode[ch]) continue;
                    for (int i = 0; i < n2; ++i) fl[ch][i] *= res[ch][i];
                }
                for (int ch = 0; ch < cfg.channels; ++ch)
                    channel_synth(fl[ch].data(), st.overlap[ch], bs0, bs1, u.block_flag, u.prev_block_flag,
                                  pcm + ((size_t)p * 2 + ch) * slot);
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

// VorbisDecoder::decode_inner from the coupling step on (lib.rs:250-315) for any channel count the reference maps (<= 8,
// lib.rs:771-788) and any number of coupling steps, which it applies in the order of the mapping (lib.rs:252-278).
int oracle_vorbis_mc_batch(oracle_vorbis_mc_state* states, const symgpu_vorbis_stream_mc* streams, const symgpu_vorbis_floor1* floors,
                           const symgpu_vorbis_unit_mc* units, const uint16_t* floor_y, const float* residue, const symgpu_vorbis_run* runs,
                           uint32_t n_runs, uint32_t channels, uint32_t slot, float* pcm) {
    std::vector<float> fl[SYMGPU_VORBIS_MAX_CHANNELS], res[SYMGPU_VORBIS_MAX_CHANNELS];
    for (uint32_t r = 0; r < n_runs; ++r) {
        const symgpu_vorbis_stream_mc& cfg = streams[runs[r].stream];
        oracle_vorbis_mc_state& st = states[runs[r].stream];
        const int bs0 = 1 << cfg.bs0_exp, bs1 = 1 << cfg.bs1_exp;
        if (cfg.channels > channels || cfg.channels > SYMGPU_VORBIS_MAX_CHANNELS) return 1;
        for (uint32_t p = runs[r].first_packet; p < runs[r].first_packet + runs[r].n_packets; ++p) {
            const symgpu_vorbis_unit_mc& u = units[p];
            const int n = u.block_flag ? bs1 : bs0, n2 = n >> 1;
            for (int ch = 0; ch < cfg.channels; ++ch) {
                const size_t plane = (size_t)p * channels + ch;
                fl[ch].assign(n2, 0.0f);
                res[ch].assign(residue + plane * slot, residue + plane * slot + n2);
                if (u.floor[ch] != 0xffff) oracle_vorbis_floor1(&floors[u.floor[ch]], floor_y + plane * 65, (uint32_t)n2, fl[ch].data());
            }
            for (int c = 0; c < cfg.n_couplings; ++c) { // lib.rs:252-278
                std::vector<float>& mag = res[cfg.magnitude_ch[c]];
                std::vector<float>& ang = res[cfg.angle_ch[c]];
                for (int i = 0; i < n2; ++i) {
                    const float m = mag[i], a = ang[i];
                    float nm, na;
                    if (m > 0.0f) {
                        if (a > 0.0f) { nm = m; na = m - a; } else { nm = m + a; na = m; }
                    } else {
                        if (a > 0.0f) { nm = m; na = m + a; } else { nm = m - a; na = m; }
                    }
                    mag[i] = nm;
                    ang[i] = na;
                }
            }
            for (int ch = 0; ch < cfg.channels; ++ch) { // lib.rs:282-292
                if (u.do_not_decode[ch]) continue;
                for (int i = 0; i < n2; ++i) fl[ch][i] *= res[ch][i];
            }
            for (int ch = 0; ch < cfg.channels; ++ch)
                channel_synth(fl[ch].data(), st.overlap[ch], bs0, bs1, u.block_flag, u.prev_block_flag, pcm + ((size_t)p * channels + ch) * slot);
        }
    }
    return 0;
}
}
