// Fused MPEG Layer III synthesis kernel for sm_100a:
//   requantize -> joint stereo -> reorder -> antialias -> IMDCT-36/12 + window + overlap-add
//   -> frequency inversion -> DCT-32 -> 512-tap polyphase window  (layer3/mod.rs:421-477)
// in ONE launch, PCM written straight to HBM.  No intermediate ever leaves the SM.
//
// Parallelisation: every piece of cross-granule state on this path is overwritten,
// never accumulated (hybrid overlap, polyphase FIFO), so a stream is cut into TILES of T
// consecutive granules; one CTA owns one tile.  A tile that starts a run loads the stream state
// from HBM; any other tile recomputes a 2-granule halo (the overlap of granule g-1 needs IMDCT of
// g-1; the 15 polyphase history slots need the time samples of g-1, which need the overlap of
// g-2).
//
// Bit-exactness rules: compiled with -fmad=false; every expression keeps the reference's operand
// order; tables come from the host (tables.cpp).  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2
// into FFMA2 even with --fmad=false, so packed f32x2 arithmetic is NOT used on mul->add chains;
// tests/test_build.py greps the SASS of this file for FFMA/FFMA2 and fails on any hit.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/symgpu.h"
#include "mp3_kernel.h"
#include "tables.h"

namespace symgpu {

// Uniformly indexed coefficient tables live in constant memory (c[3][...] operands cost no issue
// slot); lane-indexed tables stay in global memory behind the read-only path.
struct Mp3Const {
    float imdct_win[4][36];
    float half_cos12[6][6];
    float dct_iv_scale[18];
    float sdct18_scale[9];
    float sdct9_d[7];
    float lee16[16], lee8[8], lee4[4], lee2[2], lee1;
    float is_mpeg1[7][2];
    float is_mpeg2[2][32][2];
    uint8_t pre_emphasis[24];
    uint8_t mixed_switch[12];
    uint8_t n_edges[9][3];
};
__constant__ Mp3Const c_mp3;

cudaError_t mp3_upload_const(const Mp3Tables& t, cudaStream_t stream) {
    static Mp3Const h; // staging must outlive the async copy
    memcpy(h.imdct_win, t.imdct_win, sizeof h.imdct_win);
    memcpy(h.half_cos12, t.half_cos12, sizeof h.half_cos12);
    memcpy(h.dct_iv_scale, t.dct_iv_scale, sizeof h.dct_iv_scale);
    memcpy(h.sdct18_scale, t.sdct18_scale, sizeof h.sdct18_scale);
    memcpy(h.sdct9_d, t.sdct9_d, sizeof h.sdct9_d);
    memcpy(h.lee16, t.lee16, sizeof h.lee16);
    memcpy(h.lee8, t.lee8, sizeof h.lee8);
    memcpy(h.lee4, t.lee4, sizeof h.lee4);
    memcpy(h.lee2, t.lee2, sizeof h.lee2);
    h.lee1 = t.lee1;
    memcpy(h.is_mpeg1, t.is_mpeg1, sizeof h.is_mpeg1);
    memcpy(h.is_mpeg2, t.is_mpeg2, sizeof h.is_mpeg2);
    memcpy(h.pre_emphasis, t.pre_emphasis, sizeof h.pre_emphasis);
    memset(h.mixed_switch, 0, sizeof h.mixed_switch);
    memcpy(h.mixed_switch, t.mixed_switch, 9);
    memcpy(h.n_edges, t.n_edges, sizeof h.n_edges);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_mp3, &h, sizeof h, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

namespace {

constexpr int kPitch = 33; // float2 per slot row: 32 sub-bands + one always-zero column (V[16] = 0)
constexpr float kFrac1Sqrt2 = 0.707106781186547524400844362104849039f;

struct WarpScratch {
    symgpu_mp3_gc gc[2];
    float scale[2][40];
    float2 sratio[40];
    uint8_t smode[40]; // 0 none, 1 mid/side, 2 intensity
    uint8_t nz[40];    // channel-1 interval holds a non-zero line
};

__device__ __forceinline__ int kind_of(const symgpu_mp3_gc& g) {
    if (g.block_type != SYMGPU_MP3_SHORT) return kKindLong;
    return (g.flags & SYMGPU_MP3_F_MIXED) ? kKindMixed : kKindShort;
}

// ---- 9-point SDCT-II (hybrid_synthesis.rs:721-779); y[j] is the reference's y[2j] ----------
__device__ __forceinline__ void sdct9(const float (&x)[9], float (&y)[9]) {
    const float a01 = x[3] + x[5], a02 = x[3] - x[5], a03 = x[6] + x[2], a04 = x[6] - x[2];
    const float a05 = x[1] + x[7], a06 = x[1] - x[7], a07 = x[8] + x[0], a08 = x[8] - x[0];
    const float a09 = x[4] + a05, a10 = a01 + a03, a11 = a10 + a07, a12 = a03 - a07;
    const float a13 = a01 - a07, a14 = a01 - a03, a15 = a02 - a04, a16 = a15 + a08;
    const float a17 = a04 + a08, a18 = a02 - a08, a19 = a02 + a04, a20 = 2.0f * x[4] - a05;
    const float m1 = c_mp3.sdct9_d[0] * a06, m2 = c_mp3.sdct9_d[1] * a12, m3 = c_mp3.sdct9_d[2] * a13;
    const float m4 = c_mp3.sdct9_d[3] * a14, m5 = c_mp3.sdct9_d[0] * a16, m6 = c_mp3.sdct9_d[4] * a17;
    const float m7 = c_mp3.sdct9_d[5] * a18, m8 = c_mp3.sdct9_d[6] * a19;
    const float a21 = a20 + m2, a22 = a20 - m2, a23 = a20 + m3, a24 = m1 + m6, a25 = m1 - m6, a26 = m1 + m7;
    y[0] = a09 + a11;
    y[1] = m8 - a26;
    y[2] = m4 - a21;
    y[3] = m5;
    y[4] = a22 - m3;
    y[5] = a25 - m7;
    y[6] = a11 - 2.0f * a09;
    y[7] = a24 + m8;
    y[8] = a23 + m4;
}

// ---- 18-point DCT-IV via two 9-point SDCT-IIs (hybrid_synthesis.rs:608-716) -----------------
__device__ __forceinline__ void dct_iv_18(const float (&x)[18], float (&y)[18]) {
    float s[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) s[i] = c_mp3.dct_iv_scale[i] * x[i];
    float even[9], odd[9], ye[9], yo[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) even[i] = s[i] + s[17 - i];
    sdct9(even, ye);
#pragma unroll
    for (int i = 0; i < 9; ++i) odd[i] = c_mp3.sdct18_scale[i] * (s[i] - s[17 - i]);
    sdct9(odd, yo);
#pragma unroll
    for (int j = 1; j < 9; ++j) yo[j] = yo[j] - yo[j - 1]; // y[3]-=y[1]; y[5]-=y[3]; ... sequential
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        y[2 * j] = ye[j];
        y[2 * j + 1] = yo[j];
    }
    y[0] = y[0] * 0.5f; // "/ 2.0" -- exact either way
#pragma unroll
    for (int i = 1; i < 18; ++i) y[i] = (y[i] * 0.5f) - y[i - 1];
}

// imdct36 (hybrid_synthesis.rs:571-603) without the overlap add: first = windowed samples 0..17,
// second = windowed samples 18..35 (the next granule's overlap).
__device__ __forceinline__ void imdct36(const float (&x)[18], const float* __restrict__ win,
                                        float (&first)[18], float (&second)[18]) {
    float dct[18];
    dct_iv_18(x, dct);
#pragma unroll
    for (int i = 0; i < 9; ++i) first[i] = dct[9 + i] * win[i];
#pragma unroll
    for (int i = 9; i < 18; ++i) first[i] = -(dct[26 - i] * win[i]); // overlap - d*w == overlap + (-(d*w))
#pragma unroll
    for (int i = 18; i < 27; ++i) second[i - 18] = -dct[26 - i] * win[i];
#pragma unroll
    for (int i = 27; i < 36; ++i) second[i - 18] = -dct[i - 27] * win[i];
}

// imdct12_win (hybrid_synthesis.rs:363-455) without the overlap add.
__device__ __forceinline__ void imdct12x3(const float (&x)[18], float (&first)[18], float (&second)[18]) {
    float tmp[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) tmp[i] = 0.0f;
    const float* win = c_mp3.imdct_win[2];
#pragma unroll
    for (int w = 0; w < 3; ++w) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float yl = (x[w] * c_mp3.half_cos12[i][0]) + (x[3 + w] * c_mp3.half_cos12[i][1]) +
                             (x[6 + w] * c_mp3.half_cos12[i][2]) + (x[9 + w] * c_mp3.half_cos12[i][3]) +
                             (x[12 + w] * c_mp3.half_cos12[i][4]) + (x[15 + w] * c_mp3.half_cos12[i][5]);
            const float yr = (x[w] * c_mp3.half_cos12[i + 3][0]) + (x[3 + w] * c_mp3.half_cos12[i + 3][1]) +
                             (x[6 + w] * c_mp3.half_cos12[i + 3][2]) + (x[9 + w] * c_mp3.half_cos12[i + 3][3]) +
                             (x[12 + w] * c_mp3.half_cos12[i + 3][4]) + (x[15 + w] * c_mp3.half_cos12[i + 3][5]);
            tmp[6 + 6 * w + 2 - i] += -yl * win[2 - i];
            tmp[6 + 6 * w + i + 3] += yl * win[i + 3];
            tmp[6 + 6 * w + i + 6] += yr * win[i + 6];
            tmp[6 + 6 * w + 11 - i] += yr * win[11 - i];
        }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        first[i] = tmp[i];
        second[i] = tmp[i + 18];
    }
}

// ---- Lee 32-point DCT (synthesis.rs:348-844) as the recursion the reference hand-flattens ----
template <int N> struct LeeCoef;
template <> struct LeeCoef<16> { static __device__ __forceinline__ float at(int i) { return c_mp3.lee16[i]; } };
template <> struct LeeCoef<8> { static __device__ __forceinline__ float at(int i) { return c_mp3.lee8[i]; } };
template <> struct LeeCoef<4> { static __device__ __forceinline__ float at(int i) { return c_mp3.lee4[i]; } };
template <> struct LeeCoef<2> { static __device__ __forceinline__ float at(int i) { return c_mp3.lee2[i]; } };

template <int N>
__device__ __forceinline__ void lee_dct(const float (&x)[N], float (&y)[N]) {
    if constexpr (N == 2) {
        y[0] = x[0] + x[1];
        y[1] = (x[0] - x[1]) * c_mp3.lee1;
    } else {
        constexpr int H = N / 2;
        float lo[H], hi[H], lo_t[H], hi_t[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            lo[i] = x[i] + x[N - 1 - i];
            hi[i] = (x[i] - x[N - 1 - i]) * LeeCoef<H>::at(i);
        }
        lee_dct<H>(lo, lo_t);
        lee_dct<H>(hi, hi_t);
#pragma unroll
        for (int i = 0; i < H - 1; ++i) {
            y[2 * i] = lo_t[i];
            y[2 * i + 1] = hi_t[i] + hi_t[i + 1];
        }
        y[N - 2] = lo_t[H - 1];
        y[N - 1] = hi_t[H - 1];
    }
}

// Sign of the frequency inversion (hybrid_synthesis.rs:458-485): odd sample of odd sub-band.
__device__ __forceinline__ float finv(float v, int sb, int t) { return ((sb & t) & 1) ? -v : v; }

} // namespace

// =============================================================================================
template <int T, int NW>
__global__ void __launch_bounds__(NW * 32, 2) mp3_synth_kernel(Mp3Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xt = reinterpret_cast<float*>(smem_raw);                       // [(T+2)*18][kPitch][2]
    WarpScratch* wscr = reinterpret_cast<WarpScratch*>(smem_raw + (size_t)(T + 2) * 18 * kPitch * 8);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const Mp3Tile tile = a.tiles[blockIdx.x];
    const int n = tile.n_granules;
    const int n_ch = tile.n_ch;
    const int gpf = tile.gpf;
    const bool load_state = tile.flags & kTileLoadState;
    const bool store_state = tile.flags & kTileStoreState;
    const Mp3Tables* __restrict__ tab = a.tab;
    // Stream state is double-buffered: a launch reads generation g and writes generation g+1, so a
    // run-starting tile never races with the run-ending tile of the same stream (they are
    // different CTAs of the same launch).  The last CTA to finish bumps the generations.
    const uint32_t gen = a.gen[tile.stream];
    const Mp3StreamState* st_in = a.states + (size_t)tile.stream * 2 + (gen & 1);
    Mp3StreamState* st = a.states + (size_t)tile.stream * 2 + ((gen + 1) & 1);
    const int gseq0 = (int)tile.first_frame * gpf + tile.first_gr; // batch granule sequence index of region 2

    // ------------------------------------------------------------------------------------------
    // Phase A+B: one warp per granule ("job"), regions processed in descending rounds so that the
    // overlap hand-off into region r+1 always targets a region whose own samples are complete.
    // ------------------------------------------------------------------------------------------
    const int r_lo = load_state ? 2 : 0;
    const int n_regions = n + 2;
    WarpScratch& ws = wscr[warp];
    const float cs_l = tab->cs[lane & 7], ca_l = tab->ca[lane & 7];

    for (int r_hi = n_regions; r_hi > r_lo; r_hi -= NW) {
        const int r = r_hi - 1 - warp; // region of this warp's job in this round
        const bool active = r >= r_lo;
        float sec[2][18];
        float* S = xt + (size_t)r * 18 * kPitch * 2; // region memory doubles as the job's [2][576] scratch
        if (active) {
            const int gseq = gseq0 + (r - 2);
            const int frame = gseq / gpf, gr = gseq - frame * gpf;
            const symgpu_mp3_gc* u = a.units + ((size_t)frame * 2 + gr) * 2;
            const float4* spec4 = reinterpret_cast<const float4*>(a.spectra + ((size_t)frame * 2 + gr) * 2 * 576);

            // A0: descriptors -> shared (two 64-byte units)
            if (lane < 8) reinterpret_cast<uint4*>(ws.gc)[lane] = __ldg(reinterpret_cast<const uint4*>(u) + lane);
            if (lane < 10) reinterpret_cast<uint32_t*>(ws.smode)[lane] = 0;
            if (lane >= 16 && lane < 26) reinterpret_cast<uint32_t*>(ws.nz)[lane - 16] = 0;
            __syncwarp();
            const symgpu_mp3_gc& g0 = ws.gc[0];
            const symgpu_mp3_gc& g1 = ws.gc[1];
            const int sr = g0.sample_rate_idx;
            const int kind0 = kind_of(g0), kind1 = (n_ch == 2) ? kind_of(g1) : kind0;
            const bool ms = (n_ch == 2) && (g0.flags & SYMGPU_MP3_F_MID_SIDE);
            const bool is = (n_ch == 2) && (g0.flags & SYMGPU_MP3_F_INTENSITY);
            int rz0 = g0.rzero, rz1 = (n_ch == 2) ? g1.rzero : 0;

            // A1: per-interval requantisation scale (requantize.rs:240-355)
            for (int ch = 0; ch < n_ch; ++ch) {
                const symgpu_mp3_gc& g = ws.gc[ch];
                const int kind = ch ? kind1 : kind0;
                const int n_iv = c_mp3.n_edges[sr][kind] - 1;
                const int gain = (int)g.global_gain - 210;
                const int shift = (g.flags & SYMGPU_MP3_F_SCALEFAC_SCALE) ? 2 : 1;
                const int sw = c_mp3.mixed_switch[sr];
                for (int idx = lane; idx < 40; idx += 32) {
                    float s = 1.0f;
                    if (idx < n_iv) {
                        int e = 0;
                        bool scaled = true;
                        const bool long_part = (kind == kKindLong) || (kind == kKindMixed && idx < sw - 1);
                        if (long_part) {
                            const int pre = (g.flags & SYMGPU_MP3_F_PREFLAG) ? c_mp3.pre_emphasis[idx] : 0;
                            const int b = ((g.scalefacs[idx] + pre) << shift) & 0xff;
                            e = gain - b;
                        } else if (kind == kKindMixed && idx == sw - 1) {
                            scaled = false; // lines between the last long band and the first short band
                        } else {
                            const int j = (kind == kKindMixed) ? idx - sw : idx;
                            const int sfi = (kind == kKindMixed) ? idx : idx; // scalefacs[switch + j]
                            const int b = (g.scalefacs[sfi] << shift) & 0xff;
                            e = gain - 8 * (int)g.subblock_gain[j % 3] - b;
                        }
                        if (scaled) s = __ldg(&tab->pow2q[e - kPow2qMin]);
                    }
                    ws.scale[ch][idx] = s;
                }
            }
            __syncwarp();

            // A2: load spectra, requantise, record channel-1 non-zero intervals
            const uchar4* ivm0 = reinterpret_cast<const uchar4*>(tab->iv_of_line[sr][kind0]);
            const uchar4* ivm1 = reinterpret_cast<const uchar4*>(tab->iv_of_line[sr][kind1]);
#pragma unroll 3
            for (int it = 0; it < 9; ++it) {
                const int q = it * 32 + lane; // float4 index in [2][144]
                const int ch = q >= 144;
                const int l4 = q - ch * 144;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ch < n_ch) {
                    v = __ldg(spec4 + q);
                    const uchar4 iv = __ldg((ch ? ivm1 : ivm0) + l4);
                    // Lines at or beyond rzero are +0.0 by contract (requantize.rs:234), and +0 * scale is
                    // +0, so the reference's "stop at rzero" (requantize.rs:267,:284) needs no predicate.
                    v.x *= ws.scale[ch][iv.x];
                    v.y *= ws.scale[ch][iv.y];
                    v.z *= ws.scale[ch][iv.z];
                    v.w *= ws.scale[ch][iv.w];
                    if (ch && is) {
                        if (v.x != 0.0f) ws.nz[iv.x] = 1;
                        if (v.y != 0.0f) ws.nz[iv.y] = 1;
                        if (v.z != 0.0f) ws.nz[iv.z] = 1;
                        if (v.w != 0.0f) ws.nz[iv.w] = 1;
                    }
                }
                reinterpret_cast<float4*>(S)[q] = v;
            }
            __syncwarp();

            // A3/A4: joint stereo (stereo.rs:485-556)
            if (ms || is) {
                const int end = max(rz0, rz1);
                int bound = end;
                if (is) {
                    // Warp-parallel restatement of the two top-down scans (stereo.rs:198-261, :265-482).
                    // Every lane derives the same facts from one 40-bit "interval of channel 1 is
                    // non-zero" mask; lane iv then labels interval iv (and iv + 32).
                    const bool mpeg1 = g1.flags & SYMGPU_MP3_F_MPEG1;
                    const int inv_pos = mpeg1 ? 7 : 31;
                    const float(*rt)[2] = mpeg1 ? c_mp3.is_mpeg1 : c_mp3.is_mpeg2[(g1.flags & SYMGPU_MP3_F_SFC_LSB) ? 1 : 0];
                    const uint16_t* e = tab->edges[sr][kind1];
                    const int n_e = c_mp3.n_edges[sr][kind1];
                    const int n_iv = n_e - 1;
                    const uint8_t mode_hi = ms ? 1 : 0;
                    bool nza = false, nzb = false;
                    if (lane < n_iv) nza = ws.nz[lane] && (kind1 != kKindLong || (int)e[lane] < rz1);
                    if (lane + 32 < n_iv) nzb = ws.nz[lane + 32] && (kind1 != kKindLong || (int)e[lane + 32] < rz1);
                    const unsigned long long nzmask = (unsigned long long)__ballot_sync(0xffffffffu, nza) |
                                                      ((unsigned long long)__ballot_sync(0xffffffffu, nzb) << 32);
                    // is_lo: first interval labelled by the scan; iv_is[w]: first interval of window w that is
                    // intensity coded (every labelled interval of that window at or above it is).
                    int is_lo, first_is0, first_is1, first_is2;
                    if (kind1 == kKindLong) {
                        const int hb = nzmask ? 63 - __clzll((long long)nzmask) : -1; // highest non-zero band
                        is_lo = hb + 1;
                        first_is0 = first_is1 = first_is2 = hb + 1;
                        if (hb < 21) bound = e[hb + 1];
                    } else {
                        const int sw = (kind1 == kKindMixed) ? c_mp3.mixed_switch[sr] : 0;
                        const int n_quads = (n_e - sw - 1) / 3;
                        int hq0 = -1, hq1 = -1, hq2 = -1; // highest quad whose window w is non-zero
                        for (int q = 0; q < n_quads; ++q) {
                            const unsigned bits = (unsigned)(nzmask >> (sw + 3 * q)) & 7u;
                            if (bits & 1u) hq0 = q;
                            if (bits & 2u) hq1 = q;
                            if (bits & 4u) hq2 = q;
                        }
                        const int qstop = min(hq0, min(hq1, hq2)); // quad where all three windows are done, or -1
                        const int qlo = max(qstop, 0);
                        is_lo = sw + 3 * qlo;
                        first_is0 = sw + 3 * (hq0 + 1);
                        first_is1 = sw + 3 * (hq1 + 1) + 1;
                        first_is2 = sw + 3 * (hq2 + 1) + 2;
                        bound = e[is_lo];
                        if (qstop < 0 && kind1 == kKindMixed) { // continue into the long bands of a mixed block
                            const unsigned long long lmask = nzmask & ((1ull << sw) - 1ull);
                            const int hb = lmask ? 63 - __clzll((long long)lmask) : -1;
                            if (hb < sw - 1) {
                                is_lo = hb + 1;
                                bound = e[hb + 1];
                            }
                        }
                    }
                    for (int iv = lane; iv < n_iv; iv += 32) {
                        if (iv < is_lo) continue;
                        bool coded;
                        if (kind1 == kKindLong) {
                            coded = true;
                        } else {
                            const int sw = (kind1 == kKindMixed) ? c_mp3.mixed_switch[sr] : 0;
                            if (iv < sw) coded = true; // mixed long band reached by the scan: always zero
                            else {
                                const int w = (iv - sw) % 3;
                                coded = iv >= (w == 0 ? first_is0 : w == 1 ? first_is1 : first_is2);
                            }
                        }
                        uint8_t mode = mode_hi;
                        if (coded) {
                            // is_pos: long: scalefacs[b], band 21 copies band 20 (stereo.rs:228-230);
                            // short: scalefacs[k], the last three copy [33..36) (stereo.rs:352-354).
                            const int k = (kind1 == kKindLong) ? (iv == 21 ? 20 : iv) : (iv < 36 ? iv : iv - 3);
                            const int pos = g1.scalefacs[k];
                            if (pos < inv_pos) { // process_intensity, stereo.rs:168-188
                                mode = 2;
                                ws.sratio[iv] = make_float2(rt[pos][0], rt[pos][1]);
                            }
                        }
                        ws.smode[iv] = mode;
                    }
                }
                __syncwarp();
                if (!is) {
                    // mid/side only: every line below max(rzero) (stereo.rs:536-544)
                    for (int l4 = lane; l4 * 4 < bound; l4 += 32) {
                        float4 m = reinterpret_cast<float4*>(S)[l4];
                        float4 d = reinterpret_cast<float4*>(S + 576)[l4];
                        const int l = l4 * 4;
                        float4 L, R;
                        L.x = (m.x + d.x) * kFrac1Sqrt2; R.x = (m.x - d.x) * kFrac1Sqrt2;
                        L.y = (m.y + d.y) * kFrac1Sqrt2; R.y = (m.y - d.y) * kFrac1Sqrt2;
                        L.z = (m.z + d.z) * kFrac1Sqrt2; R.z = (m.z - d.z) * kFrac1Sqrt2;
                        L.w = (m.w + d.w) * kFrac1Sqrt2; R.w = (m.w - d.w) * kFrac1Sqrt2;
                        if (l + 3 >= bound) { // rzero is not always a multiple of 4: keep the tail untouched
                            if (l + 1 >= bound) { L.y = m.y; R.y = d.y; }
                            if (l + 2 >= bound) { L.z = m.z; R.z = d.z; }
                            L.w = m.w; R.w = d.w;
                        }
                        reinterpret_cast<float4*>(S)[l4] = L;
                        reinterpret_cast<float4*>(S + 576)[l4] = R;
                    }
                } else {
                for (int l4 = lane; l4 < 144; l4 += 32) {
                    float4 m = reinterpret_cast<float4*>(S)[l4];
                    float4 s = reinterpret_cast<float4*>(S + 576)[l4];
                    const uchar4 iv = __ldg(ivm1 + l4);
                    const int l = l4 * 4;
                    auto apply = [&](float& x, float& y, int line, int ivx) {
                        int mode;
                        if (line < bound) mode = ms ? 1 : 0; else mode = ws.smode[ivx];
                        if (mode == 1) { // process_mid_side, stereo.rs:143-152
                            const float left = (x + y) * kFrac1Sqrt2;
                            const float right = (x - y) * kFrac1Sqrt2;
                            x = left;
                            y = right;
                        } else if (mode == 2) {
                            const float2 rt = ws.sratio[ivx];
                            const float isv = x;
                            x = rt.x * isv;
                            y = rt.y * isv;
                        }
                    };
                    apply(m.x, s.x, l + 0, iv.x);
                    apply(m.y, s.y, l + 1, iv.y);
                    apply(m.z, s.z, l + 2, iv.z);
                    apply(m.w, s.w, l + 3, iv.w);
                    reinterpret_cast<float4*>(S)[l4] = m;
                    reinterpret_cast<float4*>(S + 576)[l4] = s;
                }
                }
                rz0 = end;
                rz1 = end;
                __syncwarp();
            }

            // A5: reorder short blocks (hybrid_synthesis.rs:153-215)
            for (int ch = 0; ch < n_ch; ++ch) {
                const int kind = ch ? kind1 : kind0;
                if (kind == kKindLong) continue;
                const int m = (kind == kKindMixed) ? 1 : 0;
                const int sw = m ? c_mp3.mixed_switch[sr] : 0;
                const uint16_t* e = tab->edges[sr][kind] + sw;
                const int n_quads = (c_mp3.n_edges[sr][kind] - sw - 1) / 3;
                int rz = ch ? rz1 : rz0;
                const bool below = (lane < n_quads) && ((int)e[3 * lane] < rz);
                const int n_done = __popc(__ballot_sync(0xffffffffu, below)); // quads form a prefix
                const int start = e[0];
                const int i_end = e[3 * n_done];
                const uint16_t* src = tab->reorder_src[sr][m];
                float* Sc = S + ch * 576;
                float tmp[18];
#pragma unroll
                for (int k = 0; k < 18; ++k) {
                    const int d = k * 32 + lane;
                    tmp[k] = (d >= start && d < i_end) ? Sc[__ldg(src + d)] : 0.0f;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 18; ++k) {
                    const int d = k * 32 + lane;
                    if (d >= start && d < i_end) Sc[d] = tmp[k];
                }
                rz = max(rz, i_end);
                if (ch) rz1 = rz; else rz0 = rz;
                __syncwarp();
            }

            // A6: antialias (hybrid_synthesis.rs:218-277)
            for (int ch = 0; ch < n_ch; ++ch) {
                const int kind = ch ? kind1 : kind0;
                if (kind == kKindShort) continue;
                const int sb_limit = (kind == kKindMixed) ? 2 : 32;
                int rz = ch ? rz1 : rz0;
                rz = 18 * min(min(sb_limit, rz / 18 + 2), 32);
                float* Sc = S + ch * 576;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int p = it * 32 + lane; // (boundary-1)*8 + i
                    const int sbb = 18 * ((p >> 3) + 1);
                    if (sbb < rz) { // p < 248 always holds when sbb <= 558
                        const int i = p & 7;
                        const int li = sbb - 1 - i, ui = sbb + i;
                        const float lower = Sc[li], upper = Sc[ui];
                        Sc[li] = lower * cs_l - upper * ca_l;
                        Sc[ui] = upper * cs_l + lower * ca_l;
                    }
                }
                if (ch) rz1 = rz; else rz0 = rz;
            }
            __syncwarp();

            // B: hybrid synthesis, lane = sub-band (hybrid_synthesis.rs:280-359)
            float xin[2][18];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                for (int i = 0; i < 18; i += 2) {
                    const float2 v = *reinterpret_cast<const float2*>(S + ch * 576 + 18 * lane + i);
                    xin[ch][i] = v.x;
                    xin[ch][i + 1] = v.y;
                }
            __syncwarp(); // every lane holds its inputs; the scratch may now be overwritten
            float* X = xt + (size_t)r * 18 * kPitch * 2;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                float first[18];
                if (ch < n_ch) {
                    const symgpu_mp3_gc& g = ws.gc[ch];
                    const int kind = ch ? kind1 : kind0;
                    const int rz = ch ? rz1 : rz0;
                    const int sb_limit = (rz + 17) / 18;
                    const int sb_split = (kind == kKindShort) ? 0 : (kind == kKindMixed) ? 2 : 32;
                    const int long_end = min(sb_split, sb_limit);
                    if (lane < long_end) {
                        const int wsel = g.block_type == SYMGPU_MP3_START ? 1 : g.block_type == SYMGPU_MP3_END ? 3 : 0;
                        imdct36(xin[ch], c_mp3.imdct_win[wsel], first, sec[ch]);
                    } else if (lane < sb_limit) {
                        imdct12x3(xin[ch], first, sec[ch]);
                    } else {
                        // samples = overlap; overlap = 0  (:351-358).  overlap + (-0.0) == overlap bit for bit.
#pragma unroll
                        for (int i = 0; i < 18; ++i) {
                            first[i] = -0.0f;
                            sec[ch][i] = 0.0f;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 18; ++i) {
                        first[i] = 0.0f;
                        sec[ch][i] = 0.0f;
                    }
                }
#pragma unroll
                for (int t = 0; t < 18; ++t) X[(t * kPitch + lane) * 2 + ch] = finv(first[t], lane, t);
            }
        }
        __syncthreads();
        // overlap hand-off: region r+1 += second(r)   (x = overlap + imdct_first, commutative)
        if (active) {
            if (r + 1 < n_regions) {
                float* Xn = xt + (size_t)(r + 1) * 18 * kPitch * 2;
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int t = 0; t < 18; ++t) {
                        float* p = Xn + (t * kPitch + lane) * 2 + ch;
                        *p = *p + finv(sec[ch][t], lane, t);
                    }
            } else if (store_state) {
#pragma unroll
                for (int ch = 0; ch < 2; ++ch)
#pragma unroll
                    for (int t = 0; t < 18; ++t) st->overlap[ch][lane][t] = sec[ch][t];
            }
        }
        __syncthreads();
    }

    // A run-starting tile takes the overlap and the 15-slot polyphase history from the stream state.
    float* hist = xt + (size_t)(18 + 3) * kPitch * 2; // region 1, slot 3
    if (load_state) {
        float* X2 = xt + (size_t)2 * 18 * kPitch * 2;
        for (int idx = threadIdx.x; idx < 2 * 32 * 18; idx += NW * 32) {
            const int ch = idx / 576, rem = idx - ch * 576, sb = rem / 18, t = rem - sb * 18;
            float* p = X2 + (t * kPitch + sb) * 2 + ch;
            *p = *p + finv(st_in->overlap[ch][sb][t], sb, t);
        }
        for (int idx = threadIdx.x; idx < 15 * 32; idx += NW * 32) {
            const int s = idx >> 5, col = idx & 31;
            *reinterpret_cast<float2*>(hist + (s * kPitch + col) * 2) = st_in->dhist[s][col];
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------------------------------
    // Phase C: DCT-32 of every time slot, in place.  Half-warp = 16 slots of one channel, so the
    // 32-bit shared accesses of a warp hit 32 distinct banks (row pitch 66 words).
    // ------------------------------------------------------------------------------------------
    {
        const int s_first = load_state ? 36 : 18 + 3;
        const int s_last = n_regions * 18;
        const int ch = lane >> 4;
        for (int base = s_first + warp * 16; base < s_last; base += NW * 16) {
            const int s = base + (lane & 15);
            if (s < s_last) {
                float* row = xt + (size_t)s * kPitch * 2 + ch;
                float x[32], y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] = row[2 * i];
                lee_dct<32>(x, y);
#pragma unroll
                for (int i = 0; i < 32; ++i) row[2 * i] = y[i];
                row[64] = 0.0f; // column 32: V[16] = 0.0 (synthesis.rs:263)
            }
        }
        if (load_state) {
            for (int idx = threadIdx.x; idx < 30; idx += NW * 32) hist[((idx >> 1) * kPitch + 32) * 2 + (idx & 1)] = 0.0f;
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------------------------------
    // Phase D: polyphase window (synthesis.rs:247-263, :309-327).  lane = PCM sample index i; each
    // warp walks a contiguous range of slots with a 16-deep register window of (V_lo[i], V_hi[i]).
    //   V_lo[i] =  d[16+i] (i<16) | 0 (i=16) | -d[48-i] (i>16);   V_hi[i] = -d[16-i] (i<=16) | -d[i-16]
    // ------------------------------------------------------------------------------------------
    {
        const int col_lo = lane < 16 ? 16 + lane : (lane == 16 ? 32 : 48 - lane);
        const int col_hi = lane <= 16 ? 16 - lane : lane - 16;
        float dlo[8], dhi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a0 = __ldg(&tab->synth_d[64 * j + lane]);
            dlo[j] = lane > 16 ? -a0 : a0;
            dhi[j] = -__ldg(&tab->synth_d[64 * j + 32 + lane]);
        }
        const int total = n * 18;
        const int per = (total + NW - 1) / NW;
        const int s_begin = 36 + warp * per;
        const int s_end = min(36 + total, s_begin + per);
        if (s_begin < s_end) {
            float2 wl[16], wh[16];
#pragma unroll
            for (int m = 0; m < 15; ++m) { // slots s_begin-15 .. s_begin-1 -> index (m+1)&15
                const float* row = xt + (size_t)(s_begin - 15 + m) * kPitch * 2;
                wl[(m + 1) & 15] = *reinterpret_cast<const float2*>(row + 2 * col_lo);
                wh[(m + 1) & 15] = *reinterpret_cast<const float2*>(row + 2 * col_hi);
            }
            // Output pointer of slot s_begin; consecutive slots of a frame are contiguous in a PCM plane
            // (plane[gr*576 + t*32 + i]), frames are SYMGPU_MP3_FRAME_FLOATS apart.
            const int slots_per_frame = 18 * gpf;
            const int q0 = (gseq0 * 18) + (s_begin - 36); // slot sequence index within the batch
            int sif = q0 % slots_per_frame;                // slot in frame
            float* out = a.pcm + (size_t)(q0 / slots_per_frame) * SYMGPU_MP3_FRAME_FLOATS + sif * 32 + lane;
            const int frame_jump = SYMGPU_MP3_FRAME_FLOATS - slots_per_frame * 32;
            const bool stereo_out = n_ch == 2;
            for (int base = s_begin; base < s_end; base += 16) {
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int s = base + u;
                    if (s < s_end) {
                        const float* row = xt + (size_t)s * kPitch * 2;
                        wl[u] = *reinterpret_cast<const float2*>(row + 2 * col_lo);
                        wh[u] = *reinterpret_cast<const float2*>(row + 2 * col_hi);
                        float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 v0 = wl[(u - 2 * j) & 15];
                            const float2 v1 = wh[(u - 2 * j - 1) & 15];
                            o0 += v0.x * dlo[j];
                            o1 += v0.y * dlo[j];
                            o0 += v1.x * dhi[j];
                            o1 += v1.y * dhi[j];
                        }
                        out[0] = o0;
                        if (stereo_out) out[1152] = o1;
                        out += 32;
                        if (++sif == slots_per_frame) {
                            sif = 0;
                            out += frame_jump;
                        }
                    }
                }
            }
        }
    }

    // The run's last tile publishes the polyphase history (last 15 DCT vectors) for the next batch.
    if (store_state) {
        const float* last = xt + (size_t)(n_regions * 18 - 15) * kPitch * 2;
        for (int idx = threadIdx.x; idx < 15 * 32; idx += NW * 32) {
            const int s = idx >> 5, col = idx & 31;
            st->dhist[s][col] = *reinterpret_cast<const float2*>(last + (s * kPitch + col) * 2);
        }
    }

    // Launch epilogue: the last CTA to retire publishes the new state generation of every run.
    __shared__ bool is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        for (unsigned i = threadIdx.x; i < gridDim.x; i += NW * 32)
            if (a.tiles[i].flags & kTileStoreState) a.gen[a.tiles[i].stream] += 1;
        if (threadIdx.x == 0) *a.done = 0;
    }
}

// ---------------------------------------------------------------------------------------------
namespace {
template <int T, int NW>
cudaError_t launch(const Mp3Args& a, int n_tiles, cudaStream_t stream) {
    constexpr size_t smem = (size_t)(T + 2) * 18 * kPitch * 8 + NW * sizeof(WarpScratch);
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(mp3_synth_kernel<T, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    mp3_synth_kernel<T, NW><<<n_tiles, NW * 32, smem, stream>>>(a);
    return cudaGetLastError();
}
} // namespace

int mp3_tile_granules() { return kMp3TileGranules; }

cudaError_t mp3_launch(const Mp3Args& a, int n_tiles, cudaStream_t stream) {
    return launch<kMp3TileGranules, kMp3Warps>(a, n_tiles, stream);
}

} // namespace symgpu
