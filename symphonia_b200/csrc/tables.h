// Lookup tables of the synthesis engine.  Everything is built ON THE HOST with libm (f64 then
// cast, except where the reference itself computes in f32) and uploaded verbatim; device math
// never regenerates a table (SURVEY.md §7 "Tables come from host libm").  The blob is one POD
// struct so that multi-GPU ranks can broadcast it as raw bytes.
#pragma once
#include <cstddef>
#include <cstdint>

namespace symgpu {

// Exponent range of the requantisation scale 2^(0.25*(A-B)) (layer3/requantize.rs:259-280,
// :317-343): A = global_gain-210-8*subblock_gain in [-266, 45]; B is a u8 in [0, 255].
constexpr int kPow2qMin = -521;
constexpr int kPow2qLen = 568;

enum Mp3Kind { kKindLong = 0, kKindShort = 1, kKindMixed = 2 };

struct alignas(16) Mp3Tables {
    // ---- f32 tables (first 915 floats are compared 1:1 with the oracle's tables in tests) ----
    float synth_d[512];        // ISO 11172-3 Table B.3 as the reference's 9-decimal literals
    float imdct_win[4][36];    // long / start / short / end
    float half_cos12[6][6];
    float cs[8], ca[8];
    float is_mpeg1[7][2];
    float is_mpeg2[2][32][2];
    float dct_iv_scale[18];
    float sdct18_scale[9];
    float sdct9_d[7];
    float lee16[16], lee8[8], lee4[4], lee2[2], lee1;
    // ---- engine-only tables ----
    float pow2q[kPow2qLen];    // (float)pow(2.0, 0.25 * k), k = kPow2qMin ..
    float pow43[8208];         // f32 powf(i, 4/3), requantize.rs:23-32 (for CPU front-ends / workloads)
    // ---- integer maps ----
    alignas(16) uint16_t edges[9][3][41];       // [sample_rate_idx][kind][edge]; interval i = [edges[i], edges[i+1])
    uint8_t n_edges[9][3];
    uint8_t mixed_switch[9];
    uint8_t pre_emphasis[24];
    alignas(16) uint8_t iv_of_line[9][3][576];  // interval index of each spectral line (read as uchar4)
    alignas(16) uint16_t reorder_src[9][2][576]; // [..][0 short | 1 mixed][dest line] -> source line
    uint16_t reorder_start[9][2];
    // One word per destination line of a short / mixed granule, so that the reordered load needs a single
    // lookup: bits 0-9 source line, 10-15 interval of the source line, 16-21 interval of the line itself.
    alignas(16) uint32_t short_map[9][2][576];
};

// Builds the tables (host libm).  Thread-safe, built once.
const Mp3Tables& mp3_tables_host();

// ---- tables of the power-of-two IMDCT codecs (AAC-LC, Vorbis) --------------------------------
struct Cplx {
    float re, im;
};

struct alignas(16) CodecTables {
    // FFT (symphonia-core/src/dsp/fft/no_simd.rs): level-16 / level-32 literal twiddles and the
    // f64-built merge tables of sizes 64..2048, size s at offset s/2 - 32.
    Cplx fft_lit16[8];
    Cplx fft_lit32[16];
    Cplx fft_merge[2048 - 32];
    // Imdct::new_scaled twiddles (symphonia-core/src/dsp/mdct.rs:35-60).
    Cplx aac_tw_long[512];   // n = 1024, scale 1/2048   (aac/dsp.rs:49)
    Cplx aac_tw_short[64];   // n = 128,  scale 1/256    (aac/dsp.rs:50)
    Cplx vorbis_tw[4096 - 16]; // unscaled, n2 = 16..2048 complex at offset n2 - 16  (lib.rs:123-124)
    // AAC windows (aac/window.rs:28-63), first halves.
    float aac_sine_long[1024], aac_sine_short[128], aac_kbd_long[1024], aac_kbd_short[128];
    // Vorbis power-sine window left halves (window.rs:11-24): blocksize bs at offset bs/2 - 32.
    float vorbis_win[8192 - 32];
    // floor1_inverse_dB_table (floor.rs:21-86).
    float vorbis_inverse_db[256];
};

const CodecTables& codec_tables_host();

} // namespace symgpu
