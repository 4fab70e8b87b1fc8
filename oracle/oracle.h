// ORACLE (test infrastructure, NOT product code).  C surface of liboracle.so; see the headers of
// oracle_mp3.cpp / oracle_mdct.cpp / oracle_aac.cpp / oracle_vorbis.cpp / oracle_conv.cpp / oracle_flac.cpp for what each restates.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../include/symgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

// Per-stream MP3 synthesis state: Layer3.overlap + Layer3.synthesis (layer3/mod.rs:254-259,
// synthesis.rs:145-154).
typedef struct oracle_mp3_state {
    float overlap[2][32][18];
    float v_vec[2][16][64];
    int32_t v_front[2];
} oracle_mp3_state;

void oracle_mp3_state_reset(oracle_mp3_state* st);
int oracle_mp3_frame(oracle_mp3_state* st, symgpu_mp3_gc* units, float* spectra, float* pcm, int n_gr, int n_ch);
int oracle_mp3_batch(oracle_mp3_state* states, const symgpu_mp3_gc* units, const float* spectra,
                     const symgpu_mp3_run* runs, uint32_t n_runs, float* pcm);
int oracle_mp3_batch_mt(oracle_mp3_state* states, const symgpu_mp3_gc* units, const float* spectra,
                        const symgpu_mp3_run* runs, uint32_t n_runs, float* pcm, int n_threads);
int oracle_mpa12_batch(oracle_mp3_state* states, const float* subbands, const symgpu_mpa12_run* runs, uint32_t n_runs,
                       uint32_t n_slots, float* pcm);
void oracle_mp3_dct32(const float* x, float* y);
void oracle_mp3_imdct36(float* x, const float* window, float* overlap);
void oracle_mp3_imdct12_win(float* x, const float* window, float* overlap);
void oracle_mp3_polyphase(oracle_mp3_state* st, int ch, int n_slots, const float* in, float* out);
const float* oracle_mp3_imdct_window(int which);
float oracle_mp3_pow43(int i);
size_t oracle_mp3_tables(float* out, size_t cap_floats);


// ---- shared IMDCT / FFT (oracle_mdct.cpp) ----
void oracle_fft_inplace(float* x, int n);
void oracle_imdct(const float* spec, float* out, int n, double scale);
void oracle_fft_twiddle(int size, int k, float* re_im);

// ---- AAC-LC filterbank (oracle_aac.cpp) ----
// Per-channel state: Ics.delay (aac/ics/mod.rs:207).
typedef struct oracle_aac_state {
    float delay[1024];
} oracle_aac_state;
void oracle_aac_tns(float* coeffs, const symgpu_aac_tns* filters, uint32_t n_filters);
void oracle_aac_synth(const float* coeffs, float* delay, int seq, int window_shape, int prev_window_shape, float* dst);
int oracle_aac_batch(oracle_aac_state* states, const symgpu_aac_unit* units, const symgpu_aac_tns* tns,
                     const float* coeffs, const symgpu_aac_run* runs, uint32_t n_runs, float* pcm, int n_threads);
const float* oracle_aac_window(int kbd, int is_short);

// ---- Vorbis synthesis (oracle_vorbis.cpp) ----
typedef struct oracle_vorbis_state {
    float overlap[2][4096]; // DspChannel.overlap, bs1/2 <= 4096
} oracle_vorbis_state;
typedef struct oracle_vorbis_mc_state {
    float overlap[SYMGPU_VORBIS_MAX_CHANNELS][4096];
} oracle_vorbis_mc_state;
// Same contract as symgpu_vorbis_mc_synth_host: every coupling step of the mapping in order (lib.rs:252-278), up to 8 channels.
int oracle_vorbis_mc_batch(oracle_vorbis_mc_state* states, const symgpu_vorbis_stream_mc* streams, const symgpu_vorbis_floor1* floors,
                           const symgpu_vorbis_unit_mc* units, const uint16_t* floor_y, const float* residue, const symgpu_vorbis_run* runs,
                           uint32_t n_runs, uint32_t channels, uint32_t slot, float* pcm);
void oracle_vorbis_floor1(const symgpu_vorbis_floor1* setup, const uint16_t* floor_y, uint32_t n, float* floor);
int oracle_vorbis_batch(oracle_vorbis_state* states, const symgpu_vorbis_stream* streams,
                        const symgpu_vorbis_floor1* floors, const symgpu_vorbis_unit* units, const uint16_t* floor_y,
                        const float* residue, const symgpu_vorbis_run* runs, uint32_t n_runs, uint32_t slot, float* pcm,
                        int n_threads);
const float* oracle_vorbis_window(int bs);
float oracle_vorbis_inverse_db(int i);

// ---- FLAC integer restoration (oracle_flac.cpp) ----
int oracle_flac_restore(const symgpu_flac_frame* frames, uint32_t n_frames, const symgpu_flac_subframe* subframes,
                        uint32_t n_subframes, int32_t* samples, size_t n_samples);

// ---- output stage (oracle_conv.cpp) ----
int16_t oracle_conv_s16(float s);
int32_t oracle_conv_s24(float s);
int32_t oracle_conv_s32(float s);
uint8_t oracle_conv_u8(float s);
int oracle_pcm_pack(const float* pcm, const symgpu_pcm_span* spans, uint32_t n_spans, uint32_t channels,
                    uint32_t plane_stride, uint32_t frames, int format, void* out);

#ifdef __cplusplus
}

#include <vector>
namespace oracle {
struct Imdct {
    int n;
    std::vector<float> tw_re, tw_im;
    Imdct(int n, double scale);
    void run(const float* spec, float* out) const; // spec[n] -> out[2n]
};
const Imdct& imdct_for(int n, double scale);
} // namespace oracle
#endif
