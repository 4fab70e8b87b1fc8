// ORACLE (test infrastructure, NOT product code).  C surface of liboracle.so; see the headers of
// oracle_mp3.cpp / oracle_mdct.cpp / oracle_aac.cpp / oracle_vorbis.cpp for what each restates.
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
void oracle_mp3_dct32(const float* x, float* y);
void oracle_mp3_imdct36(float* x, const float* window, float* overlap);
void oracle_mp3_imdct12_win(float* x, const float* window, float* overlap);
void oracle_mp3_polyphase(oracle_mp3_state* st, int ch, int n_slots, const float* in, float* out);
const float* oracle_mp3_imdct_window(int which);
float oracle_mp3_pow43(int i);
size_t oracle_mp3_tables(float* out, size_t cap_floats);

#ifdef __cplusplus
}
#endif
