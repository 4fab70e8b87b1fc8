/*
 * symgpu.h -- C ABI of libsymgpu.so, the B200 (sm_100a) batched audio-synthesis engine that
 * replaces the f32 DSP back-end of Symphonia's MP3 / AAC-LC / Vorbis decoders.
 *
 * The seam this ABI sits on is the point inside each reference decoder where the serial
 * bit-reader stage ends and the data-parallel synthesis stage begins:
 *
 *   MP3    symphonia-bundle-mp3/src/layer3/mod.rs:408 (read_main_data) | :421-477 (granule loop)
 *   AAC    symphonia-codec-aac/src/aac/mod.rs:217-220 -> ics/mod.rs:449-468 (Ics::synth_channel)
 *   Vorbis symphonia-codec-vorbis/src/lib.rs:248 (read_residue) | :250-315 (coupling/dot/synth)
 *
 * Everything left of the seam (frame sync, side info, Huffman / VQ decode) stays on the CPU in
 * the caller; everything right of it runs in one fused CUDA kernel per codec.  All entry points
 * take plain pointers and sizes.  "host" pointers are ordinary (ideally pinned) host memory and
 * the call performs the H2D / D2H copies itself; "dev" entry points take device pointers and
 * only enqueue the kernel on the context's CUDA stream (used when the spectra are already
 * resident in HBM).
 *
 * Error model (replaces symphonia_core::errors::Error, symphonia-core/src/errors.rs:43-57):
 * every function returns a symgpu_status; symgpu_strerror() returns a static string so a Rust
 * adapter can wrap it in Error::DecodeError(&'static str).  No C++ exception crosses the ABI.
 *
 * Threading (AudioDecoder: Send + Sync, symphonia-core/src/codecs/audio.rs:251): a context owns
 * one CUDA stream; calls on one context must be serialised by the caller (the trait's &mut self
 * already guarantees that per decoder).  Different contexts may be used concurrently.
 */
#ifndef SYMGPU_H
#define SYMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYMGPU_ABI_VERSION 1

typedef enum symgpu_status {
    SYMGPU_OK = 0,
    SYMGPU_ERR_DECODE = 1,       /* Error::DecodeError  - malformed unit, caller drops packet   */
    SYMGPU_ERR_UNSUPPORTED = 2,  /* Error::Unsupported                                          */
    SYMGPU_ERR_LIMIT = 3,        /* Error::LimitError   - batch / stream limits exceeded        */
    SYMGPU_ERR_RESET = 4,        /* Error::ResetRequired                                        */
    SYMGPU_ERR_CUDA = 5,         /* device failure (sticky; context must be destroyed)          */
    SYMGPU_ERR_ARG = 6           /* null pointer / bad size - a bug in the caller               */
} symgpu_status;

typedef struct symgpu_ctx symgpu_ctx; /* opaque */

/* ---- context ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` (cudaSetDevice ordinal), builds every lookup table
 * on the host with libm (see DESIGN.md "tables") and uploads them.  Fails with SYMGPU_ERR_CUDA
 * when no usable sm_100 device is present: there is NO CPU fallback behind this ABI. */
symgpu_status symgpu_ctx_create(int device, symgpu_ctx** out);
void symgpu_ctx_destroy(symgpu_ctx* ctx);

/* Static description of `status`. */
const char* symgpu_strerror(symgpu_status status);
/* Last CUDA error text seen by this context (static storage inside the context). */
const char* symgpu_last_cuda_error(const symgpu_ctx* ctx);
int symgpu_abi_version(void);

/* The host-built table blob (f32 words) exactly as uploaded, so that ranks can broadcast it
 * (ncclBroadcast / torch.distributed.broadcast) and tests can compare it with the oracle's
 * tables without a GPU.  Returns the number of bytes; `out` may be NULL to query the size. */
size_t symgpu_tables_host_blob(void* out, size_t cap);
/* Replace the device tables of `ctx` with a blob received from rank 0. */
symgpu_status symgpu_tables_upload(symgpu_ctx* ctx, const void* blob, size_t bytes);

/* Blocks until every kernel / copy enqueued on the context's stream has finished. */
symgpu_status symgpu_sync(symgpu_ctx* ctx);
/* The context's cudaStream_t, as an opaque pointer (for CUDA-event timing by the caller). */
void* symgpu_cuda_stream(symgpu_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t symgpu_launch_count(const symgpu_ctx* ctx);

/* ---- MPEG-1/2/2.5 Layer III ---------------------------------------------------------------- */

/* Block types, symphonia-bundle-mp3/src/layer3/common.rs:175-182. */
enum { SYMGPU_MP3_LONG = 0, SYMGPU_MP3_START = 1, SYMGPU_MP3_SHORT = 2, SYMGPU_MP3_END = 3 };

/* symgpu_mp3_gc.flags */
enum {
    SYMGPU_MP3_F_MIXED = 1 << 0,          /* BlockType::Short { is_mixed: true }                */
    SYMGPU_MP3_F_SCALEFAC_SCALE = 1 << 1, /* GranuleChannel::scalefac_scale                     */
    SYMGPU_MP3_F_PREFLAG = 1 << 2,        /* GranuleChannel::preflag                            */
    SYMGPU_MP3_F_SFC_LSB = 1 << 3,        /* scalefac_compress & 1 (MPEG-2 intensity scale)     */
    SYMGPU_MP3_F_MID_SIDE = 1 << 4,       /* frame: JointStereo(Layer3{mid_side})  header.rs:161 */
    SYMGPU_MP3_F_INTENSITY = 1 << 5,      /* frame: JointStereo(Layer3{intensity})              */
    SYMGPU_MP3_F_MPEG1 = 1 << 6,          /* frame: header.is_mpeg1()                           */
    SYMGPU_MP3_F_MUTE = 1 << 7            /* unit absent (mono ch1 / MPEG-2 gr1): skip, no PCM  */
};

/* One granule-channel: the fields of `GranuleChannel` (layer3/mod.rs:145-205) that the synthesis
 * stage reads, plus the three frame-header facts it needs.  64 bytes, 4 per frame laid out
 * [granule][channel].  `rzero` is the value returned by read_huffman_samples
 * (requantize.rs:236); the spectrum beyond it must be zero. */
typedef struct symgpu_mp3_gc {
    uint16_t rzero;            /* 0..576                                                        */
    uint8_t global_gain;       /*                                                               */
    uint8_t block_type;        /* SYMGPU_MP3_LONG..END                                          */
    uint8_t flags;             /* SYMGPU_MP3_F_*                                                */
    uint8_t sample_rate_idx;   /* 0..8, order of layer3/common.rs:9-55                          */
    uint8_t subblock_gain[3];  /*                                                               */
    uint8_t scalefacs[39];     /*                                                               */
    uint8_t reserved[16];      /* must be zero                                                  */
} symgpu_mp3_gc;

/* A run = `n_frames` consecutive frames of ONE stream, stored contiguously in the batch starting
 * at frame `first_frame`.  `stream` indexes the persistent per-stream synthesis state (hybrid
 * overlap + polyphase history; what `Layer3.overlap` / `Layer3.synthesis` hold,
 * layer3/mod.rs:254-259). */
typedef struct symgpu_mp3_run {
    uint32_t stream;
    uint32_t first_frame;
    uint32_t n_frames;
    uint8_t granules_per_frame; /* 2 = MPEG-1, 1 = MPEG-2 / 2.5 (header.n_granules()); 0 means 2 */
    uint8_t channels;           /* 1 or 2 (header.n_channels()); 0 means 2                      */
    uint16_t reserved;          /* must be zero                                                  */
} symgpu_mp3_run;

#define SYMGPU_MP3_LINES 576
#define SYMGPU_MP3_FRAME_FLOATS (2 * 2 * 576) /* spectra [gr][ch][576]; pcm [ch][gr*576 + i]  */

/* POW43[x] = f32 powf(x, 4/3) for x in 0..8206 (requantize.rs:23-32): the magnitude the CPU
 * Huffman stage writes for quantised value x (requantize.rs:128, :144).  Host libm, no GPU
 * needed.  Returns the table length (8207); copies min(cap, 8207) floats when `out` != NULL. */
size_t symgpu_mp3_pow43(float* out, size_t cap);

/* Allocates / zeroes device state for `n_streams` MP3 streams (Layer3::new, mod.rs:262-269). */
symgpu_status symgpu_mp3_streams_alloc(symgpu_ctx* ctx, uint32_t n_streams);
/* AudioDecoder::reset for one stream (decoder.rs:152-155): zero overlap + polyphase history. */
symgpu_status symgpu_mp3_stream_reset(symgpu_ctx* ctx, uint32_t stream);

/* Synthesises a batch of `n_frames` frames.
 *   units   [n_frames][2][2]      symgpu_mp3_gc
 *   spectra [n_frames][2][2][576] f32, values as left by read_huffman_samples
 *   runs    [n_runs]  (host)      frames of one stream are consecutive and in decode order; runs
 *                                 must tile [0, n_frames) without overlap and no stream may
 *                                 appear in two runs of the same call
 *   pcm     [n_frames][2][1152]   f32 planar per frame: plane(ch)[gr*576 .. gr*576+576]
 * Host variant: copies in, launches, copies out, and returns after the PCM is in `pcm`. */
symgpu_status symgpu_mp3_synth_host(symgpu_ctx* ctx, const symgpu_mp3_gc* units,
                                    const float* spectra, const symgpu_mp3_run* runs,
                                    uint32_t n_runs, uint32_t n_frames, float* pcm);
/* Device variant: `units`, `spectra` and `pcm` are device memory already resident in HBM; `runs`
 * is HOST memory (control plane: the library cuts runs into per-CTA tiles on the host).
 * Asynchronous on the context stream; call symgpu_sync() before reading `pcm`. */
symgpu_status symgpu_mp3_synth_dev(symgpu_ctx* ctx, const symgpu_mp3_gc* units,
                                   const float* spectra, const symgpu_mp3_run* runs,
                                   uint32_t n_runs, uint32_t n_frames, float* pcm);

#ifdef __cplusplus
}
#endif
#endif /* SYMGPU_H */
