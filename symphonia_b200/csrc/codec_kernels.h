// Host <-> kernel interface of the AAC and Vorbis synthesis kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/symgpu.h"
#include "tables.h"

namespace symgpu {

enum : uint8_t { kChunkLoadState = 1, kChunkStoreState = 2 };

// One CTA's work: `count` consecutive frames / packets of one stream (AAC: of one channel).
struct CodecChunk {
    uint32_t first;   // batch index of the first frame / packet
    uint32_t stream;  // state slot
    uint16_t count;
    uint8_t channel;  // AAC only
    uint8_t flags;    // kChunkLoadState: starts its run (take state from HBM) | kChunkStoreState: ends it
    uint32_t pad;
};
static_assert(sizeof(CodecChunk) == 16, "CodecChunk is 16 bytes");

constexpr int kAacChunkFrames = 6;      // two warps per frame
constexpr int kAacChunkFramesWarp = 13; // one warp per frame
constexpr int kAacChunkFramesZ = 15;    // one warp per frame, Z layout (16 warps, two CTAs per SM)
constexpr int kAacDefaultVariant = 2;    // 0 pair | 1 warp | 2 z (SYMGPU_AAC_KERNEL overrides)
int aac_kernel_variant();
int aac_launch_count(bool any_tns);        // kernels one aac_launch starts
bool aac_warp_per_frame();
int aac_chunk_frames();                 // frames per chunk of the variant in use (long runs)
int aac_chunk_frames_for(uint32_t mean_run_frames); // ... for a batch whose runs hold that many frames on average

constexpr int kVorbisStateFloats = 2 * 4096; // overlap of both channels, blocksize_1 <= 8192

struct AacArgs {
    const symgpu_aac_unit* units;
    const symgpu_aac_tns* tns;
    const float* coeffs;
    const float* tns_scratch;   // same buffer as tns_scratch_rw, read side
    float* tns_scratch_rw;
    uint32_t* tns_sorted;       // [n_tns] filter indices ordered by filter order
    uint32_t* tns_owner;        // [n_tns] channel-frame of each filter
    uint32_t n_tns;
    int n_chunks;               // filled in by aac_launch
    int z_frames;               // frames per chunk the plan was made for (Z kernel: 15 | 9 | 7)
    int tns_inline;             // filled in by aac_launch: the Z kernel applies the filters itself (no pre-pass)
    int n_groups;               // filled in by aac_launch: group_first = (const uint32_t*)(chunks + n_chunks), n_groups + 1 entries
    float* pcm;
    const CodecChunk* chunks;
    float* states;              // [n_streams][2 generations][2 channels][1024]
    uint32_t* gen;
    unsigned* done;
    const CodecTables* tab;
};

// Per floor-1 setup, computed on the host when the setup is registered: post i (>= 2) depends on its two
// neighbours among the earlier posts; level[i] = 1 + max(level[low[i]], level[high[i]]), level[0] = level[1] = 0.
struct FloorAux {
    uint8_t level[65];
    uint8_t max_level;
    uint8_t pad[6];
};
static_assert(sizeof(FloorAux) == 72, "FloorAux is 72 bytes");

struct VorbisArgs {
    const symgpu_vorbis_unit* units;
    const uint16_t* floor_y;
    const float* residue;
    float* pcm;
    const CodecChunk* chunks;
    const symgpu_vorbis_stream* streams;
    const symgpu_vorbis_floor1* floors;
    const FloorAux* floor_aux;
    uint32_t n_floors;
    uint32_t slot;              // floats per channel slot in residue / pcm
    uint32_t pkt_ch;            // channel planes per packet in floor_y / residue / pcm: 2, or C for the multichannel entry points
    uint32_t ch_base;           // first of the (at most two) planes this launch works on
    float* states;              // [n_streams][2 generations][kVorbisStateFloats]
    uint32_t* gen;
    unsigned* done;
    const CodecTables* tab;
};

cudaError_t aac_launch(const AacArgs& a, uint32_t n_units, bool any_tns, int n_chunks, int n_groups, cudaStream_t stream);
cudaError_t vorbis_launch(const VorbisArgs& a, int n_chunks, int max_bs1_exp, int n_slots, cudaStream_t stream);
// Multichannel helpers (symgpu_vorbis_mc_*): inverse coupling over every step of a mapping (lib.rs:252-278), in place on
// residue [n_packets][channels][slot]; and the (block flags, floor, do-not-decode) records of channel pair `pair`.
cudaError_t vorbis_mc_decouple_launch(const symgpu_vorbis_unit_mc* units, const uint32_t* stream_of_packet, const symgpu_vorbis_stream_mc* streams,
                                      float* residue, uint32_t n_packets, uint32_t channels, uint32_t slot, cudaStream_t stream);
cudaError_t vorbis_mc_split_units_launch(const symgpu_vorbis_unit_mc* units, uint32_t n_packets, uint32_t pair, symgpu_vorbis_unit* out,
                                         cudaStream_t stream);
// Packet slots per CTA (chunk packets + 1) for a batch whose largest blocksize_1 is 2^max_bs1_exp.
int vorbis_slots_for(int max_bs1_exp);
bool vorbis_kernel_z();

} // namespace symgpu
