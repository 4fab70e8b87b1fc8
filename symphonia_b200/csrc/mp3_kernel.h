// Host <-> kernel interface of the fused MP3 synthesis kernel (mp3_kernel.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/symgpu.h"
#include "tables.h"

namespace symgpu {

// Granules per tile and warps per CTA.  One persistent CTA of 16 warps per SM (4 per scheduler: 5
// would cap the kernel at 96 registers because registers are partitioned per scheduler): 13 + 2 halo
// granule jobs in the hybrid phase; (15 + 234) time slots x 2 channels = 498 DCT jobs for 512
// threads; 234 slots = 15 per warp in the window phase.  Shared memory: 252 XT rows (66.5 KB) + the
// TMA stage of 15 granules (71 KB) + scratch.
#ifndef SYMGPU_MP3_T
#define SYMGPU_MP3_T 13
#define SYMGPU_MP3_NW 16
#endif
constexpr int kMp3TileGranules = SYMGPU_MP3_T;
constexpr int kMp3Warps = SYMGPU_MP3_NW;

enum : uint8_t { kTileLoadState = 1, kTileStoreState = 2 };

// One CTA's work: `n_granules` consecutive granules of one stream.  Built on the host from the
// caller's runs (symgpu.cpp: build_tiles).
struct Mp3Tile {
    uint32_t first_frame; // batch frame index holding the tile's first granule
    uint32_t stream;      // per-stream state slot
    uint16_t first_gr;    // granule-in-frame of the tile's first granule
    uint16_t n_granules;  // 1..kMp3TileGranules
    uint8_t gpf;          // granules per frame: 2 (MPEG-1) or 1 (MPEG-2 / 2.5)
    uint8_t n_ch;         // 1 or 2
    uint8_t flags;        // kTileLoadState | kTileStoreState
    uint8_t pad;
};
static_assert(sizeof(Mp3Tile) == 16, "Mp3Tile is 16 bytes");

// Persistent per-stream state in HBM: what Layer3.overlap and Layer3.synthesis hold in the
// reference (layer3/mod.rs:254-259, synthesis.rs:145-154), in feed-forward form: instead of the
// 16x64 v_vec FIFO we keep the last 15 DCT-32 output vectors (both channels interleaved), from
// which every FIFO entry the next 15 slots can read is a copy or a negation.
struct Mp3StreamState {
    float overlap[2][32][18];
    float2 dhist[15][32];
};

struct Mp3Args {
    const symgpu_mp3_gc* units;
    const float* spectra;
    float* pcm;
    const Mp3Tile* tiles;
    int n_tiles;
    Mp3StreamState* states; // [n_streams][2] double-buffered, see gen
    uint32_t* gen;          // [n_streams] state generation; buffer (gen & 1) is current
    unsigned* done;         // retired-CTA counter (self-resetting)
    const Mp3Tables* tab;
};

cudaError_t mp3_upload_const(const Mp3Tables& t, cudaStream_t stream);
cudaError_t mp3_launch(const Mp3Args& a, cudaStream_t stream);
int mp3_tile_granules();

} // namespace symgpu
