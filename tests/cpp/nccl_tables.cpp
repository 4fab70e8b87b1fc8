// symgpu_tables_broadcast over a caller-made NCCL communicator (the one collective of the path), without Python:
//   nccl_tables IN OUT    IN = parsed MP3 packets of one stream (GpuMpaDecoder::kPacketBytes each); every rank decodes them
//                         after the broadcast and rank (n - 1)'s PCM is written to OUT.
// With >= 2 visible GPUs: two ranks (one thread each); rank 1's tables are ZEROED before the broadcast, so its PCM is only right
// if the tables of rank 0 really arrived (device blob and constant memory).  With one GPU: a 1-rank communicator, which still
// walks the whole path (dlopen, ncclBroadcast, constant refresh).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>
#include <vector>

#include "../../include/symgpu.h"

int main(int argc, char** argv) {
    if (argc < 3) return 64;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        std::fprintf(stderr, "libnccl.so.2 not found\n");
        return 3;
    }
    auto init_all = reinterpret_cast<int (*)(void**, int, const int*)>(dlsym(h, "ncclCommInitAll"));
    auto destroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    if (!init_all || !destroy) return 3;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) return 2;
    int n = n_dev >= 2 ? 2 : 1;
    int devs[2] = {0, 1};
    void* comms[2] = {nullptr, nullptr};
    int nccl_rc = init_all(comms, n, devs);
    if (nccl_rc != 0 && n == 2) {
        // the communicator is the host application's business, not the library's: when this box's NCCL cannot make a 2-rank one
        // (seen once on an 8-GPU NVSwitch box with the system libnccl), run the same code path over a 1-rank communicator
        std::fprintf(stderr, "ncclCommInitAll(2 ranks) failed with ncclResult_t %d: falling back to one rank\n", nccl_rc);
        n = 1;
        comms[0] = comms[1] = nullptr;
        nccl_rc = init_all(comms, n, devs);
    }
    if (nccl_rc != 0) {
        std::fprintf(stderr, "ncclCommInitAll failed with ncclResult_t %d\n", nccl_rc);
        return 4;
    }

    std::ifstream in(argv[1], std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const size_t pkt = 4 * sizeof(symgpu_mp3_gc) + SYMGPU_MP3_FRAME_FLOATS * sizeof(float);
    const uint32_t n_frames = (uint32_t)(bytes.size() / pkt);
    std::vector<symgpu_mp3_gc> units((size_t)n_frames * 4);
    std::vector<float> spectra((size_t)n_frames * SYMGPU_MP3_FRAME_FLOATS);
    for (uint32_t f = 0; f < n_frames; ++f) {
        std::memcpy(&units[(size_t)f * 4], bytes.data() + f * pkt, 4 * sizeof(symgpu_mp3_gc));
        std::memcpy(&spectra[(size_t)f * SYMGPU_MP3_FRAME_FLOATS], bytes.data() + f * pkt + 4 * sizeof(symgpu_mp3_gc), SYMGPU_MP3_FRAME_FLOATS * sizeof(float));
    }
    std::vector<std::vector<float>> pcm(n, std::vector<float>((size_t)n_frames * SYMGPU_MP3_FRAME_FLOATS, 0.0f));
    int status[2] = {0, 0};
    std::vector<std::thread> ranks;
    for (int r = 0; r < n; ++r)
        ranks.emplace_back([&, r] {
            symgpu_ctx* ctx = nullptr;
            if (symgpu_ctx_create(devs[r], &ctx) != SYMGPU_OK) {
                status[r] = 10;
                return;
            }
            if (r > 0) { // a rank whose own tables are useless: only the broadcast can make it decode
                std::vector<unsigned char> zeros(symgpu_tables_host_blob(nullptr, 0), 0);
                if (symgpu_tables_upload(ctx, zeros.data(), zeros.size()) != SYMGPU_OK) status[r] = 11;
            }
            const symgpu_status st = symgpu_tables_broadcast(ctx, comms[r], 0);
            if (st != SYMGPU_OK) {
                std::fprintf(stderr, "rank %d: symgpu_tables_broadcast: %s (%s)\n", r, symgpu_strerror(st), symgpu_last_cuda_error(ctx));
                status[r] = 12;
            }
            symgpu_mp3_run run{};
            run.stream = 0, run.first_frame = 0, run.n_frames = n_frames, run.granules_per_frame = 2, run.channels = 2;
            if (!status[r] && (symgpu_mp3_streams_alloc(ctx, 1) != SYMGPU_OK ||
                               symgpu_mp3_synth_host(ctx, units.data(), spectra.data(), &run, 1, n_frames, pcm[r].data()) != SYMGPU_OK))
                status[r] = 13;
            symgpu_ctx_destroy(ctx);
        });
    for (auto& t : ranks) t.join();
    for (int r = 0; r < n; ++r) destroy(comms[r]);
    for (int r = 0; r < n; ++r)
        if (status[r]) {
            std::fprintf(stderr, "rank %d failed with %d\n", r, status[r]);
            return status[r];
        }
    std::ofstream out(argv[2], std::ios::binary);
    out.write(reinterpret_cast<const char*>(pcm[n - 1].data()), (std::streamsize)(pcm[n - 1].size() * sizeof(float)));
    std::printf("ranks %d frames %u\n", n, n_frames);
    return 0;
}
