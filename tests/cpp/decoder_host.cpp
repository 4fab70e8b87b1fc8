// Host-side plug-in interface exercised from C++ (include/symgpu/decoder.hpp).
//   decoder_host registry            CPU tier: tier ordering / Unsupported / no-GPU error behaviour
//   decoder_host decode IN OUT       GPU tier: decode a stream of parsed MP3 packets one decode() call at a
//                                    time (BASELINE config 0 "plumbing") and write planar PCM
//   decoder_host threads S IN OUT    GPU tier: S decoder threads, one stream each, sharing ONE context: every decode() is a
//                                    symgpu_mp3_submit + symgpu_mp3_wait, the context batches the packets of all threads
//   decoder_host file LAYER IN OUT   GPU tier: an MPEG audio FILE end to end in C++ -- packetiser (packetizer.hpp), registry,
//                                    GpuMpaDecoder::decode on real frames with the packetiser's gapless trims -- planar PCM out
//   decoder_host file aac IN OUT     GPU tier: an ADTS file: frame index, registry, GpuAacDecoder::decode per raw_data_block
//   decoder_host file vorbis IN OUT  GPU tier: a Vorbis-in-Ogg file: pages -> packets -> mapping (durations, discards, end trims
//                                    against the page granule positions), registry, GpuVorbisDecoder::decode per audio packet
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <vector>

#include "../../include/symgpu/decoder.hpp"
#include "../../include/symgpu/packetizer.hpp"

using namespace symgpu_host;

struct StubDecoder final : AudioDecoder {
    AudioCodecParameters p;
    int tag;
    StubDecoder(AudioCodecParameters pp, int t) : p(std::move(pp)), tag(t) {}
    void reset() override {}
    const AudioCodecParameters& codec_params() const override { return p; }
    Result<AudioBufferRef> decode(const Packet&) override { return {{}, {ErrorKind::DecodeError, "stub"}}; }
    AudioBufferRef last_decoded() const override { return {}; }
};

static int check(bool cond, const char* what) {
    if (!cond) std::fprintf(stderr, "FAILED: %s\n", what);
    return cond ? 0 : 1;
}

static int test_registry() {
    int bad = 0;
    CodecRegistry reg;
    AudioCodecParameters mp3;
    mp3.codec = CODEC_ID_MP3;
    AudioCodecParameters aac;
    aac.codec = CODEC_ID_AAC;
    bad += check(reg.make_audio_decoder(mp3, {}).error.kind == ErrorKind::Unsupported, "empty registry -> Unsupported");
    auto stub = [](int tag) {
        return [tag](const AudioCodecParameters& p, const AudioDecoderOptions&) -> Result<std::unique_ptr<AudioDecoder>> {
            return {std::unique_ptr<AudioDecoder>(new StubDecoder(p, tag)), {}};
        };
    };
    reg.register_audio_decoder_at_tier(Tier::Fallback, CODEC_ID_MP3, stub(3));
    reg.register_audio_decoder_at_tier(Tier::Standard, CODEC_ID_MP3, stub(2));
    auto d = reg.make_audio_decoder(mp3, {});
    bad += check(d.ok() && static_cast<StubDecoder*>(d.value.get())->tag == 2, "standard beats fallback");
    reg.register_audio_decoder_at_tier(Tier::Preferred, CODEC_ID_MP3, stub(1));
    d = reg.make_audio_decoder(mp3, {});
    bad += check(d.ok() && static_cast<StubDecoder*>(d.value.get())->tag == 1, "preferred beats standard");
    bad += check(reg.make_audio_decoder(aac, {}).error.kind == ErrorKind::Unsupported, "unknown codec -> Unsupported");
    bad += check(map_status(SYMGPU_ERR_DECODE).kind == ErrorKind::DecodeError, "status mapping: decode");
    bad += check(map_status(SYMGPU_ERR_RESET).kind == ErrorKind::ResetRequired, "status mapping: reset");
    bad += check(map_status(SYMGPU_ERR_CUDA).kind == ErrorKind::IoError, "status mapping: cuda");
    // without a GPU the context cannot be created: the error is surfaced, nothing falls back to a CPU path
    auto gpu = GpuContext::create(0, 4);
    if (!gpu.ok()) {
        bad += check(gpu.error.kind == ErrorKind::IoError || gpu.error.kind == ErrorKind::Unsupported, "no GPU -> error");
        std::printf("no usable GPU: %s\n", gpu.error.message);
    } else {
        CodecRegistry r2;
        register_gpu_decoders(r2, gpu.value);
        auto g = r2.make_audio_decoder(mp3, {});
        bad += check(g.ok(), "GPU decoder registered at Tier::Preferred");
        uint8_t junk[16] = {0};
        Packet p;
        p.data = junk;
        p.len = sizeof junk;
        auto res = g.value->decode(p);
        bad += check(res.error.kind == ErrorKind::DecodeError && g.value->last_decoded().frames == 0,
                     "malformed packet -> DecodeError and an empty buffer");
    }
    std::printf(bad ? "registry: FAILED\n" : "registry: ok\n");
    return bad;
}

static int run_decode(const char* in_path, const char* out_path) {
    std::ifstream in(in_path, std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const size_t n = bytes.size() / GpuMpaDecoder::kPacketBytes;
    auto gpu = GpuContext::create(0, 2);
    if (!gpu.ok()) {
        std::fprintf(stderr, "%s\n", gpu.error.message);
        return 2;
    }
    CodecRegistry reg;
    register_gpu_decoders(reg, gpu.value);
    AudioCodecParameters params;
    params.codec = CODEC_ID_MP3;
    params.channels = 2;
    AudioDecoderOptions opts;
    opts.gapless = false;
    auto dec = reg.make_audio_decoder(params, opts);
    if (!dec.ok()) return 3;
    std::ofstream out(out_path, std::ios::binary);
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = 0; i < n; ++i) {
        Packet p;
        p.data = bytes.data() + i * GpuMpaDecoder::kPacketBytes;
        p.len = GpuMpaDecoder::kPacketBytes;
        auto res = dec.value->decode(p);
        if (!res.ok()) {
            std::fprintf(stderr, "decode failed at packet %zu: %s\n", i, res.error.message);
            return 4;
        }
        for (size_t ch = 0; ch < 2; ++ch)
            out.write(reinterpret_cast<const char*>(res.value.planes[ch]), (std::streamsize)(res.value.frames * sizeof(float)));
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("decoded %zu packets us_per_packet %.2f (AudioDecoder::decode, one packet per call, file write included)\n", n, 1e6 * sec / (double)n);
    return 0;
}


// S decoders on S threads share one context (the server shape of SURVEY 8b): IN = S streams x F parsed packets, stream-major.
static int run_threads(int n_streams, const char* in_path, const char* out_path) {
    std::ifstream in(in_path, std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const size_t n = bytes.size() / GpuMpaDecoder::kPacketBytes;
    if (n_streams <= 0 || n % (size_t)n_streams) return 6;
    const size_t per = n / (size_t)n_streams;
    auto gpu = GpuContext::create(0, (uint32_t)n_streams);
    if (!gpu.ok()) {
        std::fprintf(stderr, "%s\n", gpu.error.message);
        return 2;
    }
    CodecRegistry reg;
    register_gpu_decoders(reg, gpu.value);
    std::vector<float> pcm(n * 2 * 1152, 0.0f);
    std::atomic<int> failed{0};
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int s = 0; s < n_streams; ++s)
        threads.emplace_back([&, s] {
            AudioCodecParameters params;
            params.codec = CODEC_ID_MP3;
            params.channels = 2;
            AudioDecoderOptions opts;
            opts.gapless = false;
            auto dec = reg.make_audio_decoder(params, opts);
            if (!dec.ok()) {
                failed++;
                return;
            }
            for (size_t i = 0; i < per; ++i) {
                Packet p;
                p.data = bytes.data() + ((size_t)s * per + i) * GpuMpaDecoder::kPacketBytes;
                p.len = GpuMpaDecoder::kPacketBytes;
                auto res = dec.value->decode(p);
                if (!res.ok() || res.value.frames != 1152) {
                    failed++;
                    return;
                }
                for (size_t ch = 0; ch < 2; ++ch)
                    std::memcpy(pcm.data() + (((size_t)s * per + i) * 2 + ch) * 1152, res.value.planes[ch], 1152 * sizeof(float));
            }
        });
    for (auto& t : threads) t.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t batches = 0, frames = 0;
    symgpu_mp3_async_stats(gpu.value->raw(), &batches, &frames);
    std::ofstream out(out_path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(pcm.data()), (std::streamsize)(pcm.size() * sizeof(float)));
    std::printf("threads %d packets %zu batches %llu frames %llu frames_per_batch %.2f seconds %.4f packets_per_s %.0f failed %d\n", n_streams, n,
                (unsigned long long)batches, (unsigned long long)frames, batches ? (double)frames / (double)batches : 0.0, sec, (double)n / sec,
                failed.load());
    return failed.load() ? 4 : 0;
}

static int run_file(int layer, const char* in_path, const char* out_path) {
    std::ifstream in(in_path, std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    symgpu::packet::MpaTrack track;
    std::vector<symgpu::packet::MpaPacket> packets;
    if (symgpu::packet::MpaIndexer::index(bytes.data(), bytes.size(), track, packets) != symgpu::packet::Status::Ok) return 5;
    auto gpu = GpuContext::create(0, 2);
    if (!gpu.ok()) {
        std::fprintf(stderr, "%s\n", gpu.error.message);
        return 2;
    }
    CodecRegistry reg;
    register_gpu_decoders(reg, gpu.value);
    AudioCodecParameters params;
    params.codec = layer == 1 ? CODEC_ID_MP1 : layer == 2 ? CODEC_ID_MP2 : CODEC_ID_MP3;
    params.sample_rate = track.first.sample_rate;
    params.channels = (uint32_t)track.first.n_channels();
    auto dec = reg.make_audio_decoder(params, AudioDecoderOptions{});  // gapless: the tag's delay and padding are trimmed
    if (!dec.ok()) return 3;
    std::ofstream out(out_path, std::ios::binary);
    size_t good = 0, samples = 0;
    for (const auto& pk : packets) {
        Packet p;
        p.data = bytes.data() + pk.offset, p.len = pk.size, p.pts = (uint64_t)pk.pts, p.dur = pk.dur;
        p.trim_start = pk.trim_start, p.trim_end = (uint32_t)std::min<uint64_t>(pk.trim_end, pk.dur);
        auto res = dec.value->decode(p);
        if (!res.ok()) continue;  // a refused frame yields no audio, the stream goes on (what a player does)
        ++good, samples += res.value.frames;
        for (size_t ch = 0; ch < res.value.n_planes; ++ch)
            out.write(reinterpret_cast<const char*>(res.value.planes[ch]), (std::streamsize)(res.value.frames * sizeof(float)));
    }
    std::printf("decoded %zu of %zu packets, %zu samples per channel, delay %u padding %u\n", good, packets.size(), samples, track.delay, track.padding);
    return 0;
}

static int write_planes(std::ofstream& out, const AudioBufferRef& b) {
    for (size_t ch = 0; ch < b.n_planes; ++ch) out.write(reinterpret_cast<const char*>(b.planes[ch]), (std::streamsize)(b.frames * sizeof(float)));
    return 0;
}

static int run_adts(const char* in_path, const char* out_path) {
    std::ifstream in(in_path, std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    size_t count = 0;
    symgpu_status stop;
    if (symgpu_adts_index(bytes.data(), bytes.size(), nullptr, 0, &count, &stop) != SYMGPU_OK || count == 0) return 5;
    std::vector<symgpu_adts_packet> packets(count);
    symgpu_adts_index(bytes.data(), bytes.size(), packets.data(), count, &count, &stop);
    auto gpu = GpuContext::create(0, 2);
    if (!gpu.ok()) {
        std::fprintf(stderr, "%s\n", gpu.error.message);
        return 2;
    }
    CodecRegistry reg;
    register_gpu_decoders(reg, gpu.value);
    AudioCodecParameters params;
    params.codec = CODEC_ID_AAC, params.sample_rate = packets[0].sample_rate, params.channels = packets[0].channels;
    auto dec = reg.make_audio_decoder(params, AudioDecoderOptions{});
    if (!dec.ok()) return 3;
    std::ofstream out(out_path, std::ios::binary);
    size_t good = 0, samples = 0;
    for (const auto& pk : packets) {
        Packet p;
        p.data = bytes.data() + pk.offset, p.len = pk.size, p.pts = (uint64_t)pk.pts, p.dur = 1024;
        auto res = dec.value->decode(p);
        if (!res.ok()) continue;
        ++good, samples += res.value.frames;
        write_planes(out, res.value);
    }
    std::printf("decoded %zu of %zu packets, %zu samples per channel\n", good, packets.size(), samples);
    return 0;
}

static int run_ogg_vorbis(const char* in_path, const char* out_path) {
    using namespace symgpu::packet;
    std::ifstream in(in_path, std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    OggIndex ix;
    OggIndex::build(bytes.data(), bytes.size(), ix, false);
    if (ix.streams.empty()) return 5;
    auto& stream = ix.streams.begin()->second;
    OggVorbisMapper mapper;
    std::vector<std::vector<uint8_t>> audio;
    std::vector<uint32_t> seq, dur, discard;
    std::vector<uint64_t> absgp;
    bool first = true;
    for (const OggPacket& pk : stream.packets()) {
        std::vector<uint8_t> b(pk.len);
        stream.gather(bytes.data(), pk, b.data());
        if (first) {
            first = false;
            if (!mapper.detect(b.data(), b.size())) return 6;
            continue;
        }
        const auto m = mapper.map(b.data(), b.size());
        if (m.kind != OggVorbisMapper::Kind::Audio || !mapper.ready()) continue;
        seq.push_back(pk.page_sequence), absgp.push_back(pk.page_absgp), dur.push_back((uint32_t)m.dur), discard.push_back((uint32_t)m.discard);
        audio.push_back(std::move(b));
    }
    std::vector<uint32_t> trim_end(audio.size());
    symgpu_ogg_page_end_trims(seq.data(), absgp.data(), dur.data(), discard.data(), audio.size(), trim_end.data());
    auto gpu = GpuContext::create(0, 2);
    if (!gpu.ok()) {
        std::fprintf(stderr, "%s\n", gpu.error.message);
        return 2;
    }
    CodecRegistry reg;
    register_gpu_decoders(reg, gpu.value);
    AudioCodecParameters params;
    params.codec = CODEC_ID_VORBIS, params.sample_rate = mapper.ident().sample_rate, params.channels = mapper.ident().n_channels;
    params.extra_data = mapper.extra_data();
    auto dec = reg.make_audio_decoder(params, AudioDecoderOptions{});
    if (!dec.ok()) {
        std::fprintf(stderr, "%s\n", dec.error.message);
        return 3;
    }
    std::ofstream out(out_path, std::ios::binary);
    size_t good = 0, samples = 0;
    for (size_t k = 0; k < audio.size(); ++k) {
        Packet p;
        p.data = audio[k].data(), p.len = audio[k].size(), p.dur = dur[k], p.trim_start = discard[k], p.trim_end = trim_end[k];
        auto res = dec.value->decode(p);
        if (!res.ok()) continue;
        ++good, samples += res.value.frames;
        write_planes(out, res.value);
    }
    std::printf("decoded %zu of %zu packets, %zu samples per channel\n", good, audio.size(), samples);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "registry") return test_registry();
    if (argc >= 4 && std::string(argv[1]) == "decode") return run_decode(argv[2], argv[3]);
    if (argc >= 5 && std::string(argv[1]) == "threads") return run_threads(std::atoi(argv[2]), argv[3], argv[4]);
    if (argc >= 5 && std::string(argv[1]) == "file" && std::string(argv[2]) == "aac") return run_adts(argv[3], argv[4]);
    if (argc >= 5 && std::string(argv[1]) == "file" && std::string(argv[2]) == "vorbis") return run_ogg_vorbis(argv[3], argv[4]);
    if (argc >= 5 && std::string(argv[1]) == "file") return run_file(std::atoi(argv[2]), argv[3], argv[4]);
    std::fprintf(stderr, "usage: decoder_host registry | decode IN OUT | file LAYER|aac|vorbis IN OUT\n");
    return 64;
}
