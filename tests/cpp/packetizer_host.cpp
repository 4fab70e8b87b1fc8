// Test driver for include/symgpu/packetizer.hpp: prints what the index builders find in a file, one record per
// line, for tests/test_packetizer.py to compare with oracle/packetizer_oracle.py.
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/symgpu/packetizer.hpp"

using namespace symgpu::packet;

static std::vector<uint8_t> slurp(const char* path) {
    std::vector<uint8_t> v;
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        std::fprintf(stderr, "cannot open %s\n", path);
        std::exit(2);
    }
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n);
    std::fclose(f);
    return v;
}

static const char* name(Status s) {
    switch (s) {
        case Status::Ok: return "ok";
        case Status::EndOfStream: return "eof";
        case Status::DecodeError: return "decode";
        default: return "unsupported";
    }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string mode = argv[1];
    if (mode == "hdr") {  // every header word given on the command line
        for (int i = 2; i < argc; ++i) {
            const uint32_t w = uint32_t(std::strtoul(argv[i], nullptr, 16));
            MpaHeader h;
            const Status s = mpa_parse_header(w, h);
            if (s != Status::Ok) {
                std::printf("%08x %s synced=%d check=%d\n", w, name(s), int(mpa_is_synced(w)), int(mpa_check_header(w)));
                continue;
            }
            std::printf("%08x ok synced=%d check=%d version=%d layer=%d mode=%d rate=%u rate_idx=%d bitrate=%u ms=%d is=%d bound=%d emph=%d "
                        "copy=%d orig=%d pad=%d crc=%d size=%u ch=%d samples=%u side=%u hsize=%u\n",
                        w, int(mpa_is_synced(w)), int(mpa_check_header(w)), int(h.version), h.layer, int(h.mode), h.sample_rate, h.sample_rate_idx,
                        h.bitrate, h.mid_side, h.intensity, h.bound, h.emphasis, h.copyrighted, h.original, h.padding, h.crc, h.frame_size,
                        h.n_channels(), h.samples_per_frame(), h.side_info_len(), h.header_size());
        }
        return 0;
    }
    const std::vector<uint8_t> d = slurp(argv[2]);
    if (mode == "crc32") {
        const uint32_t init = argc > 3 ? uint32_t(std::strtoul(argv[3], nullptr, 16)) : 0;
        // whole buffer, and split at every point of a short prefix: the sliced loop and the byte loop must agree
        const uint32_t whole = crc32_update(init, d.data(), d.size());
        for (size_t cut = 0; cut <= d.size() && cut < 40; ++cut) {
            const uint32_t two = crc32_update(crc32_update(init, d.data(), cut), d.data() + cut, d.size() - cut);
            if (two != whole) return std::printf("split mismatch at %zu\n", cut), 1;
        }
        std::printf("%08x\n", whole);
        return 0;
    }
    if (mode == "crc16") return std::printf("%04x\n", crc16_ansi_le_update(0, d.data(), d.size())), 0;
    if (mode == "tag") {  // one frame: the heuristics and the parsed tags
        MpaHeader h;
        if (d.size() < 4 || mpa_parse_header(detail::be32(d.data()), h) != Status::Ok) return std::printf("bad header\n"), 1;
        MpaInfoTag it;
        MpaVbriTag vt;
        const bool mi = mpa_is_maybe_info_tag(d.data(), d.size(), h), mv = mpa_is_maybe_vbri_tag(d.data(), d.size(), h);
        const bool ri = mpa_read_info_tag(d.data(), d.size(), h, it), rv = mpa_read_vbri_tag(d.data(), d.size(), h, vt);
        std::printf("maybe_info=%d maybe_vbri=%d info=%d vbri=%d mdb=%d\n", mi, mv, ri, rv, mpa_main_data_begin(d.data(), d.size(), h));
        if (ri)
            std::printf("info frames=%d:%u bytes=%d:%u toc=%d quality=%d:%u cbr=%d lame=%d delay=%u padding=%u peak=%u\n", it.has_num_frames,
                        it.num_frames, it.has_num_bytes, it.num_bytes, it.has_toc, it.has_quality, it.quality, it.is_cbr, it.has_lame,
                        it.lame.delay, it.lame.padding, it.lame.peak);
        if (rv) std::printf("vbri bytes=%u frames=%u\n", vt.num_bytes, vt.num_mpeg_frames);
        return 0;
    }
    if (mode == "mpa" || mode == "mpa-noseek") {
        MpaTrack t;
        std::vector<MpaPacket> pk;
        const Status s = MpaIndexer::index(d.data(), d.size(), t, pk, mode == "mpa");
        if (s != Status::Ok) return std::printf("open %s\n", name(s)), 0;
        std::printf("track %08x delay=%d:%u:%u frames=%d:%" PRIu64 " tag=%d first=%" PRIu64 "\n", t.first_word, t.has_delay, t.delay, t.padding,
                    t.has_num_frames, t.num_frames, int(t.tag), t.first_packet_pos);
        for (const MpaPacket& p : pk) {
            MpaHeader h;
            mpa_parse_header(p.header, h);
            std::printf("p %" PRIu64 " %u %08x %" PRId64 " %u %u %" PRIu64 " %d\n", p.offset, p.size, p.header, p.pts, p.dur, p.trim_start, p.trim_end,
                        h.layer == 3 ? mpa_main_data_begin(d.data() + p.offset, p.size, h) : -1);
        }
        return 0;
    }
    if (mode == "adts") {
        std::vector<AdtsPacket> pk;
        bool truncated = false;
        const Status s = AdtsIndexer::index(d.data(), d.size(), pk, &truncated);
        for (const AdtsPacket& p : pk)
            std::printf("p %" PRIu64 " %u %" PRId64 " %u %d %d\n", p.offset, p.size, p.pts, p.sample_rate, p.channels, p.profile);
        std::printf("stop %s\n", truncated ? "truncated" : name(s));
        return 0;
    }
    if (mode == "ogg") {
        OggIndex ix;
        const Status s = OggIndex::build(d.data(), d.size(), ix);
        for (const OggPage& pg : ix.pages)
            std::printf("page %" PRIu64 " %u %u %" PRIu64 " %u %u\n", pg.offset, pg.serial, pg.sequence, pg.absgp, unsigned(pg.n_packets), pg.body_len);
        for (const auto& kv : ix.streams) {
            std::printf("stream %u\n", kv.first);
            const OggLogicalStream& ls = kv.second;
            for (const OggPacket& p : ls.packets()) {
                std::vector<uint8_t> bytes(p.len);
                ls.gather(d.data(), p, bytes.data());
                std::printf("k %u %" PRIu64 " %d %" PRIu64 " %08x", p.page_sequence, p.page_absgp, int(p.last_on_page), p.len,
                            crc32_update(0, bytes.data(), bytes.size()));
                for (uint32_t k = 0; k < p.n_pieces; ++k) std::printf(" %" PRIu64 ":%u", ls.pieces()[p.first_piece + k].offset, ls.pieces()[p.first_piece + k].len);
                std::printf("\n");
            }
        }
        std::printf("end %s rejected=%zu orphans=%zu\n", name(s), ix.rejected, ix.orphans);
        return 0;
    }
    if (mode == "vsetup") {  // extra_data layout: the 30-byte identification packet, then the setup packet
        VorbisIdent id;
        const Status si = vorbis_read_ident(d.data(), d.size(), id);
        if (si != Status::Ok) return std::printf("ident %s\n", name(si)), 0;
        std::printf("ident ok ch=%d rate=%u bs=%d,%d\n", id.n_channels, id.sample_rate, id.bs0_exp, id.bs1_exp);
        uint8_t modes = 0;
        uint64_t mask = 0;
        const Status ss = vorbis_read_setup_modes(d.data() + 30, d.size() - 30, id, modes, mask);
        if (ss != Status::Ok) return std::printf("setup %s\n", name(ss)), 0;
        std::printf("setup ok modes=%d mask=%" PRIx64 "\n", modes, mask);
        return 0;
    }
    if (mode == "xiph") {
        Piece a, b;
        const Status s = vorbis_unpack_xiph_laced(d.data(), d.size(), a, b);
        if (s != Status::Ok) return std::printf("%s\n", name(s)), 0;
        std::printf("ok %" PRIu64 ":%u %" PRIu64 ":%u\n", a.offset, a.len, b.offset, b.len);
        return 0;
    }
    if (mode == "oggvorbis") {  // pages -> packets -> mapper, per logical stream
        OggIndex ix;
        OggIndex::build(d.data(), d.size(), ix, false);
        for (const auto& kv : ix.streams) {
            const OggLogicalStream& ls = kv.second;
            OggVorbisMapper mp;
            bool detected = false;
            std::vector<uint8_t> bytes;
            static const char* kinds[] = {"audio", "comment", "setup", "unknown", "error"};
            for (size_t i = 0; i < ls.packets().size(); ++i) {
                const OggPacket& p = ls.packets()[i];
                bytes.resize(p.len);
                ls.gather(d.data(), p, bytes.data());
                if (i == 0) {
                    detected = mp.detect(bytes.data(), bytes.size());
                    std::printf("stream %u vorbis=%d\n", kv.first, int(detected));
                    if (!detected) break;
                    continue;
                }
                const OggVorbisMapper::Mapped m = mp.map(bytes.data(), bytes.size());
                std::printf("m %s %" PRIu64 " %" PRIu64 "\n", kinds[int(m.kind)], m.dur, m.discard);
            }
            if (detected)
                std::printf("extra %zu %08x ready=%d rap=%" PRIu64 "\n", mp.extra_data().size(),
                            crc32_update(0, mp.extra_data().data(), mp.extra_data().size()), int(mp.ready()), mp.max_rap_period());
        }
        return 0;
    }
    if (mode == "bench-mpa" || mode == "bench-adts" || mode == "bench-ogg") {  // host throughput of the index builders
        const int reps = argc > 3 ? std::atoi(argv[3]) : 5;
        size_t packets = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; ++r) {
            if (mode == "bench-mpa") {
                MpaTrack t;
                std::vector<MpaPacket> pk;
                MpaIndexer::index(d.data(), d.size(), t, pk);
                packets = pk.size();
            } else if (mode == "bench-adts") {
                std::vector<AdtsPacket> pk;
                AdtsIndexer::index(d.data(), d.size(), pk);
                packets = pk.size();
            } else {
                OggIndex ix;
                OggIndex::build(d.data(), d.size(), ix, false);
                packets = 0;
                for (const auto& kv : ix.streams) packets += kv.second.packets().size();
            }
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::printf("%s: %zu bytes, %zu packets, %.3f ms per pass, %.2f GB/s, %.2f M packets/s\n", mode.c_str(), d.size(), packets, s * 1e3,
                    double(d.size()) / s * 1e-9, double(packets) / s * 1e-6);
        return 0;
    }
    return 2;
}
