// Sanitizer fuzz driver for the host-side parsers (packetisers, MP3 / Layer I-II / FLAC / Vorbis / AAC front-ends, plan + jobs):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all ... tests/cpp/fuzz_frontends.cpp <csrc/*.cpp>
//   fuzz_frontends SEEDFILE... : every seed is mutated (bit flips, byte runs, truncation, splices) ITER times and pushed through
// every entry point; any out-of-bounds access, overflow or leak aborts.  Run by tests/test_fuzz_sanitized.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <vector>

#include "../../include/symgpu.h"
#include "../../include/symgpu/packetizer.hpp"

using namespace symgpu::packet;

static void run_all(const std::vector<uint8_t>& d) {
    const uint8_t* p = d.data();
    const size_t n = d.size();
    // ---- MPEG audio: index, serial front-end, plan + jobs, Layer I / II
    {
        symgpu_mpa_track track;
        size_t count = 0;
        if (symgpu_mpa_index(p, n, 1, &track, nullptr, 0, &count) == SYMGPU_OK && count) {
            std::vector<symgpu_mpa_packet> pk(count);
            symgpu_mpa_index(p, n, 1, &track, pk.data(), count, &count);
            std::vector<symgpu_mp3_gc> units(count * 4);
            std::vector<int16_t> quant(count * 4 * 576);
            std::vector<uint32_t> frame_of(count);
            size_t good = 0;
            symgpu_mp3_frame_info info;
            symgpu_mp3_fe* fe = nullptr;
            symgpu_mp3_fe_create(&fe);
            symgpu_mp3_fe_decode_packets(fe, p, n, pk.data(), count, units.data(), quant.data(), frame_of.data(), &good, &info);
            symgpu_mp3_fe_destroy(fe);
            uint32_t rounds = 0;
            symgpu_mp3_entropy_decode_cpu(p, n, pk.data(), count, units.data(), quant.data(), frame_of.data(), &good, &info, &rounds);
            std::vector<float> sub(count * 2 * 32 * 36);
            for (int layer = 1; layer <= 2; ++layer)
                symgpu_mpa12_fe_decode_packets(p, n, pk.data(), count, layer, sub.data(), frame_of.data(), &good, &info);
        }
        // a packet that is just the raw bytes
        symgpu_mp3_gc u[4];
        std::vector<int16_t> q(4 * 576);
        symgpu_mp3_fe* fe = nullptr;
        symgpu_mp3_fe_create(&fe);
        symgpu_mp3_frame_info info;
        symgpu_mp3_fe_decode(fe, p, n < 3000 ? n : 3000, u, q.data(), &info);
        symgpu_mp3_fe_destroy(fe);
    }
    // ---- ADTS, Ogg (+ Vorbis mapping), FLAC
    {
        size_t count = 0;
        symgpu_status stop;
        symgpu_adts_index(p, n, nullptr, 0, &count, &stop);
        {   // C entry points: index, then the gathered copy of everything
            size_t np_ = 0, nq = 0;
            symgpu_ogg_index(p, n, nullptr, 0, &np_, nullptr, 0, &nq);
            std::vector<symgpu_ogg_packet> pk(np_);
            std::vector<symgpu_piece> pc(nq), tab(np_);
            if (np_ && symgpu_ogg_index(p, n, pk.data(), np_, &np_, pc.data(), nq, &nq) == SYMGPU_OK) {
                size_t total = 0, used = 0;
                for (const auto& q : pk) total += q.len;
                std::vector<uint8_t> out(total + 1);
                symgpu_ogg_gather(p, n, pk.data(), np_, pc.data(), nq, out.data(), total, tab.data(), &used);
            }
        }
        OggIndex ix;
        OggIndex::build(p, n, ix, false);
        for (auto& kv : ix.streams) {
            OggVorbisMapper mp;
            std::vector<uint8_t> bytes;
            bool first = true;
            for (const OggPacket& pk : kv.second.packets()) {
                if (pk.len > (1u << 22)) continue;
                bytes.resize(pk.len);
                kv.second.gather(p, pk, bytes.data());
                if (first) {
                    first = false;
                    if (!mp.detect(bytes.data(), bytes.size())) break;
                } else {
                    mp.map(bytes.data(), bytes.size());
                }
            }
        }
        // the raw bytes as a Vorbis setup packet and as Xiph-laced extra data
        VorbisIdent id{2, 44100, 8, 11};
        uint8_t modes;
        uint64_t mask;
        vorbis_read_setup_modes(p, n, id, modes, mask);
        VorbisSetup full;
        vorbis_read_setup(p, n, id, full);
        symgpu_vorbis_ident cid{44100, 2, 8, 11, 0};
        symgpu_vorbis_setup_info sinfo;
        std::vector<symgpu_vorbis_floor1> fl(64);
        symgpu_vorbis_setup_parse(p, n, &cid, &sinfo, fl.data());
        Piece a, b;
        vorbis_unpack_xiph_laced(p, n, a, b);
        symgpu_flac_stream_info si;
        symgpu_flac_index(p, n, &si, nullptr, 0, &count);
        if (count) {
            std::vector<symgpu_flac_packet> fp(count);
            symgpu_flac_index(p, n, &si, fp.data(), count, &count);
            std::vector<symgpu_piece> tab(count);
            size_t total = 0;
            for (size_t i = 0; i < count; ++i) tab[i] = symgpu_piece{fp[i].offset, fp[i].size, 0}, total += size_t(fp[i].dur) * si.channels;
            std::vector<symgpu_flac_frame> fr(count);
            std::vector<symgpu_flac_frame_info> fi(count);
            std::vector<uint32_t> fo(count);
            std::vector<symgpu_flac_subframe> sf(count * 8);
            std::vector<int32_t> smp(total + 8);
            size_t g, ns, nm;
            symgpu_flac_fe_decode_packets(p, n, tab.data(), count, si.bits_per_sample, si.channels, si.block_max, fr.data(), fi.data(), fo.data(), sf.data(), sf.size(),
                                          smp.data(), smp.size(), &g, &ns, &nm);
        }
        // the raw bytes as one FLAC packet
        symgpu_piece one{0, uint32_t(n), 0};
        symgpu_flac_frame fr;
        symgpu_flac_frame_info fi;
        uint32_t fo;
        symgpu_flac_subframe sf[8];
        std::vector<int32_t> smp(8 * 65536);
        size_t g, ns, nm;
        symgpu_flac_fe_decode_packets(p, n, &one, 1, 16, 0, 0, &fr, &fi, &fo, sf, 8, smp.data(), smp.size(), &g, &ns, &nm);
    }
    {   // the raw bytes as an AudioSpecificConfig
        symgpu_aac_asc asc;
        symgpu_aac_asc_parse(p, n < 64 ? n : 64, &asc);
        symgpu_aac_fe* fe = nullptr;
        if (symgpu_aac_fe_create_asc(p, n < 64 ? n : 64, &fe, &asc) == SYMGPU_OK) symgpu_aac_fe_destroy(fe);
    }
    // ---- AAC entropy front-end: "AFE1", rate index byte, channel byte, then length-prefixed (u16 LE) raw_data_blocks;
    //      and every ADTS frame of the input as a packet
    if (n > 8 && std::memcmp(p, "AFE1", 4) == 0) {
        static const uint32_t rates[8] = {44100, 48000, 8000, 96000, 22050, 32000, 16000, 64000};
        symgpu_aac_fe* fe = nullptr;
        if (symgpu_aac_fe_create(rates[p[4] & 7], 1 + (p[5] & 1), &fe) == SYMGPU_OK) {
            symgpu_aac_unit units[2];
            std::vector<symgpu_aac_tns> tns(16);
            std::vector<float> co(2048);
            uint32_t nt = 0;
            size_t at = 6, k = 0;
            while (at + 2 <= n) {
                const size_t len = size_t(p[at]) | size_t(p[at + 1]) << 8;
                at += 2;
                const size_t take = len < n - at ? len : n - at;
                symgpu_aac_fe_decode(fe, p + at, take, 0, units, tns.data(), &nt, co.data());
                if (++k % 5 == 0) symgpu_aac_fe_reset(fe);
                at += take;
            }
            symgpu_aac_fe_destroy(fe);
        }
        {   // the same blocks as independent jobs
            std::vector<symgpu_piece> tab;
            size_t at = 6;
            while (at + 2 <= n && tab.size() < 64) {
                const size_t len = size_t(p[at]) | size_t(p[at + 1]) << 8;
                at += 2;
                const size_t take = len < n - at ? len : n - at;
                tab.push_back(symgpu_piece{at, uint32_t(take), 0});
                at += take;
            }
            std::vector<symgpu_aac_unit> u(2 * tab.size() + 2);
            std::vector<symgpu_aac_tns> t(16 * tab.size() + 16);
            std::vector<float> co(2048 * tab.size() + 2048);
            size_t nt = 0;
            symgpu_aac_fe_decode_packets_jobs(rates[p[4] & 7], 1 + (p[5] & 1), p, n, tab.data(), tab.size(), 0, u.data(), t.data(), t.size(), co.data(), &nt, 2);
        }
    }
    {
        size_t count = 0;
        symgpu_status stop;
        if (symgpu_adts_index(p, n, nullptr, 0, &count, &stop) == SYMGPU_OK && count) {
            std::vector<symgpu_adts_packet> pk(count);
            symgpu_adts_index(p, n, pk.data(), count, &count, &stop);
            symgpu_aac_fe* fe = nullptr;
            if (symgpu_aac_fe_create(pk[0].sample_rate, pk[0].channels == 1 ? 1 : 2, &fe) == SYMGPU_OK) {
                symgpu_aac_unit units[2];
                std::vector<symgpu_aac_tns> tns(16);
                std::vector<float> co(2048);
                uint32_t nt = 0;
                for (size_t i = 0; i < count; ++i)
                    if (pk[i].offset + pk[i].size <= n) symgpu_aac_fe_decode(fe, p + pk[i].offset, pk[i].size, 0, units, tns.data(), &nt, co.data());
                symgpu_aac_fe_destroy(fe);
            }
        }
    }
    // ---- Vorbis entropy front-end: "VFE1", then length-prefixed (u16 LE) identification, setup and audio packets
    if (n > 8 && std::memcmp(p, "VFE1", 4) == 0) {
        std::vector<Piece> parts;
        size_t at = 4;
        while (at + 2 <= n) {
            const size_t len = size_t(p[at]) | size_t(p[at + 1]) << 8;
            at += 2;
            const size_t take = len < n - at ? len : n - at;
            parts.push_back(Piece{at, uint32_t(take)});
            at += take;
        }
        symgpu_vorbis_fe* fe = nullptr;
        if (parts.size() >= 2 && symgpu_vorbis_fe_create(p + parts[0].offset, parts[0].len, p + parts[1].offset, parts[1].len, &fe) == SYMGPU_OK) {
            symgpu_vorbis_stream st;
            std::vector<symgpu_vorbis_floor1> fl(64);
            uint32_t nf = 0;
            symgpu_vorbis_fe_config(fe, &st, fl.data(), &nf);
            const uint32_t slot = (1u << st.bs1_exp) >> 1;
            std::vector<uint16_t> fy(2 * 65);
            std::vector<float> res(2 * size_t(slot));
            symgpu_vorbis_unit unit;
            for (size_t k = 2; k < parts.size(); ++k) {
                symgpu_vorbis_fe_decode(fe, p + parts[k].offset, parts[k].len, slot, 0, &unit, fy.data(), res.data());
                if (k % 5 == 0) symgpu_vorbis_fe_reset(fe);
            }
            symgpu_vorbis_fe_destroy(fe);
        }
        if (parts.size() >= 3 && parts.size() < 200) {  // the same packets as independent jobs
            std::vector<symgpu_piece> tab;
            for (size_t k = 2; k < parts.size(); ++k) tab.push_back(symgpu_piece{parts[k].offset, parts[k].len, 0});
            const uint32_t slot = 4096;
            std::vector<symgpu_vorbis_unit> u(tab.size());
            std::vector<uint16_t> fy(130 * tab.size());
            std::vector<float> res(2 * size_t(slot) * tab.size());
            std::vector<uint32_t> acc(tab.size());
            size_t good = 0;
            symgpu_vorbis_fe_decode_packets_jobs(p + parts[0].offset, parts[0].len, p + parts[1].offset, parts[1].len, p, n, tab.data(), tab.size(), slot, 0,
                                                 u.data(), fy.data(), res.data(), acc.data(), &good, 2);
        }
    }
}

int main(int argc, char** argv) {
    const int iters = std::getenv("FUZZ_ITERS") ? std::atoi(std::getenv("FUZZ_ITERS")) : 200;
    std::mt19937_64 rng(12345);
    size_t runs = 0;
    for (int a = 1; a < argc; ++a) {
        std::ifstream in(argv[a], std::ios::binary);
        const std::vector<uint8_t> seed((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        run_all(seed), ++runs;
        for (int it = 0; it < iters && !seed.empty(); ++it) {
            std::vector<uint8_t> d = seed;
            const int kinds = 1 + int(rng() % 3);
            for (int k = 0; k < kinds; ++k) {
                const size_t at = rng() % d.size();
                switch (rng() % 6) {
                    case 0: d[at] ^= uint8_t(1u << (rng() % 8)); break;
                    case 1: d[at] = uint8_t(rng()); break;
                    case 2: {
                        const size_t len = std::min<size_t>(1 + rng() % 16, d.size() - at);
                        std::memset(d.data() + at, (rng() & 1) ? 0xff : 0x00, len);
                        break;
                    }
                    case 3: d.resize(1 + at); break;
                    case 4: {
                        const size_t from = rng() % d.size(), len = std::min<size_t>(1 + rng() % 64, std::min(d.size() - from, d.size() - at));
                        std::memmove(d.data() + at, d.data() + from, len);
                        break;
                    }
                    default: {
                        const size_t len = 1 + rng() % 32;
                        std::vector<uint8_t> junk(len);
                        for (auto& x : junk) x = uint8_t(rng());
                        d.insert(d.begin() + at, junk.begin(), junk.end());
                    }
                }
                if (d.empty()) break;
            }
            if (!d.empty()) run_all(d), ++runs;
        }
    }
    std::printf("fuzz: %zu inputs, no sanitizer report\n", runs);
    return 0;
}
