// MP3 entropy front-end (include/symgpu.h "MP3 entropy front-end"): the serial, per-stream half of
// the Layer III decoder -- side information, bit reservoir, scale factors, Huffman-coded spectrum -- producing the
// batch format the synthesis kernels consume.  CPU only.
//
// Same results as the reference's reader (symphonia-bundle-mp3/src/layer3/{mod,bitstream,requantize}.rs), different
// construction: the Huffman tables are two-level direct-lookup tables built once from the standard's (code, length)
// lists (mp3_huffman_data.inc), the bit reader is a 64-bit big-endian window that reads zeros past the end of the
// data and reports over-reads by position, and the spectrum is written as int16 sign * x -- the POW43 lookup that the
// reference folds into this loop (requantize.rs:128, :144) runs on the device.
#include <algorithm>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "../../include/symgpu.h"
#include "../../include/symgpu/packetizer.hpp"
#include "mp3_entropy.h"
#include "tables.h"

namespace {

#include "mp3_huffman_data.inc"

using symgpu::packet::MpaHeader;
using symgpu::packet::MpaVersion;

// ---------------------------------------------------------------------------------------------- Huffman tables
// Two-level direct lookup: the first kFirstBits bits of the window select either a finished entry or a second-level
// table sized for the longest code under that prefix (entry format: mp3_entropy.h).
constexpr unsigned kFirstBits = 9;

void append_table(std::vector<uint32_t>& lut, uint32_t& base, uint8_t& first_bits, const uint32_t* packed, size_t n, unsigned wrap, bool quad) {
    unsigned max_len = 0;
    for (size_t i = 0; i < n; ++i) max_len = std::max(max_len, packed[i] >> 24);
    const unsigned first = std::min(max_len, kFirstBits);
    base = uint32_t(lut.size()), first_bits = uint8_t(first);
    std::vector<uint32_t> t(size_t(1) << first, 0);
    std::vector<unsigned> deepest(t.size(), 0);  // longest code under each first-level prefix
    for (size_t i = 0; i < n; ++i) {
        const unsigned len = packed[i] >> 24, code = packed[i] & 0x7ffff;
        if (len > first) deepest[code >> (len - first)] = std::max(deepest[code >> (len - first)], len - first);
    }
    for (size_t pfx = 0; pfx < deepest.size(); ++pfx)
        if (deepest[pfx]) {
            t[pfx] = 0x80000000u | uint32_t(t.size()) | (deepest[pfx] << 24);
            t.resize(t.size() + (size_t(1) << deepest[pfx]), 0);
        }
    for (size_t i = 0; i < n; ++i) {
        const unsigned len = packed[i] >> 24, code = packed[i] & 0x7ffff;
        const uint32_t entry = (quad ? unsigned(i) : unsigned(((i / wrap) << 4) | (i % wrap))) | (len << 8);
        if (len <= first) {
            const unsigned pad = first - len;
            for (unsigned k = 0; k < (1u << pad); ++k) t[(code << pad) + k] = entry;
        } else {
            const unsigned rest = len - first, prefix = code >> rest;
            const unsigned sub = (t[prefix] >> 24) & 31, at = t[prefix] & 0xffffff, pad = sub - rest;
            for (unsigned k = 0; k < (1u << pad); ++k) t[at + ((code & ((1u << rest) - 1)) << pad) + k] = entry;
        }
    }
    lut.insert(lut.end(), t.begin(), t.end());
}

struct HostTables {
    std::vector<uint32_t> lut;
    symgpu::mp3e::HuffSet set{};
    HostTables() {
        static const uint8_t linbits[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13};
        std::memcpy(set.linbits, linbits, 32);
#define SYMGPU_BIG(T, W) append_table(lut, set.base[T], set.first_bits[T], kHuff_##T, sizeof(kHuff_##T) / 4, W, false)
        SYMGPU_BIG(1, 2), SYMGPU_BIG(2, 3), SYMGPU_BIG(3, 3), SYMGPU_BIG(5, 4), SYMGPU_BIG(6, 4), SYMGPU_BIG(7, 6), SYMGPU_BIG(8, 6), SYMGPU_BIG(9, 6);
        SYMGPU_BIG(10, 8), SYMGPU_BIG(11, 8), SYMGPU_BIG(12, 8), SYMGPU_BIG(13, 16), SYMGPU_BIG(15, 16), SYMGPU_BIG(16, 16), SYMGPU_BIG(24, 16);
#undef SYMGPU_BIG
        for (int t = 17; t < 24; ++t) set.base[t] = set.base[16], set.first_bits[t] = set.first_bits[16];  // same codes, other linbits
        for (int t = 25; t < 32; ++t) set.base[t] = set.base[24], set.first_bits[t] = set.first_bits[24];
        append_table(lut, set.base[32], set.first_bits[32], kHuff_quadA, 16, 16, true);
        append_table(lut, set.base[33], set.first_bits[33], kHuff_quadB, 16, 16, true);
        set.lut = lut.data();
    }
};
const HostTables& host_tables() {
    static const HostTables t;
    return t;
}

// ---------------------------------------------------------------------------------------------- frame data
using symgpu::mp3e::Bits;
using symgpu::mp3e::GcSide;

struct Frame {
    unsigned main_data_begin;
    unsigned scfsi[2];  // bit g: group g of granule 1 repeats granule 0's scale factors
    GcSide gc[2][2];
    uint8_t scalefacs[2][2][39];
    uint16_t rzero[2][2];
};

// bitstream.rs:57-236
bool read_side_info(Bits& bs, const MpaHeader& h, Frame& f) {
    const bool mpeg1 = h.version == MpaVersion::Mpeg1;
    const int n_ch = h.n_channels(), n_gr = h.n_granules();
    const uint16_t* long_edges = symgpu::mp3_tables_host().edges[h.sample_rate_idx][symgpu::kKindLong];
    uint32_t v;
    if (mpeg1) {
        if (!bs.read(9, v)) return false;
        f.main_data_begin = v;
        if (!bs.skip(n_ch == 1 ? 5 : 3)) return false;
        for (int ch = 0; ch < n_ch; ++ch) {
            if (!bs.read(4, v)) return false;
            f.scfsi[ch] = (v >> 3 & 1) | (v >> 1 & 2) | (v << 1 & 4) | (v << 3 & 8);  // first bit read = group 0
        }
    } else {
        if (!bs.read(8, v)) return false;
        f.main_data_begin = v;
        if (!bs.skip(n_ch == 1 ? 1 : 2)) return false;
    }
    for (int gr = 0; gr < n_gr; ++gr)
        for (int ch = 0; ch < n_ch; ++ch) {
            GcSide& c = f.gc[gr][ch];
            if (!bs.read(12, v)) return false;
            c.part2_3_length = uint16_t(v);
            if (!bs.read(9, v)) return false;
            c.big_values = uint16_t(v);
            if (c.big_values > 288) return false;
            if (!bs.read(8, v)) return false;
            c.global_gain = uint8_t(v);
            if (!bs.read(mpeg1 ? 4 : 9, v)) return false;
            c.scalefac_compress = uint16_t(v);
            if (!bs.read(1, v)) return false;
            if (v) {  // window switching
                uint32_t type, mixed;
                if (!bs.read(2, type) || !bs.read(1, mixed)) return false;
                if (type == 0) return false;
                c.block_type = uint8_t(type == 1 ? SYMGPU_MP3_START : type == 2 ? SYMGPU_MP3_SHORT : SYMGPU_MP3_END);
                c.mixed = type == 2 && mixed;
                for (int i = 0; i < 2; ++i) {
                    if (!bs.read(5, v)) return false;
                    c.table_select[i] = uint8_t(v);
                }
                for (int i = 0; i < 3; ++i) {
                    if (!bs.read(3, v)) return false;
                    c.subblock_gain[i] = uint8_t(v);
                }
                // region0 ends after 36 lines (MPEG-1, and short blocks of MPEG-2), 54 (MPEG-2 long transitions), or,
                // for MPEG-2.5, after 6 (pure short) / 8 long bands of the rate's table (bitstream.rs:108-150)
                if (h.version == MpaVersion::Mpeg2p5) c.region1_start = long_edges[type == 2 && !mixed ? 6 : 8];
                else c.region1_start = (mpeg1 || type == 2) ? 36 : 54;
                c.region2_start = 576;
            } else {
                c.block_type = SYMGPU_MP3_LONG;
                for (int i = 0; i < 3; ++i) {
                    if (!bs.read(5, v)) return false;
                    c.table_select[i] = uint8_t(v);
                }
                uint32_t r0, r1;
                if (!bs.read(4, r0) || !bs.read(3, r1)) return false;
                const unsigned a = r0 + 1, b = r1 + a + 1;
                c.region1_start = long_edges[a];
                c.region2_start = b <= 22 ? long_edges[b] : 576;
            }
            if (mpeg1) {
                if (!bs.read(1, v)) return false;
                c.preflag = uint8_t(v);
            }
            if (!bs.read(1, v)) return false;
            c.scalefac_scale = uint8_t(v);
            if (!bs.read(1, v)) return false;
            c.count1table = uint8_t(v);
        }
    return true;
}

// decoder.rs:84-131: synchronise inside the packet, parse, insist that the packet is exactly one frame of the stream's
// signal specification (fixed by the first packet ever shown, good or bad) and of Layer III; skip the CRC.
struct Spec {
    bool have = false;
    uint32_t rate = 0;
    int channels = 0;
};
symgpu_status open_packet(Spec& spec, const uint8_t* frame, size_t n, MpaHeader& h, const uint8_t*& buf, size_t& buf_len) {
    using namespace symgpu::packet;
    size_t q = 0;
    uint32_t word = 0;
    for (;; ++q) {
        if (q + 4 > n) return SYMGPU_ERR_DECODE;
        word = detail::be32(frame + q);
        if (mpa_is_synced(word) && mpa_check_header(word)) break;
    }
    const Status hs = mpa_parse_header(word, h);
    if (hs != Status::Ok) return hs == Status::Unsupported ? SYMGPU_ERR_UNSUPPORTED : SYMGPU_ERR_DECODE;
    const size_t body_len = n - q - 4;
    if (h.frame_size != body_len) return SYMGPU_ERR_DECODE;
    if (!spec.have) spec.have = true, spec.rate = h.sample_rate, spec.channels = h.n_channels();
    else if (spec.rate != h.sample_rate || spec.channels != h.n_channels()) return SYMGPU_ERR_DECODE;
    if (h.layer != 3) return SYMGPU_ERR_DECODE;
    const size_t crc_len = h.crc ? 2 : 0;
    if (body_len < crc_len) return SYMGPU_ERR_DECODE;
    buf = frame + q + 4 + crc_len, buf_len = body_len - crc_len;
    return SYMGPU_OK;
}

}  // namespace

struct symgpu_mp3_fe {
    uint8_t reservoir[2048];
    size_t len = 0, consumed = 0;
    Spec spec;
    void clear() { len = consumed = 0; }
};

const symgpu::mp3e::HuffSet& symgpu::mp3_huffset_host(size_t* words) {
    if (words) *words = host_tables().lut.size();
    return host_tables().set;
}

extern "C" symgpu_status symgpu_mp3_fe_create(symgpu_mp3_fe** out) {
    if (!out) return SYMGPU_ERR_ARG;
    host_tables();
    *out = new (std::nothrow) symgpu_mp3_fe();
    return *out ? SYMGPU_OK : SYMGPU_ERR_LIMIT;
}
extern "C" void symgpu_mp3_fe_destroy(symgpu_mp3_fe* fe) { delete fe; }
extern "C" void symgpu_mp3_fe_reset(symgpu_mp3_fe* fe) {
    if (fe) fe->clear(), fe->spec = Spec{};
}

extern "C" symgpu_status symgpu_mp3_fe_decode(symgpu_mp3_fe* fe, const uint8_t* frame, size_t n, symgpu_mp3_gc* units, int16_t* quant,
                                              symgpu_mp3_frame_info* info) {
    using namespace symgpu::packet;
    if (!fe || (!frame && n) || !units || !quant) return SYMGPU_ERR_ARG;
    MpaHeader h{};
    const uint8_t* buf = nullptr;
    size_t buf_len = 0;
    {
        const symgpu_status hs = open_packet(fe->spec, frame, n, h, buf, buf_len);
        if (hs != SYMGPU_OK) return hs;
    }
    Frame f{};
    Bits side(buf, buf_len);
    if (!read_side_info(side, h, f)) return fe->clear(), SYMGPU_ERR_DECODE;
    const size_t side_len = h.side_info_len();
    if (side_len > buf_len) return fe->clear(), SYMGPU_ERR_DECODE;  // (the reference would panic slicing; cannot happen for a sized frame)

    // BitResevoir::fill (mod.rs:42-95)
    const uint8_t* md = buf + side_len;
    const size_t md_len = buf_len - side_len, begin = f.main_data_begin;
    if (begin + md_len > sizeof fe->reservoir) return SYMGPU_ERR_DECODE;  // returned before anything changes, reservoir kept
    const size_t unread = fe->len - fe->consumed;
    uint32_t underflow = 0;
    if (begin <= unread) {
        std::memmove(fe->reservoir, fe->reservoir + fe->len - begin, begin);
        std::memcpy(fe->reservoir + begin, md, md_len);
        fe->len = begin + md_len;
    } else {
        std::memmove(fe->reservoir, fe->reservoir + fe->len - unread, unread);
        std::memcpy(fe->reservoir + unread, md, md_len);
        fe->len = unread + md_len;
        underflow = uint32_t(begin - unread);
    }
    fe->consumed = 0;

    // read_main_data (mod.rs:272-370)
    const int n_ch = h.n_channels(), n_gr = h.n_granules();
    const bool mpeg1 = h.version == MpaVersion::Mpeg1;
    const bool intensity = h.mode == MpaMode::JointStereo && h.intensity;
    const uint32_t underflow_bits = 8 * underflow;
    size_t part_begin = 0;
    uint32_t skipped = 0;
    std::memset(quant, 0, 4 * 576 * sizeof(int16_t));
    for (int gr = 0; gr < n_gr; ++gr) {
        if (skipped < underflow_bits) {  // the granule's bits are (partly) in frames never seen: silence it
            for (int ch = 0; ch < n_ch; ++ch) skipped += f.gc[gr][ch].part2_3_length;
            if (skipped > underflow_bits) part_begin = skipped - underflow_bits;
            continue;
        }
        for (int ch = 0; ch < n_ch; ++ch) {
            if ((part_begin >> 3) > fe->len) return fe->clear(), SYMGPU_ERR_DECODE;
            Bits bs(fe->reservoir, fe->len, part_begin);
            if (bs.at > bs.n_bits) return fe->clear(), SYMGPU_ERR_DECODE;
            GcSide& c = f.gc[gr][ch];
            const int part2 = mpeg1 ? symgpu::mp3e::read_scale_factors_mpeg1(bs, c, gr ? f.scalefacs[0][ch] : nullptr, f.scfsi[ch], f.scalefacs[gr][ch])
                                    : symgpu::mp3e::read_scale_factors_mpeg2(bs, ch > 0 && intensity, c, &c.preflag, f.scalefacs[gr][ch]);
            if (part2 < 0 || uint32_t(part2) > c.part2_3_length) return fe->clear(), SYMGPU_ERR_DECODE;
            const int rz = symgpu::mp3e::read_huffman(bs, host_tables().set, c, uint32_t(c.part2_3_length) - uint32_t(part2), quant + (gr * 2 + ch) * 576);
            if (rz < 0) return fe->clear(), SYMGPU_ERR_DECODE;
            f.rzero[gr][ch] = uint16_t(rz);
            part_begin += c.part2_3_length;
        }
    }
    const size_t used = (part_begin + 7) >> 3;
    fe->consumed = std::min(fe->len, fe->consumed + used);

    // GranuleChannel -> symgpu_mp3_gc
    const uint8_t frame_flags = uint8_t((mpeg1 ? SYMGPU_MP3_F_MPEG1 : 0) | (h.mode == MpaMode::JointStereo && h.mid_side ? SYMGPU_MP3_F_MID_SIDE : 0) |
                                        (intensity ? SYMGPU_MP3_F_INTENSITY : 0));
    for (int gr = 0; gr < 2; ++gr)
        for (int ch = 0; ch < 2; ++ch) {
            symgpu_mp3_gc& u = units[gr * 2 + ch];
            std::memset(&u, 0, sizeof u);
            u.sample_rate_idx = h.sample_rate_idx;
            if (gr >= n_gr || ch >= n_ch) {
                u.flags = uint8_t(frame_flags | SYMGPU_MP3_F_MUTE);
                continue;
            }
            const GcSide& c = f.gc[gr][ch];
            u.rzero = f.rzero[gr][ch], u.global_gain = c.global_gain, u.block_type = c.block_type;
            u.flags = uint8_t(frame_flags | (c.mixed ? SYMGPU_MP3_F_MIXED : 0) | (c.scalefac_scale ? SYMGPU_MP3_F_SCALEFAC_SCALE : 0) |
                              (c.preflag ? SYMGPU_MP3_F_PREFLAG : 0) | ((c.scalefac_compress & 1) ? SYMGPU_MP3_F_SFC_LSB : 0));
            std::memcpy(u.subblock_gain, c.subblock_gain, 3);
            std::memcpy(u.scalefacs, f.scalefacs[gr][ch], 39);
        }
    if (info) {
        info->sample_rate = h.sample_rate, info->channels = uint8_t(n_ch), info->granules = uint8_t(n_gr);
        info->sample_rate_idx = h.sample_rate_idx, info->version = uint8_t(h.version);
        info->underflow_bytes = underflow, info->main_data_bytes = uint32_t(used);
    }
    return SYMGPU_OK;
}

extern "C" symgpu_status symgpu_mp3_fe_decode_packets(symgpu_mp3_fe* fe, const uint8_t* data, size_t n, const symgpu_mpa_packet* packets,
                                                      size_t n_packets, symgpu_mp3_gc* units, int16_t* quant, uint32_t* frame_of,
                                                      size_t* n_good, symgpu_mp3_frame_info* info) {
    if (!fe || !data || !n_good || (n_packets && (!packets || !units || !quant || !frame_of))) return SYMGPU_ERR_ARG;
    size_t good = 0;
    for (size_t i = 0; i < n_packets; ++i) {
        if (packets[i].offset > n || packets[i].size > n - packets[i].offset) return SYMGPU_ERR_ARG;
        symgpu_mp3_frame_info fi;
        if (symgpu_mp3_fe_decode(fe, data + packets[i].offset, packets[i].size, units + good * 4, quant + good * 4 * 576, &fi) != SYMGPU_OK) continue;
        if (good == 0 && info) *info = fi;
        frame_of[good++] = uint32_t(i);
    }
    *n_good = good;
    return SYMGPU_OK;
}

// ---------------------------------------------------------------------------------------------- plan + jobs
static_assert(sizeof(symgpu_mp3_gc_job) == sizeof(symgpu::mp3e::GcJob), "the public job record is the GcJob");

extern "C" symgpu_status symgpu_mp3_entropy_plan(const uint8_t* data, size_t n, const symgpu_mpa_packet* packets, size_t n_packets,
                                                 const uint8_t* bad, uint8_t* md, size_t md_cap, size_t* md_len, symgpu_mp3_gc_job* jobs_out,
                                                 uint32_t* frame_of, size_t* n_good, symgpu_mp3_frame_info* info) {
    using namespace symgpu::mp3e;
    if ((!data && n) || (n_packets && !packets) || !md_len || !n_good || (!md && md_cap)) return SYMGPU_ERR_ARG;
    GcJob* jobs = reinterpret_cast<GcJob*>(jobs_out);
    Spec spec;
    size_t len = 0, consumed = 0;  // the reservoir, as byte counts only
    size_t md_at = 0, good = 0;
    for (size_t i = 0; i < n_packets; ++i) {
        if (packets[i].offset > n || packets[i].size > n - packets[i].offset) return SYMGPU_ERR_ARG;
        MpaHeader h{};
        const uint8_t* buf = nullptr;
        size_t buf_len = 0;
        if (open_packet(spec, data + packets[i].offset, packets[i].size, h, buf, buf_len) != SYMGPU_OK) continue;
        Frame f{};
        Bits side(buf, buf_len);
        const size_t side_len = h.side_info_len();
        if (!read_side_info(side, h, f) || side_len > buf_len) {
            len = consumed = 0;
            continue;
        }
        const size_t slot = buf_len - side_len, begin = f.main_data_begin;
        if (begin + slot > 2048) continue;  // refused before the reservoir is touched (mod.rs:49-51)
        const size_t unread = len - consumed;
        const size_t reuse = begin <= unread ? begin : unread;
        const uint32_t underflow = uint32_t(begin - reuse);
        if (bad && bad[i] == 1) {  // known to fail while its main data is read: the reference then empties the reservoir (mod.rs:409-414)
            len = consumed = 0;
            continue;
        }
        const bool leave_out = bad && bad[i] == 2;  // refused AFTER its main data was read (stereo.rs:503-505): the reservoir moves on, no audio
        if (md_at + slot > md_cap) return SYMGPU_ERR_LIMIT;
        if (md) std::memcpy(md + md_at, buf + side_len, slot);
        const uint64_t seg_begin = md_at - reuse;
        const uint32_t seg_len = uint32_t(reuse + slot);
        md_at += slot;
        len = seg_len, consumed = 0;

        const int n_ch = h.n_channels(), n_gr = h.n_granules();
        const bool mpeg1 = h.version == MpaVersion::Mpeg1;
        const bool intensity = h.mode == symgpu::packet::MpaMode::JointStereo && h.intensity;
        const uint8_t frame_flags = uint8_t((mpeg1 ? SYMGPU_MP3_F_MPEG1 : 0) | (h.mode == symgpu::packet::MpaMode::JointStereo && h.mid_side ? SYMGPU_MP3_F_MID_SIDE : 0) |
                                            (intensity ? SYMGPU_MP3_F_INTENSITY : 0));
        const uint32_t underflow_bits = 8 * underflow;
        uint32_t skipped = 0, gr0_begin[2] = {~0u, ~0u};
        size_t part_begin = 0;
        for (int gr = 0; gr < 2; ++gr) {
            const bool silent = gr < n_gr && skipped < underflow_bits;
            for (int ch = 0; ch < 2; ++ch) {
                GcJob j{};
                j.seg_begin = seg_begin, j.seg_len = seg_len, j.out_index = uint32_t(good * 4 + gr * 2 + ch);
                j.unit_flags = frame_flags, j.sample_rate_idx = h.sample_rate_idx, j.mpeg1 = mpeg1, j.gr0_bit_begin = ~0u;
                if (gr >= n_gr || ch >= n_ch) {
                    j.kind = kJobMute;
                } else {
                    j.side = f.gc[gr][ch];
                    j.intensity_channel = ch > 0 && intensity, j.scfsi = uint8_t(f.scfsi[ch]);
                    if (silent) {
                        j.kind = kJobSilent;
                        skipped += f.gc[gr][ch].part2_3_length;
                    } else {
                        j.kind = kJobDecode;
                        j.bit_begin = uint32_t(part_begin);
                        if (gr == 0) gr0_begin[ch] = j.bit_begin;
                        else {
                            j.gr0_bit_begin = gr0_begin[ch];
                            j.gr0_scalefac_compress = f.gc[0][ch].scalefac_compress, j.gr0_block_type = f.gc[0][ch].block_type, j.gr0_mixed = f.gc[0][ch].mixed;
                        }
                        part_begin += f.gc[gr][ch].part2_3_length;
                    }
                }
                if (jobs && !leave_out) jobs[good * 4 + gr * 2 + ch] = j;
            }
            if (silent && skipped > underflow_bits) part_begin = skipped - underflow_bits;
        }
        consumed = std::min(len, (part_begin + 7) >> 3);
        if (leave_out) continue;
        if (good == 0 && info) {
            info->sample_rate = h.sample_rate, info->channels = uint8_t(n_ch), info->granules = uint8_t(n_gr);
            info->sample_rate_idx = h.sample_rate_idx, info->version = uint8_t(h.version), info->underflow_bytes = underflow;
            info->main_data_bytes = uint32_t((part_begin + 7) >> 3);
        }
        if (frame_of) frame_of[good] = uint32_t(i);
        ++good;
    }
    *md_len = md_at, *n_good = good;
    return SYMGPU_OK;
}

extern "C" symgpu_status symgpu_mp3_entropy_run_cpu(const uint8_t* md, size_t md_len, const symgpu_mp3_gc_job* jobs_in, size_t n_jobs,
                                                    symgpu_mp3_gc* units, int16_t* quant, uint8_t* failed) {
    using namespace symgpu::mp3e;
    if ((n_jobs && (!jobs_in || !units || !quant)) || (!md && md_len)) return SYMGPU_ERR_ARG;
    const GcJob* jobs = reinterpret_cast<const GcJob*>(jobs_in);
    const HuffSet& hs = host_tables().set;
    const uint32_t first = n_jobs ? jobs[0].out_index & ~3u : 0;
    for (size_t k = 0; k < n_jobs; ++k) {
        const GcJob& j = jobs[k];
        if (j.out_index < first || j.seg_begin > md_len || j.seg_len > md_len - j.seg_begin) return SYMGPU_ERR_ARG;
        const uint32_t slot = j.out_index - first;
        if (decode_gc_job(j, md, hs, units + slot, quant + size_t(slot) * 576) && failed) failed[slot >> 2] = 1;
    }
    return SYMGPU_OK;
}

extern "C" symgpu_status symgpu_mp3_entropy_run_cpu_mt(const uint8_t* md, size_t md_len, const symgpu_mp3_gc_job* jobs, size_t n_jobs,
                                                       symgpu_mp3_gc* units, int16_t* quant, uint8_t* failed, uint32_t n_threads) {
    if (n_jobs % 4) return SYMGPU_ERR_ARG;
    const size_t n_frames = n_jobs / 4;
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = uint32_t(std::min<size_t>(n_threads, std::max<size_t>(n_frames, 1)));
    if (n_threads <= 1) return symgpu_mp3_entropy_run_cpu(md, md_len, jobs, n_jobs, units, quant, failed);
    host_tables();  // built before the threads start
    std::vector<symgpu_status> status(n_threads, SYMGPU_OK);
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t] {
            const size_t a = n_frames * t / n_threads, b = n_frames * (t + 1) / n_threads;  // whole frames: a job's slot is relative to its range's first frame
            status[t] = symgpu_mp3_entropy_run_cpu(md, md_len, jobs + a * 4, (b - a) * 4, units + a * 4, quant + a * 4 * 576, failed ? failed + a : nullptr);
        });
    for (auto& th : pool) th.join();
    for (symgpu_status s : status)
        if (s != SYMGPU_OK) return s;
    return SYMGPU_OK;
}

extern "C" symgpu_status symgpu_mp3_entropy_decode_cpu(const uint8_t* data, size_t n, const symgpu_mpa_packet* packets, size_t n_packets,
                                                       symgpu_mp3_gc* units, int16_t* quant, uint32_t* frame_of, size_t* n_good,
                                                       symgpu_mp3_frame_info* info, uint32_t* n_rounds) {
    if ((!data && n) || !n_good || (n_packets && (!packets || !units || !quant || !frame_of))) return SYMGPU_ERR_ARG;
    size_t total = 0;
    for (size_t i = 0; i < n_packets; ++i) total += packets[i].size;
    std::vector<uint8_t> md(total + 8), bad(n_packets, 0), failed;
    std::vector<symgpu_mp3_gc_job> jobs(n_packets * 4);
    uint32_t rounds = 0;
    for (;;) {
        ++rounds;
        size_t md_len = 0, good = 0;
        const symgpu_status s = symgpu_mp3_entropy_plan(data, n, packets, n_packets, bad.data(), md.data(), md.size(), &md_len, jobs.data(), frame_of, &good, info);
        if (s != SYMGPU_OK) return s;
        failed.assign(good, 0);
        const symgpu_status r = symgpu_mp3_entropy_run_cpu(md.data(), md_len, jobs.data(), good * 4, units, quant, failed.data());
        if (r != SYMGPU_OK) return r;
        // only the FIRST failure is certain: frames behind it were planned with a reservoir the reference would have emptied
        size_t first_bad = good;
        for (size_t f = 0; f < good; ++f)
            if (failed[f]) {
                first_bad = f;
                break;
            }
        if (first_bad == good) {
            *n_good = good;
            break;
        }
        bad[frame_of[first_bad]] = 1;
    }
    if (n_rounds) *n_rounds = rounds;
    return SYMGPU_OK;
}
