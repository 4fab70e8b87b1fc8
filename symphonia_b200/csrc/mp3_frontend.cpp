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
#include <vector>

#include "../../include/symgpu.h"
#include "../../include/symgpu/packetizer.hpp"
#include "tables.h"

namespace {

#include "mp3_huffman_data.inc"

using symgpu::packet::MpaHeader;
using symgpu::packet::MpaVersion;

// ---------------------------------------------------------------------------------------------- bit reader
struct Bits {
    const uint8_t* p;
    size_t n_bits;
    size_t at = 0;
    Bits(const uint8_t* data, size_t n_bytes) : p(data), n_bits(n_bytes * 8) {}
    // The next 32 bits, left-aligned at bit 31, zeros past the end.
    uint32_t window() const {
        const size_t byte = at >> 3, n = n_bits >> 3;
        uint64_t v = 0;
        if (byte + 8 <= n) {
            std::memcpy(&v, p + byte, 8);
            v = __builtin_bswap64(v);
        } else {
            for (size_t k = 0; k < 8; ++k) v = v << 8 | (byte + k < n ? p[byte + k] : 0);
        }
        return uint32_t((v << (at & 7)) >> 32);
    }
    size_t left() const { return n_bits - at; }
    // false: not enough bits (the reference's reader errors; nothing is consumed here)
    bool read(unsigned width, uint32_t& v) {
        if (width > left()) return false;
        v = width ? window() >> (32 - width) : 0;
        at += width;
        return true;
    }
    bool skip(size_t width) {
        if (width > left()) return false;
        at += width;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------- Huffman tables
// entry: bits 0-7 value, 8-12 code length (direct) | bit 31 set: bits 0-23 offset of a second-level table whose
// index width is in bits 24-28.
struct HuffTable {
    std::vector<uint32_t> lut;
    unsigned first_bits = 0;
    bool empty() const { return lut.empty(); }
    // value and length of the code at the head of `win` (left-aligned 32-bit window)
    inline void decode(uint32_t win, unsigned& value, unsigned& len) const {
        uint32_t e = lut[win >> (32 - first_bits)];
        if (e & 0x80000000u) {
            const unsigned sub = (e >> 24) & 31;
            e = lut[(e & 0xffffff) + ((win << first_bits) >> (32 - sub))];
        }
        value = e & 0xff, len = (e >> 8) & 31;
    }
};

constexpr unsigned kFirstBits = 9;

HuffTable build(const uint32_t* packed, size_t n, unsigned wrap, bool quad) {
    HuffTable t;
    unsigned max_len = 0;
    for (size_t i = 0; i < n; ++i) max_len = std::max(max_len, packed[i] >> 24);
    t.first_bits = std::min(max_len, kFirstBits);
    t.lut.assign(size_t(1) << t.first_bits, 0);
    // longest code under each first-level prefix
    std::vector<unsigned> deepest(t.lut.size(), 0);
    for (size_t i = 0; i < n; ++i) {
        const unsigned len = packed[i] >> 24, code = packed[i] & 0x7ffff;
        if (len > t.first_bits) {
            const unsigned prefix = code >> (len - t.first_bits);
            deepest[prefix] = std::max(deepest[prefix], len - t.first_bits);
        }
    }
    for (size_t pfx = 0; pfx < deepest.size(); ++pfx)
        if (deepest[pfx]) {
            t.lut[pfx] = 0x80000000u | uint32_t(t.lut.size()) | (deepest[pfx] << 24);
            t.lut.resize(t.lut.size() + (size_t(1) << deepest[pfx]), 0);
        }
    for (size_t i = 0; i < n; ++i) {
        const unsigned len = packed[i] >> 24, code = packed[i] & 0x7ffff;
        const unsigned value = quad ? unsigned(i) : unsigned(((i / wrap) << 4) | (i % wrap));
        const uint32_t entry = value | (len << 8);
        if (len <= t.first_bits) {
            const unsigned pad = t.first_bits - len;
            for (unsigned k = 0; k < (1u << pad); ++k) t.lut[(code << pad) + k] = entry;
        } else {
            const unsigned rest = len - t.first_bits, prefix = code >> rest;
            const unsigned sub = (t.lut[prefix] >> 24) & 31, base = t.lut[prefix] & 0xffffff, pad = sub - rest;
            const unsigned low = code & ((1u << rest) - 1);
            for (unsigned k = 0; k < (1u << pad); ++k) t.lut[base + (low << pad) + k] = entry;
        }
    }
    return t;
}

struct Tables {
    HuffTable big[32];  // by table_select; 4 and 14 stay empty
    HuffTable quad[2];
    uint8_t linbits[32] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13};
    Tables() {
#define SYMGPU_BIG(T, W) big[T] = build(kHuff_##T, sizeof(kHuff_##T) / 4, W, false)
        SYMGPU_BIG(1, 2), SYMGPU_BIG(2, 3), SYMGPU_BIG(3, 3), SYMGPU_BIG(5, 4), SYMGPU_BIG(6, 4), SYMGPU_BIG(7, 6), SYMGPU_BIG(8, 6), SYMGPU_BIG(9, 6);
        SYMGPU_BIG(10, 8), SYMGPU_BIG(11, 8), SYMGPU_BIG(12, 8), SYMGPU_BIG(13, 16), SYMGPU_BIG(15, 16), SYMGPU_BIG(16, 16), SYMGPU_BIG(24, 16);
#undef SYMGPU_BIG
        for (int t = 17; t < 24; ++t) big[t] = big[16];  // 16..23 and 24..31 share codes and differ in linbits
        for (int t = 25; t < 32; ++t) big[t] = big[24];
        quad[0] = build(kHuff_quadA, 16, 16, true), quad[1] = build(kHuff_quadB, 16, 16, true);
    }
};
const Tables& tables() {
    static const Tables t;
    return t;
}

// ---------------------------------------------------------------------------------------------- frame data
struct Gc {  // GranuleChannel (layer3/mod.rs:145-205), side-information part
    uint16_t part2_3_length, big_values, scalefac_compress;
    uint8_t global_gain, block_type, mixed, subblock_gain[3], table_select[3], preflag, scalefac_scale, count1table;
    uint16_t region1_start, region2_start, rzero;
    uint8_t scalefacs[39];
};
struct Frame {
    unsigned main_data_begin;
    bool scfsi[2][4];
    Gc gc[2][2];
};

// bitstream.rs:57-195
bool read_side_info(Bits& bs, const MpaHeader& h, Frame& f) {
    const bool mpeg1 = h.version == MpaVersion::Mpeg1;
    const int n_ch = h.n_channels(), n_gr = h.n_granules();
    const uint16_t* long_edges = symgpu::mp3_tables_host().edges[h.sample_rate_idx][symgpu::kKindLong];
    uint32_t v;
    if (mpeg1) {
        if (!bs.read(9, v)) return false;
        f.main_data_begin = v;
        if (!bs.skip(n_ch == 1 ? 5 : 3)) return false;
        for (int ch = 0; ch < n_ch; ++ch)
            for (int b = 0; b < 4; ++b) {
                if (!bs.read(1, v)) return false;
                f.scfsi[ch][b] = v != 0;
            }
    } else {
        if (!bs.read(8, v)) return false;
        f.main_data_begin = v;
        if (!bs.skip(n_ch == 1 ? 1 : 2)) return false;
    }
    for (int gr = 0; gr < n_gr; ++gr)
        for (int ch = 0; ch < n_ch; ++ch) {
            Gc& c = f.gc[gr][ch];
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

// bitstream.rs:240-318: part 2 of an MPEG-1 granule-channel.  Returns the bits used, -1 when they run out.
int read_scale_factors_mpeg1(Bits& bs, int gr, int ch, Frame& f) {
    static const uint8_t slen[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {3, 0}, {1, 1}, {1, 2}, {1, 3},
                                        {2, 1}, {2, 2}, {2, 3}, {3, 1}, {3, 2}, {3, 3}, {4, 2}, {4, 3}};
    Gc& c = f.gc[gr][ch];
    const unsigned s1 = slen[c.scalefac_compress][0], s2 = slen[c.scalefac_compress][1];
    int bits = 0;
    uint32_t v;
    if (c.block_type == SYMGPU_MP3_SHORT) {
        const int n1 = c.mixed ? 17 : 18;
        if (s1)
            for (int i = 0; i < n1; ++i) {
                if (!bs.read(s1, v)) return -1;
                c.scalefacs[i] = uint8_t(v);
            }
        if (s2)
            for (int i = n1; i < n1 + 18; ++i) {
                if (!bs.read(s2, v)) return -1;
                c.scalefacs[i] = uint8_t(v);
            }
        return n1 * int(s1) + 18 * int(s2);
    }
    static const uint8_t groups[5] = {0, 6, 11, 16, 21};
    for (int g = 0; g < 4; ++g) {
        const unsigned s = g < 2 ? s1 : s2;
        if (gr > 0 && f.scfsi[ch][g]) {
            std::memcpy(c.scalefacs + groups[g], f.gc[0][ch].scalefacs + groups[g], groups[g + 1] - groups[g]);
        } else if (s) {
            for (int i = groups[g]; i < groups[g + 1]; ++i) {
                if (!bs.read(s, v)) return -1;
                c.scalefacs[i] = uint8_t(v);
            }
            bits += int(s) * (groups[g + 1] - groups[g]);
        }
    }
    return bits;
}

// bitstream.rs:320-427: MPEG-2 / 2.5.  `intensity_channel`: channel 1 of an intensity-stereo frame.
int read_scale_factors_mpeg2(Bits& bs, bool intensity_channel, Gc& c) {
    static const uint8_t nsfb[6][3][4] = {
        {{7, 7, 7, 0}, {12, 12, 12, 0}, {6, 15, 12, 0}}, {{6, 6, 6, 3}, {12, 9, 9, 6}, {6, 12, 9, 6}}, {{8, 8, 5, 0}, {15, 12, 9, 0}, {6, 18, 9, 0}},
        {{6, 5, 5, 5}, {9, 9, 9, 9}, {6, 9, 9, 9}},      {{6, 5, 7, 3}, {9, 9, 12, 6}, {6, 9, 12, 6}}, {{11, 10, 0, 0}, {18, 18, 0, 0}, {15, 18, 0, 0}}};
    const int block = c.block_type == SYMGPU_MP3_SHORT ? (c.mixed ? 2 : 1) : 0;
    unsigned slen[4] = {0, 0, 0, 0};
    int row;
    if (intensity_channel) {
        const unsigned sfc = c.scalefac_compress >> 1;
        if (sfc < 180) row = 0, slen[0] = sfc / 36, slen[1] = (sfc % 36) / 6, slen[2] = (sfc % 36) % 6;
        else if (sfc < 244) row = 1, slen[0] = ((sfc - 180) % 64) >> 4, slen[1] = ((sfc - 180) % 16) >> 2, slen[2] = (sfc - 180) % 4;
        else row = 2, slen[0] = (sfc - 244) / 3, slen[1] = (sfc - 244) % 3;
    } else {
        const unsigned sfc = c.scalefac_compress;
        c.preflag = sfc >= 500;
        if (sfc < 400) row = 3, slen[0] = (sfc >> 4) / 5, slen[1] = (sfc >> 4) % 5, slen[2] = (sfc % 16) >> 2, slen[3] = sfc % 4;
        else if (sfc < 500) row = 4, slen[0] = ((sfc - 400) >> 2) / 5, slen[1] = ((sfc - 400) >> 2) % 5, slen[2] = (sfc - 400) % 4;
        else row = 5, slen[0] = (sfc - 500) / 3, slen[1] = (sfc - 500) % 3;
    }
    int bits = 0, start = 0;
    uint32_t v;
    for (int k = 0; k < 4; ++k) {
        const int n = nsfb[row][block][k];
        if (slen[k]) {
            for (int i = start; i < start + n; ++i) {
                if (!bs.read(slen[k], v)) return -1;
                c.scalefacs[i] = uint8_t(v);
            }
            bits += int(slen[k]) * n;
        }
        start += n;
    }
    return bits;
}

// requantize.rs:47-237 with the magnitudes left as integers.  Returns rzero, or -1 when the data runs out
// ("huffman decode overrun").  `q` gets all 576 lines.
int read_huffman(Bits& bs, const Gc& c, uint32_t part3_bits, int16_t* q) {
    if (part3_bits == 0) return std::memset(q, 0, 576 * sizeof(int16_t)), 0;
    const Tables& T = tables();
    const size_t begin = bs.at, end = begin + part3_bits;  // "bits_read < part3_bits" == bs.at < end
    const int big_len = 2 * int(c.big_values);
    const int region_end[3] = {std::min<int>(c.region1_start, big_len), std::min<int>(c.region2_start, big_len), std::min(576, big_len)};
    int i = 0;
    for (int r = 0; r < 3; ++r) {
        const HuffTable& tab = T.big[c.table_select[r]];
        const unsigned linbits = T.linbits[c.table_select[r]];
        if (tab.empty()) {
            for (; i < region_end[r]; ++i) q[i] = 0;
            continue;
        }
        while (i < region_end[r] && bs.at < end) {
            unsigned value, len;
            uint32_t win = bs.window();
            tab.decode(win, value, len);
            if (len > bs.left()) return -1;
            bs.at += len;
            unsigned xy[2] = {value >> 4, value & 15};
            for (int k = 0; k < 2; ++k) {
                unsigned x = xy[k];
                if (x) {
                    uint32_t extra = 0, sign;
                    if (x == 15 && linbits) {
                        if (!bs.read(linbits, extra)) return -1;
                        x += extra;
                    }
                    if (!bs.read(1, sign)) return -1;
                    q[i + k] = int16_t(sign ? -int(x) : int(x));
                } else {
                    q[i + k] = 0;
                }
            }
            i += 2;
        }
    }
    const HuffTable& quad = T.quad[c.count1table];
    while (i <= 572 && bs.at < end) {
        unsigned value, len;
        quad.decode(bs.window(), value, len);
        if (len > bs.left()) return -1;
        bs.at += len;
        const unsigned ones = unsigned(__builtin_popcount(value & 15));
        uint32_t signs;
        if (!bs.read(ones, signs)) return -1;
        // the sign bits follow in the order v, w, x, y; the reference peels them off from the last one
        for (int k = 3; k >= 0; --k) {
            if (value & (1u << (3 - k))) {
                q[i + k] = (signs & 1) ? -1 : 1;
                signs >>= 1;
            } else {
                q[i + k] = 0;
            }
        }
        i += 4;
    }
    if (bs.at < end) {
        if (!bs.skip(end - bs.at)) return -1;  // stuffing
    } else if (bs.at > end && i > big_len) {
        i -= 4;  // the last quad was read out of bits that belong to the next granule: undo it (requantize.rs:222-226)
    }
    for (int k = i; k < 576; ++k) q[k] = 0;
    return i;
}

}  // namespace

struct symgpu_mp3_fe {
    uint8_t reservoir[2048];
    size_t len = 0, consumed = 0;
    bool have_spec = false;
    uint32_t spec_rate = 0;
    int spec_channels = 0;
    void clear() { len = consumed = 0; }
};

extern "C" symgpu_status symgpu_mp3_fe_create(symgpu_mp3_fe** out) {
    if (!out) return SYMGPU_ERR_ARG;
    tables();
    *out = new (std::nothrow) symgpu_mp3_fe();
    return *out ? SYMGPU_OK : SYMGPU_ERR_LIMIT;
}
extern "C" void symgpu_mp3_fe_destroy(symgpu_mp3_fe* fe) { delete fe; }
extern "C" void symgpu_mp3_fe_reset(symgpu_mp3_fe* fe) {
    if (fe) fe->clear(), fe->have_spec = false;
}

extern "C" symgpu_status symgpu_mp3_fe_decode(symgpu_mp3_fe* fe, const uint8_t* frame, size_t n, symgpu_mp3_gc* units, int16_t* quant,
                                              symgpu_mp3_frame_info* info) {
    using namespace symgpu::packet;
    if (!fe || (!frame && n) || !units || !quant) return SYMGPU_ERR_ARG;
    // decoder.rs:87-93: synchronise inside the packet, parse, and insist that the packet is exactly one frame
    size_t q = 0;
    uint32_t word = 0;
    for (;; ++q) {
        if (q + 4 > n) return SYMGPU_ERR_DECODE;
        word = detail::be32(frame + q);
        if (mpa_is_synced(word) && mpa_check_header(word)) break;
    }
    MpaHeader h{};
    const Status hs = mpa_parse_header(word, h);
    if (hs != Status::Ok) return hs == Status::Unsupported ? SYMGPU_ERR_UNSUPPORTED : SYMGPU_ERR_DECODE;
    const uint8_t* body = frame + q + 4;
    const size_t body_len = n - q - 4;
    if (h.frame_size != body_len) return SYMGPU_ERR_DECODE;
    // decoder.rs:96-108: the signal specification (rate, channel layout) is fixed by the first frame
    if (!fe->have_spec) fe->have_spec = true, fe->spec_rate = h.sample_rate, fe->spec_channels = h.n_channels();
    else if (fe->spec_rate != h.sample_rate || fe->spec_channels != h.n_channels()) return SYMGPU_ERR_DECODE;
    if (h.layer != 3) return SYMGPU_ERR_DECODE;

    const size_t crc_len = h.crc ? 2 : 0;
    if (body_len < crc_len) return SYMGPU_ERR_DECODE;
    const uint8_t* buf = body + crc_len;
    const size_t buf_len = body_len - crc_len;
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
            Bits bs(fe->reservoir, fe->len);
            bs.at = part_begin;  // (a bit offset inside the last byte's padding cannot occur: the byte index is checked above)
            if (bs.at > bs.n_bits) return fe->clear(), SYMGPU_ERR_DECODE;
            Gc& c = f.gc[gr][ch];
            const int part2 = mpeg1 ? read_scale_factors_mpeg1(bs, gr, ch, f) : read_scale_factors_mpeg2(bs, ch > 0 && intensity, c);
            if (part2 < 0 || uint32_t(part2) > c.part2_3_length) return fe->clear(), SYMGPU_ERR_DECODE;
            const int rz = read_huffman(bs, c, uint32_t(c.part2_3_length) - uint32_t(part2), quant + (gr * 2 + ch) * 576);
            if (rz < 0) return fe->clear(), SYMGPU_ERR_DECODE;
            c.rzero = uint16_t(rz);
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
            const Gc& c = f.gc[gr][ch];
            u.rzero = c.rzero, u.global_gain = c.global_gain, u.block_type = c.block_type;
            u.flags = uint8_t(frame_flags | (c.mixed ? SYMGPU_MP3_F_MIXED : 0) | (c.scalefac_scale ? SYMGPU_MP3_F_SCALEFAC_SCALE : 0) |
                              (c.preflag ? SYMGPU_MP3_F_PREFLAG : 0) | ((c.scalefac_compress & 1) ? SYMGPU_MP3_F_SFC_LSB : 0));
            std::memcpy(u.subblock_gain, c.subblock_gain, 3);
            std::memcpy(u.scalefacs, c.scalefacs, 39);
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
