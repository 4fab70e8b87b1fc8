// Layer III entropy decoding of ONE granule-channel -- scale factors (part 2) and the Huffman-coded spectrum (part 3) --
// written once for host and device.  The CPU front-end (mp3_frontend.cpp) calls these functions frame after frame; the
// device path (mp3_entropy_kernel.cu) calls the SAME functions with one thread per granule-channel, which is possible
// because a granule-channel's first bit is known from the side information alone: the sum of the part2_3_length fields
// before it (layer3/mod.rs:272-370 advances `part2_3_begin` by exactly that).  So every CPU test of the front-end is a
// test of the code the kernel runs.
//
// Reference: read_scale_factors_mpeg1 / _mpeg2 (symphonia-bundle-mp3/src/layer3/bitstream.rs:240-427),
// read_huffman_samples (layer3/requantize.rs:47-237), BitReaderLtr::read_codebook (symphonia-core/src/io/bit.rs:771-808).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/symgpu.h"

#ifdef __CUDACC__
#define SYMGPU_HD __host__ __device__ __forceinline__
#else
#define SYMGPU_HD inline
#endif
#ifdef __CUDA_ARCH__
#define SYMGPU_UNROLL _Pragma("unroll")
#else
#define SYMGPU_UNROLL
#endif

namespace symgpu {
namespace mp3e {

// All Huffman lookup tables in one flat array (host: built once; device: a copy in global memory).
// Table t (0..31 big values by table_select, 32 / 33 the two quad tables) starts at base[t]; first_bits[t] == 0 marks the
// tables that hold no codes (0, 4, 14).  Entry: bits 0-7 value, 8-12 code length | bit 31: bits 0-23 offset (from the
// table's base) of a second-level table indexed by the next (bits 24-28) bits.
struct HuffSet {
    const uint32_t* lut;
    uint32_t base[34];
    uint8_t first_bits[34];
    uint8_t linbits[32];
};

// What the side information says about one granule-channel (GranuleChannel, layer3/mod.rs:145-205).
struct GcSide {
    uint16_t part2_3_length, big_values, scalefac_compress;
    uint16_t region1_start, region2_start;
    uint8_t global_gain, block_type, mixed, preflag, scalefac_scale, count1table;
    uint8_t subblock_gain[3], table_select[3];
};

// Most-significant-bit-first reader over [p, p + n_bits / 8): reads past the end fail, the window pads with zeros.
struct Bits {
    const uint8_t* p;
    size_t n_bits;
    size_t at;
    SYMGPU_HD Bits(const uint8_t* data, size_t n_bytes, size_t start_bit = 0) : p(data), n_bits(n_bytes * 8), at(start_bit) {}
    SYMGPU_HD uint32_t window() const {  // the next 32 bits, left-aligned
        const size_t byte = at >> 3, n = n_bits >> 3;
#if !defined(__CUDA_ARCH__) && !defined(SYMGPU_MP3E_DEVICE_WINDOW)
        if (byte + 8 <= n) {  // host fast path: one unaligned load (SYMGPU_MP3E_DEVICE_WINDOW: build the host code with the device's path, for tests)
            uint64_t w;
            __builtin_memcpy(&w, p + byte, 8);
            return uint32_t((__builtin_bswap64(w) << (at & 7)) >> 32);
        }
#endif
        uint64_t v = 0;
SYMGPU_UNROLL
        for (int k = 0; k < 5; ++k) v = v << 8 | (byte + k < n ? p[byte + k] : 0);
        return uint32_t((v << (at & 7)) >> 8);
    }
    SYMGPU_HD size_t left() const { return n_bits - at; }
    SYMGPU_HD bool read(unsigned width, uint32_t& v) {  // width <= 25
        if (width > left()) return false;
        v = width ? window() >> (32 - width) : 0;
        at += width;
        return true;
    }
    SYMGPU_HD bool skip(size_t width) {
        if (width > left()) return false;
        at += width;
        return true;
    }
};

SYMGPU_HD void huff_decode(const HuffSet& hs, int table, uint32_t win, unsigned& value, unsigned& len) {
    const uint32_t* lut = hs.lut + hs.base[table];
    const unsigned first = hs.first_bits[table];
    uint32_t e = lut[win >> (32 - first)];
    if (e & 0x80000000u) {
        const unsigned sub = (e >> 24) & 31;
        e = lut[(e & 0xffffff) + ((win << first) >> (32 - sub))];
    }
    value = e & 0xff, len = (e >> 8) & 31;
}

// Part 2 of an MPEG-1 granule-channel.  `copy_from`: granule 0's scale factors of the same channel when this is granule
// 1 (groups flagged in scfsi are copied instead of read), else null.  Returns the bits read, -1 when they run out.
SYMGPU_HD int read_scale_factors_mpeg1(Bits& bs, const GcSide& c, const uint8_t* copy_from, unsigned scfsi_mask, uint8_t* scalefacs) {
    const unsigned sfc = c.scalefac_compress & 15;
    // slen1 = 0 0 0 0 3 1 1 1 2 2 2 3 3 3 4 4 and slen2 = 0 1 2 3 0 1 2 3 1 2 3 1 2 3 2 3 by scalefac_compress, one nibble each
    const unsigned s1 = (0x4433322211130000ull >> (4 * sfc)) & 15, s2 = (0x3232132132103210ull >> (4 * sfc)) & 15;
    uint32_t v;
    if (c.block_type == 2) {
        const int n1 = c.mixed ? 17 : 18;
        if (s1)
            for (int i = 0; i < n1; ++i) {
                if (!bs.read(s1, v)) return -1;
                scalefacs[i] = uint8_t(v);
            }
        if (s2)
            for (int i = n1; i < n1 + 18; ++i) {
                if (!bs.read(s2, v)) return -1;
                scalefacs[i] = uint8_t(v);
            }
        return n1 * int(s1) + 18 * int(s2);
    }
    int bits = 0;
    for (int g = 0; g < 4; ++g) {
        const int a = g == 0 ? 0 : g == 1 ? 6 : g == 2 ? 11 : 16, b = g == 0 ? 6 : g == 1 ? 11 : g == 2 ? 16 : 21;
        const unsigned s = g < 2 ? s1 : s2;
        if (copy_from && (scfsi_mask >> g & 1)) {
            for (int i = a; i < b; ++i) scalefacs[i] = copy_from[i];
        } else if (s) {
            for (int i = a; i < b; ++i) {
                if (!bs.read(s, v)) return -1;
                scalefacs[i] = uint8_t(v);
            }
            bits += int(s) * (b - a);
        }
    }
    return bits;
}

// Part 2 of an MPEG-2 / 2.5 granule-channel; sets *preflag for a channel that is not the intensity channel.
SYMGPU_HD int read_scale_factors_mpeg2(Bits& bs, bool intensity_channel, const GcSide& c, uint8_t* preflag, uint8_t* scalefacs) {
    // band counts per partition [table row][long | short | mixed], ISO 13818-3 2.4.3.2
    const uint8_t nsfb[6][3][4] = {
        {{7, 7, 7, 0}, {12, 12, 12, 0}, {6, 15, 12, 0}}, {{6, 6, 6, 3}, {12, 9, 9, 6}, {6, 12, 9, 6}}, {{8, 8, 5, 0}, {15, 12, 9, 0}, {6, 18, 9, 0}},
        {{6, 5, 5, 5}, {9, 9, 9, 9}, {6, 9, 9, 9}},      {{6, 5, 7, 3}, {9, 9, 12, 6}, {6, 9, 12, 6}}, {{11, 10, 0, 0}, {18, 18, 0, 0}, {15, 18, 0, 0}}};
    const int block = c.block_type == 2 ? (c.mixed ? 2 : 1) : 0;
    unsigned slen[4] = {0, 0, 0, 0};
    int row;
    if (intensity_channel) {
        const unsigned sfc = c.scalefac_compress >> 1;
        if (sfc < 180) row = 0, slen[0] = sfc / 36, slen[1] = (sfc % 36) / 6, slen[2] = (sfc % 36) % 6;
        else if (sfc < 244) row = 1, slen[0] = ((sfc - 180) % 64) >> 4, slen[1] = ((sfc - 180) % 16) >> 2, slen[2] = (sfc - 180) % 4;
        else row = 2, slen[0] = (sfc - 244) / 3, slen[1] = (sfc - 244) % 3;
    } else {
        const unsigned sfc = c.scalefac_compress;
        *preflag = sfc >= 500;
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
                scalefacs[i] = uint8_t(v);
            }
            bits += int(slen[k]) * n;
        }
        start += n;
    }
    return bits;
}

// Part 3: the spectrum as sign * x (the reference writes sign * POW43[x], requantize.rs:128, :144); all 576 lines of `q`
// are written.  Returns rzero, -1 on an over-read.
SYMGPU_HD int read_huffman(Bits& bs, const HuffSet& hs, const GcSide& c, uint32_t part3_bits, int16_t* q) {
    if (part3_bits == 0) {
        for (int k = 0; k < 576; ++k) q[k] = 0;
        return 0;
    }
    const size_t end = bs.at + part3_bits;  // the reference's "bits_read < part3_bits" is "bs.at < end"
    const int big_len = 2 * int(c.big_values);
    int i = 0;
    for (int r = 0; r < 3; ++r) {
        const int limit = r == 0 ? c.region1_start : r == 1 ? c.region2_start : 576;
        const int region_end = limit < big_len ? limit : big_len;
        const int table = c.table_select[r];
        const unsigned linbits = hs.linbits[table];
        if (hs.first_bits[table] == 0) {  // tables 0, 4, 14: a silent region that costs no bits
            for (; i < region_end; ++i) q[i] = 0;
            continue;
        }
        while (i < region_end && bs.at < end) {
            unsigned value, len;
            const uint32_t win = bs.window();
            huff_decode(hs, table, win, value, len);
            if (len > bs.left()) return -1;
            if (!linbits) {
                // tables without linbits (0..15): the code (<= 19 bits) and its at most two sign bits sit in one window
                const unsigned x = value >> 4, y = value & 15;
                const unsigned need = len + (x != 0) + (y != 0);
                if (need > bs.left()) return -1;
                uint32_t tail = win << len;  // the bits behind the code, left-aligned
                int16_t ox = 0, oy = 0;
                if (x) ox = int16_t((tail >> 31) ? -int(x) : int(x)), tail <<= 1;
                if (y) oy = int16_t((tail >> 31) ? -int(y) : int(y));
                q[i] = ox, q[i + 1] = oy;
                bs.at += need;
                i += 2;
                continue;
            }
            bs.at += len;
SYMGPU_UNROLL
            for (int k = 0; k < 2; ++k) {
                unsigned x = k == 0 ? value >> 4 : value & 15;
                int16_t out = 0;
                if (x) {
                    uint32_t extra, sign;
                    if (x == 15) {
                        if (!bs.read(linbits, extra)) return -1;
                        x += extra;
                    }
                    if (!bs.read(1, sign)) return -1;
                    out = int16_t(sign ? -int(x) : int(x));
                }
                q[i + k] = out;
            }
            i += 2;
        }
    }
    const int quad = 32 + c.count1table;
    while (i <= 572 && bs.at < end) {
        unsigned value, len;
        huff_decode(hs, quad, bs.window(), value, len);
        if (len > bs.left()) return -1;
        bs.at += len;
        const unsigned ones = (value >> 3 & 1) + (value >> 2 & 1) + (value >> 1 & 1) + (value & 1);
        uint32_t signs;
        if (!bs.read(ones, signs)) return -1;
        // sign bits come in the order v, w, x, y; the reference peels them off from the last one (requantize.rs:170-203)
SYMGPU_UNROLL
        for (int k = 3; k >= 0; --k) {
            int16_t out = 0;
            if (value & (1u << (3 - k))) {
                out = (signs & 1) ? -1 : 1;
                signs >>= 1;
            }
            q[i + k] = out;
        }
        i += 4;
    }
    if (bs.at < end) {
        if (!bs.skip(end - bs.at)) return -1;  // stuffing
    } else if (bs.at > end && i > big_len) {
        i -= 4;  // the last quad came out of bits that belong to the next granule: undo it (requantize.rs:222-226)
    }
    for (int k = i; k < 576; ++k) q[k] = 0;
    return i;
}

// ---- one granule-channel as a unit of parallel work ---------------------------------------------------------------
// The host walks the frames once (headers, side information, bit-reservoir arithmetic: none of it needs the Huffman
// data) and emits one job per unit slot; the main data of all frames is compacted into one byte stream `md`, in which
// a frame's reservoir -- the bytes main_data_begin reaches back to plus its own -- is one contiguous window.
enum : uint8_t { kJobDecode = 0, kJobSilent = 1, kJobMute = 2 };
struct GcJob {  // 64 bytes
    uint64_t seg_begin;      // byte offset in md of the frame's reservoir window
    uint32_t seg_len;        // its length: reads beyond it fail exactly where the reference's reservoir ends
    uint32_t bit_begin;      // first bit of this granule-channel's part 2, from seg_begin
    uint32_t gr0_bit_begin;  // MPEG-1 granule 1 with scfsi: where granule 0's part 2 (same channel) starts; ~0u = it was never read
    uint32_t out_index;      // unit slot: frame * 4 + granule * 2 + channel
    GcSide side;             // 22 bytes
    uint16_t gr0_scalefac_compress;
    uint8_t gr0_block_type, gr0_mixed;
    uint8_t kind;            // kJobDecode | kJobSilent (bits lost to an underflow: zeros, side information kept) | kJobMute (absent unit)
    uint8_t mpeg1, intensity_channel, scfsi;
    uint8_t unit_flags;      // frame-level SYMGPU_MP3_F_* bits
    uint8_t sample_rate_idx;
    uint8_t reserved[8];
};
static_assert(sizeof(GcJob) == 64, "GcJob is 64 bytes");

// Fills one unit and its 576 quantised lines.  Returns 0, or 1 when the reference would refuse the frame here
// ("part2_3_length is not valid", "huffman decode overrun", an offset past the reservoir: layer3/mod.rs:318-358).
SYMGPU_HD int decode_gc_job(const GcJob& j, const uint8_t* md, const HuffSet& hs, symgpu_mp3_gc* unit, int16_t* q) {
    symgpu_mp3_gc u;
    u.rzero = 0, u.global_gain = 0, u.block_type = 0, u.flags = j.unit_flags, u.sample_rate_idx = j.sample_rate_idx;
    for (int k = 0; k < 3; ++k) u.subblock_gain[k] = 0;
    for (int k = 0; k < 39; ++k) u.scalefacs[k] = 0;
    for (int k = 0; k < 16; ++k) u.reserved[k] = 0;
    int status = 0;
    if (j.kind == kJobMute) {
        u.flags |= SYMGPU_MP3_F_MUTE;
        for (int k = 0; k < 576; ++k) q[k] = 0;
    } else {
        const GcSide& c = j.side;
        uint8_t preflag = c.preflag;
        int rzero = 0;
        if (j.kind == kJobSilent) {
            for (int k = 0; k < 576; ++k) q[k] = 0;
        } else {
            const uint8_t* seg = md + j.seg_begin;
            Bits bs(seg, j.seg_len, j.bit_begin);
            int part2 = -1;
            if ((j.bit_begin >> 3) <= j.seg_len && bs.at <= bs.n_bits) {
                if (j.mpeg1) {
                    uint8_t first[39];
                    const uint8_t* copy_from = nullptr;
                    if (j.scfsi && c.block_type != 2 && (j.out_index & 2)) {  // granule 1 repeating groups of granule 0
                        for (int k = 0; k < 39; ++k) first[k] = 0;
                        copy_from = first;
                        if (j.gr0_bit_begin != ~0u) {
                            GcSide g0 = c;
                            g0.scalefac_compress = j.gr0_scalefac_compress, g0.block_type = j.gr0_block_type, g0.mixed = j.gr0_mixed;
                            Bits b0(seg, j.seg_len, j.gr0_bit_begin);
                            read_scale_factors_mpeg1(b0, g0, nullptr, 0, first);  // its own failure is granule 0's job to report
                        }
                    }
                    part2 = read_scale_factors_mpeg1(bs, c, copy_from, j.scfsi, u.scalefacs);
                } else {
                    part2 = read_scale_factors_mpeg2(bs, j.intensity_channel != 0, c, &preflag, u.scalefacs);
                }
            }
            if (part2 < 0 || uint32_t(part2) > c.part2_3_length) {
                status = 1;
                for (int k = 0; k < 576; ++k) q[k] = 0;
            } else {
                rzero = read_huffman(bs, hs, c, uint32_t(c.part2_3_length) - uint32_t(part2), q);
                if (rzero < 0) status = 1, rzero = 0;
            }
        }
        u.rzero = uint16_t(rzero), u.global_gain = c.global_gain, u.block_type = c.block_type;
        u.flags |= uint8_t((c.mixed ? SYMGPU_MP3_F_MIXED : 0) | (c.scalefac_scale ? SYMGPU_MP3_F_SCALEFAC_SCALE : 0) | (preflag ? SYMGPU_MP3_F_PREFLAG : 0) |
                           ((c.scalefac_compress & 1) ? SYMGPU_MP3_F_SFC_LSB : 0));
        for (int k = 0; k < 3; ++k) u.subblock_gain[k] = c.subblock_gain[k];
    }
    *unit = u;
    return status;
}

}  // namespace mp3e

// The flat tables (host memory, built once; thread-safe).  `words` = length of lut.
const mp3e::HuffSet& mp3_huffset_host(size_t* words);

}  // namespace symgpu
