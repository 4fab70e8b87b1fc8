// symgpu packetisers: cut MPEG audio, ADTS and Ogg byte streams into codec packets the way
// the reference's format readers do, but as INDEX BUILDERS over one resident byte buffer.
//
// The reference pulls packets one at a time out of a consuming reader and copies each into its own allocation
// (symphonia-bundle-mp3/src/demuxer.rs:598-604, symphonia-codec-aac/src/adts.rs:303-308,
// symphonia-format-ogg/src/logical.rs:577-597).  A batched device decoder wants the opposite: the whole file
// goes to HBM in ONE copy and the front-end kernels are handed a table of (offset, length) references into it.
// So nothing here copies payload bytes: an MPEG / ADTS packet is a byte range of the source, an Ogg packet --
// which may straddle pages -- is a short gather list of ranges.  The byte RULES (what counts as sync, what is
// skipped as junk, which frames are tags, how packets continue across pages, where the time stamps and trims come
// from) are the reference's, cited at each function, and are checked bit for bit against oracle/packetizer_oracle.py.
//
// Header-only C++17, no dependencies, no device code.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

namespace symgpu {
namespace packet {

enum class Status : uint8_t {
    Ok = 0,
    EndOfStream,  // the bytes ran out (the reference: IoError UnexpectedEof)
    DecodeError,  // malformed stream (Error::DecodeError)
    Unsupported,  // well-formed but not handled by the reference either (Error::Unsupported)
};

// A byte range of the source buffer.
struct Piece {
    uint64_t offset;
    uint32_t len;
};

namespace detail {
inline uint32_t be32(const uint8_t* p) { return uint32_t(p[0]) << 24 | uint32_t(p[1]) << 16 | uint32_t(p[2]) << 8 | p[3]; }
inline uint32_t be24(const uint8_t* p) { return uint32_t(p[0]) << 16 | uint32_t(p[1]) << 8 | p[2]; }
inline uint32_t be16(const uint8_t* p) { return uint32_t(p[0]) << 8 | p[1]; }
inline uint32_t le32(const uint8_t* p) { return uint32_t(p[3]) << 24 | uint32_t(p[2]) << 16 | uint32_t(p[1]) << 8 | p[0]; }
inline uint64_t le64(const uint8_t* p) { return uint64_t(le32(p + 4)) << 32 | le32(p); }

struct Crc32Table {  // slicing-by-8: t[k][b] = contribution of byte b seen k bytes before the end of an 8-byte block
    uint32_t t[8][256];
    constexpr Crc32Table() : t() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i << 24;
            for (int k = 0; k < 8; ++k) c = (c & 0x80000000u) ? (c << 1) ^ 0x04c11db7u : c << 1;
            t[0][i] = c;
        }
        for (int k = 1; k < 8; ++k)
            for (uint32_t i = 0; i < 256; ++i) t[k][i] = (t[k - 1][i] << 8) ^ t[0][t[k - 1][i] >> 24];
    }
};
struct Crc16Table {
    uint16_t t[256];
    constexpr Crc16Table() : t() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0xa001u : c >> 1;
            t[i] = uint16_t(c);
        }
    }
};
}  // namespace detail

// CRC-32, polynomial 0x04c11db7, most-significant bit first, no final xor; the caller supplies the initial state
// (Ogg pages: 0).  symphonia-core/src/checksum/crc32.rs:543-570.
inline uint32_t crc32_update(uint32_t state, const uint8_t* p, size_t n) {
    static constexpr detail::Crc32Table tab{};
    for (; n >= 8; p += 8, n -= 8) {
        const uint32_t hi = state ^ detail::be32(p), lo = detail::be32(p + 4);
        state = tab.t[7][hi >> 24] ^ tab.t[6][(hi >> 16) & 0xff] ^ tab.t[5][(hi >> 8) & 0xff] ^ tab.t[4][hi & 0xff] ^
                tab.t[3][lo >> 24] ^ tab.t[2][(lo >> 16) & 0xff] ^ tab.t[1][(lo >> 8) & 0xff] ^ tab.t[0][lo & 0xff];
    }
    for (size_t i = 0; i < n; ++i) state = (state << 8) ^ tab.t[0][(state >> 24) ^ p[i]];
    return state;
}

// CRC-16, polynomial 0x8005, least-significant bit first, no final xor (the LAME tag's checksum).
// symphonia-core/src/checksum/crc16.rs:377-404.
inline uint16_t crc16_ansi_le_update(uint16_t state, const uint8_t* p, size_t n) {
    static constexpr detail::Crc16Table tab{};
    for (size_t i = 0; i < n; ++i) state = uint16_t((state >> 8) ^ tab.t[(state ^ p[i]) & 0xff]);
    return state;
}

// =====================================================================================================================
// MPEG audio (Layers I-III)
// =====================================================================================================================

enum class MpaVersion : uint8_t { Mpeg1 = 0, Mpeg2 = 1, Mpeg2p5 = 2 };
enum class MpaMode : uint8_t { Stereo = 0, JointStereo = 1, DualMono = 2, Mono = 3 };

// The 32-bit frame header, decoded.  symphonia-bundle-mp3/src/header.rs:107-233, common.rs:155-212.
struct MpaHeader {
    MpaVersion version;
    uint8_t layer;            // 1, 2 or 3
    MpaMode mode;
    uint8_t sample_rate_idx;  // 0..8: 44.1 / 48 / 32 kHz, then the halved and quartered rates (the reference's table index)
    bool mid_side;            // Layer III joint stereo
    bool intensity;           // Layer III joint stereo
    uint8_t bound;            // Layers I / II joint stereo: first sub-band coded in intensity stereo, else 32
    uint8_t emphasis;         // 0 none, 1 50/15 us, 3 CCITT J.17
    bool copyrighted, original, padding, crc;
    uint32_t bitrate;         // bit/s
    uint32_t sample_rate;     // Hz
    uint32_t frame_size;      // bytes AFTER the 4-byte header word

    int n_channels() const { return mode == MpaMode::Mono ? 1 : 2; }
    int n_granules() const { return version == MpaVersion::Mpeg1 ? 2 : 1; }
    uint32_t samples_per_frame() const { return layer == 1 ? 384u : layer == 2 ? 1152u : 576u * uint32_t(n_granules()); }
    uint32_t header_size() const { return 4u + (crc ? 2u : 0u); }
    uint32_t side_info_len() const {
        const bool mono = mode == MpaMode::Mono;
        return version == MpaVersion::Mpeg1 ? (mono ? 17u : 32u) : (mono ? 9u : 17u);
    }
};

// header.rs:71-75: eleven set bits.
inline bool mpa_is_synced(uint32_t w) { return (w & 0xffe00000u) == 0xffe00000u; }

// header.rs:49-69: the cheap plausibility test applied while hunting for sync.
inline bool mpa_check_header(uint32_t w) {
    return ((w >> 19) & 3) != 1 && ((w >> 17) & 3) != 0 && ((w >> 12) & 15) != 15 && ((w >> 10) & 3) != 3;
}

inline Status mpa_parse_header(uint32_t w, MpaHeader& h) {
    static constexpr uint16_t kbps[5][15] = {
        {0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448},  // MPEG-1 Layer I
        {0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384},     // MPEG-1 Layer II
        {0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320},      // MPEG-1 Layer III
        {0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256},     // MPEG-2 / 2.5 Layer I
        {0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160},          // MPEG-2 / 2.5 Layers II, III
    };
    static constexpr uint32_t rates[3] = {44100, 48000, 32000};
    const uint32_t v = (w >> 19) & 3, l = (w >> 17) & 3, bi = (w >> 12) & 15, ri = (w >> 10) & 3, m = (w >> 6) & 3;
    if (v == 1 || l == 0) return Status::DecodeError;
    h.version = v == 3 ? MpaVersion::Mpeg1 : v == 2 ? MpaVersion::Mpeg2 : MpaVersion::Mpeg2p5;
    h.layer = uint8_t(4 - l);
    if (bi == 0) return Status::Unsupported;  // free format
    if (bi == 15) return Status::DecodeError;
    const int row = h.version == MpaVersion::Mpeg1 ? h.layer - 1 : (h.layer == 1 ? 3 : 4);
    h.bitrate = uint32_t(kbps[row][bi]) * 1000u;
    if (ri == 3) return Status::DecodeError;
    const int shift = int(h.version);  // 0, 1, 2: full, half, quarter rate
    h.sample_rate = rates[ri] >> shift;
    h.sample_rate_idx = uint8_t(ri + 3 * shift);
    h.mode = MpaMode(m == 0 ? 0 : m == 1 ? 1 : m == 2 ? 2 : 3);
    h.mid_side = h.intensity = false;
    h.bound = 32;
    if (h.mode == MpaMode::JointStereo) {
        if (h.layer == 3) {
            h.mid_side = (w & 0x20) != 0;
            h.intensity = (w & 0x10) != 0;
        } else {
            h.bound = uint8_t((1 + ((w >> 4) & 3)) << 2);
        }
    }
    if (h.layer == 2) {  // header.rs:176-187: combinations Layer II forbids
        const uint32_t k = h.bitrate / 1000;
        if (h.mode == MpaMode::Mono ? (k == 224 || k == 256 || k == 320 || k == 384) : (k == 32 || k == 48 || k == 56 || k == 80))
            return Status::DecodeError;
    }
    const uint32_t e = w & 3;
    h.emphasis = uint8_t(e == 1 ? 1 : e == 3 ? 3 : 0);
    h.copyrighted = (w & 8) != 0;
    h.original = (w & 4) != 0;
    h.padding = (w & 0x200) != 0;
    h.crc = (w & 0x10000) == 0;
    const uint32_t factor = h.layer == 1 ? 12u : (h.layer == 3 && h.version != MpaVersion::Mpeg1) ? 72u : 144u;
    const uint32_t slots = factor * h.bitrate / h.sample_rate + (h.padding ? 1u : 0u);
    h.frame_size = slots * (h.layer == 1 ? 4u : 1u) - 4u;
    return Status::Ok;
}

// Longest frame the format can express, header included (header.rs:17).
constexpr uint32_t kMpaMaxFrameSize = 2881;

struct MpaLameInfo {
    char encoder[9];
    uint32_t delay, padding;  // samples the decoder output starts / ends with that are not audio
    uint32_t peak;            // raw 9.23 fixed-point replay-gain peak, 0 = absent
};

// Xing / Info tag of a Layer III frame (demuxer.rs:761-925).
struct MpaInfoTag {
    bool has_num_frames, has_num_bytes, has_toc, has_quality, is_cbr, has_lame;
    uint32_t num_frames, num_bytes, quality;
    MpaLameInfo lame;
};
struct MpaVbriTag {
    uint32_t num_bytes, num_mpeg_frames;
};

// demuxer.rs:942-968.  `f` = the whole frame, header word first.
inline bool mpa_is_maybe_info_tag(const uint8_t* f, size_t n, const MpaHeader& h) {
    if (h.layer != 3) return false;
    const size_t at = 4 + h.side_info_len();
    if (n < at + 8) return false;
    if (std::memcmp(f + at, "Xing", 4) != 0 && std::memcmp(f + at, "Info", 4) != 0) return false;
    for (size_t i = h.header_size(); i < at; ++i)
        if (f[i]) return false;
    return true;
}

// True when the frame is a tag the reference would act on; false for everything else, INCLUDING a tag whose
// flagged fields do not fit the frame (the reference flattens that read error to "no tag").
inline bool mpa_read_info_tag(const uint8_t* f, size_t n, const MpaHeader& h, MpaInfoTag& t) {
    if (!mpa_is_maybe_info_tag(f, n, h)) return false;
    const size_t base = 4 + h.side_info_len();
    size_t at = base;
    auto need = [&](size_t k) { return at + k <= n; };
    t = MpaInfoTag{};
    t.is_cbr = std::memcmp(f + at, "Info", 4) == 0;
    const uint32_t flags = detail::be32(f + at + 4);
    at += 8;
    if (flags & 1) {
        if (!need(4)) return false;
        t.has_num_frames = true, t.num_frames = detail::be32(f + at), at += 4;
    }
    if (flags & 2) {
        if (!need(4)) return false;
        t.has_num_bytes = true, t.num_bytes = detail::be32(f + at), at += 4;
    }
    if (flags & 4) {
        if (!need(100)) return false;
        t.has_toc = true, at += 100;
    }
    if (flags & 8) {
        if (!need(4)) return false;
        t.has_quality = true, t.quality = detail::be32(f + at), at += 4;
    }
    // LAME extension: 24 bytes up to the delay / padding field, 12 more up to its CRC-16 (over everything before it).
    if (n - at >= 24) {
        const uint8_t* e = f + at;
        MpaLameInfo li{};
        std::memcpy(li.encoder, e, 9);
        li.peak = detail::be32(e + 11);
        const uint32_t trim = detail::be24(e + 21);
        const bool known = !std::memcmp(e, "LAME", 4) || !std::memcmp(e, "Lavf", 4) || !std::memcmp(e, "Lavc", 4);
        if (known) {
            li.delay = 528 + 1 + (trim >> 12);
            const uint32_t pad = trim & 0xfff;
            li.padding = pad > 529 ? pad - 529 : 0;
        }
        at += 24;
        bool ok = true;
        if (n - at >= 12) {
            at += 10;
            if (h.crc || !std::memcmp(e, "LAME", 4)) {
                const uint32_t written = detail::be16(f + at);
                ok = written == 0 || written == crc16_ansi_le_update(0, f, at);
            }
        }
        if (ok) t.has_lame = true, t.lame = li;
    }
    return true;
}

// demuxer.rs:1023-1047, 980-1019.
inline bool mpa_is_maybe_vbri_tag(const uint8_t* f, size_t n, const MpaHeader& h) {
    if (h.layer != 3 || n < 36 + 26 || std::memcmp(f + 36, "VBRI", 4) != 0) return false;
    for (size_t i = h.header_size(); i < 36; ++i)
        if (f[i]) return false;
    return true;
}
inline bool mpa_read_vbri_tag(const uint8_t* f, size_t n, const MpaHeader& h, MpaVbriTag& t) {
    if (!mpa_is_maybe_vbri_tag(f, n, h) || detail::be16(f + 40) != 1) return false;
    t.num_bytes = detail::be32(f + 46);
    t.num_mpeg_frames = detail::be32(f + 50);
    return true;
}

// main_data_begin of a Layer III frame: how many bytes of THIS frame's main data live in earlier frames
// (demuxer.rs:664-680).  The bit-reservoir front end needs it per frame; -1 when the frame is too short.
inline int mpa_main_data_begin(const uint8_t* f, size_t n, const MpaHeader& h) {
    const size_t at = h.header_size();
    if (h.version == MpaVersion::Mpeg1) return at + 2 <= n ? int(detail::be16(f + at) >> 7) : -1;
    return at + 1 <= n ? int(f[at]) : -1;
}

// One packet = one frame, a byte range of the source.
struct MpaPacket {
    uint64_t offset;      // of the header word
    uint32_t size;        // header word included
    uint32_t header;      // the header word
    int64_t pts;          // in samples; starts at -delay
    uint32_t dur;         // samples the frame decodes to
    uint32_t trim_start;  // leading samples to drop (encoder delay)
    uint64_t trim_end;    // trailing samples to drop; NOT capped to dur, as in the reference (packet.rs:334-338)
};

struct MpaTrack {
    MpaHeader first;           // codec parameters come from the first frame
    uint32_t first_word;
    bool has_delay, has_num_frames;
    uint32_t delay, padding;
    uint64_t num_frames;       // samples of audio, delay and padding removed (exact, from a tag) or estimated
    enum Tag : uint8_t { None, Xing, Info, Vbri } tag;
    uint64_t first_packet_pos;
};

// The reference's MpaReader over a resident buffer: open() = try_new, next() = next_packet.
class MpaIndexer {
  public:
    MpaIndexer(const uint8_t* data, size_t n) : d_(data), n_(n) {}

    // demuxer.rs:414-487.  EndOfStream: the buffer holds no frame.  `seekable` = the reference's is_seekable():
    // without it no duration is estimated for an untagged stream and nothing is trimmed from its end.
    Status open(bool seekable = true) {
        size_t at = 0, size = 0;
        MpaHeader h;
        uint32_t w;
        // demuxer.rs:610-640: accept a first frame only when the next word looks like the same kind of stream;
        // rejected candidates restart the hunt one byte further.
        for (size_t from = 0;;) {
            if (!find_frame(from, at, w, h)) return Status::EndOfStream;
            size = 4 + size_t(h.frame_size);
            if (at + size + 4 <= n_) {
                const uint32_t nxt = detail::be32(d_ + at + size);
                MpaHeader c;
                const bool similar = mpa_is_synced(nxt) && mpa_parse_header(nxt, c) == Status::Ok && c.version == h.version &&
                                     c.layer == h.layer && c.sample_rate == h.sample_rate && c.n_channels() == h.n_channels();
                if (!similar) {
                    from = at + 1;
                    continue;
                }
            }
            break;
        }
        track_ = MpaTrack{};
        track_.first = h;
        track_.first_word = w;
        track_.tag = MpaTrack::None;
        pos_ = at + size;
        MpaInfoTag info;
        MpaVbriTag vbri;
        if (mpa_read_info_tag(d_ + at, size, h, info)) {
            track_.tag = info.is_cbr ? MpaTrack::Info : MpaTrack::Xing;
            if (info.has_lame) track_.has_delay = true, track_.delay = info.lame.delay, track_.padding = info.lame.padding;
            if (info.has_num_frames) {
                const uint64_t total = uint64_t(info.num_frames) * h.samples_per_frame();
                const uint64_t cut = uint64_t(track_.delay) + track_.padding;
                track_.has_num_frames = true, track_.num_frames = total > cut ? total - cut : 0;
            }
        } else if (mpa_read_vbri_tag(d_ + at, size, h, vbri)) {
            track_.tag = MpaTrack::Vbri;
            track_.has_num_frames = true, track_.num_frames = uint64_t(vbri.num_mpeg_frames) * h.samples_per_frame();
        } else {
            pos_ = at;  // an ordinary frame: it is the first packet
            uint64_t frames;
            if (seekable && estimate_frames(at, frames)) track_.has_num_frames = true, track_.num_frames = frames * h.samples_per_frame();
        }
        track_.first_packet_pos = pos_;
        ts_ = -int64_t(track_.delay);
        open_ = true;
        return Status::Ok;
    }

    const MpaTrack& track() const { return track_; }

    // demuxer.rs:160-218.  Frames that are Xing / Info / VBRI tags are dropped wherever they turn up.
    Status next(MpaPacket& p) {
        if (!open_) return Status::DecodeError;
        for (;;) {
            size_t at;
            MpaHeader h;
            uint32_t w;
            if (!find_frame(pos_, at, w, h)) return Status::EndOfStream;
            const size_t size = 4 + size_t(h.frame_size);
            pos_ = at + size;
            MpaInfoTag info;
            MpaVbriTag vbri;
            if (mpa_is_maybe_info_tag(d_ + at, size, h)) {
                if (mpa_read_info_tag(d_ + at, size, h, info)) continue;
            } else if (mpa_read_vbri_tag(d_ + at, size, h, vbri)) {
                continue;
            }
            const uint32_t dur = h.samples_per_frame();
            p.offset = at, p.size = uint32_t(size), p.header = w, p.pts = ts_, p.dur = dur;
            p.trim_start = ts_ < 0 ? uint32_t(-ts_ < int64_t(dur) ? -ts_ : int64_t(dur)) : 0u;
            p.trim_end = 0;
            if (track_.has_num_frames) {
                const int64_t over = ts_ + int64_t(dur) - int64_t(track_.num_frames);
                if (over > 0) p.trim_end = uint64_t(over);
            }
            ts_ += dur;
            return Status::Ok;
        }
    }

    // Everything at once.
    static Status index(const uint8_t* data, size_t n, MpaTrack& track, std::vector<MpaPacket>& out, bool seekable = true) {
        MpaIndexer ix(data, n);
        const Status s = ix.open(seekable);
        if (s != Status::Ok) return s;
        track = ix.track();
        MpaPacket p;
        while (ix.next(p) == Status::Ok) out.push_back(p);
        return Status::Ok;
    }

  private:
    // header.rs:77-103 + demuxer.rs:585-607 as a window search: the first offset >= from whose word has the sync
    // bits and passes the plausibility test; a word that then fails the full parse (free format, a forbidden Layer II
    // combination) costs its 4 bytes, and a frame whose body runs past the buffer ends the stream.
    bool find_frame(size_t from, size_t& at, uint32_t& w, MpaHeader& h) const {
        for (size_t q = from; q + 4 <= n_;) {
            // cheap reject on the first byte keeps the scan at memchr speed through payload bytes
            if (d_[q] != 0xff) {
                const void* hit = std::memchr(d_ + q, 0xff, n_ - q);
                if (!hit) return false;
                q = size_t(static_cast<const uint8_t*>(hit) - d_);
                if (q + 4 > n_) return false;
            }
            const uint32_t word = detail::be32(d_ + q);
            if (!mpa_is_synced(word) || !mpa_check_header(word)) {
                ++q;
                continue;
            }
            if (mpa_parse_header(word, h) != Status::Ok) {
                q += 4;
                continue;
            }
            if (q + 4 + size_t(h.frame_size) > n_) return false;
            at = q, w = word;
            return true;
        }
        return false;
    }

    // demuxer.rs:683-733: average the first frames (more than 16 of them or more than 16 KiB) and extrapolate.
    bool estimate_frames(size_t from, uint64_t& frames) const {
        const double total_len = double(n_ - from);
        size_t q = from, len = 0;
        unsigned count = 0;
        for (;;) {
            MpaHeader h;
            if (q + 4 > n_ || mpa_parse_header(detail::be32(d_ + q), h) != Status::Ok) return false;
            len += 4 + h.frame_size, ++count;
            if (q + 4 + h.frame_size > n_) return false;
            q += 4 + h.frame_size;
            if (count > 16 || len > 16 * 1024) break;
        }
        frames = uint64_t(total_len / (double(len) / double(count)));
        return true;
    }

    const uint8_t* d_;
    size_t n_;
    size_t pos_ = 0;
    int64_t ts_ = 0;
    bool open_ = false;
    MpaTrack track_{};
};

// =====================================================================================================================
// ADTS (AAC)
// =====================================================================================================================

struct AdtsHeader {
    uint8_t profile;       // MPEG-4 audio object type: 1 Main, 2 LC, 3 SSR, 4 LTP
    uint8_t channels;      // 0: configured in-band (program config element)
    uint8_t header_len;    // 7, or 9 with a CRC
    bool has_crc;
    uint16_t crc;
    uint16_t frame_len;    // sync word, header and payload
    uint32_t sample_rate;
    uint32_t payload_len() const { return uint32_t(frame_len) - header_len; }
};

// adts.rs:202-204: 0xfff, layer bits zero; the MPEG-2/4 bit and the protection bit are free.
inline bool adts_is_sync(uint32_t w16) { return (w16 & 0xfff6u) == 0xfff0u; }

// adts.rs:137-198.  `p` points at the sync word; `n` bytes are readable.  EndOfStream: fewer than header_len bytes.
inline Status adts_parse_header(const uint8_t* p, size_t n, AdtsHeader& h) {
    static constexpr uint32_t rates[13] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350};
    static constexpr uint8_t chans[8] = {0, 1, 2, 3, 4, 5, 6, 8};
    if (n < 2) return Status::EndOfStream;
    h.has_crc = (p[1] & 1) == 0;
    h.header_len = h.has_crc ? 9 : 7;
    if (n < h.header_len) return Status::EndOfStream;
    h.profile = uint8_t((p[2] >> 6) + 1);
    const uint32_t ri = (p[2] >> 2) & 15;
    if (ri > 12) return Status::DecodeError;  // 15 is the escape ADTS forbids, 13 / 14 are reserved
    h.sample_rate = rates[ri];
    h.channels = chans[((p[2] & 1) << 2) | (p[3] >> 6)];
    h.frame_len = uint16_t(((p[3] & 3) << 11) | (p[4] << 3) | (p[5] >> 5));
    if (h.frame_len < h.header_len) return Status::DecodeError;
    if ((p[6] & 3) != 0) return Status::Unsupported;  // more than one raw data block per frame
    h.crc = h.has_crc ? uint16_t(detail::be16(p + 7)) : 0;
    return Status::Ok;
}

struct AdtsPacket {
    uint64_t offset;  // of the PAYLOAD (the raw data block), as the reference's packets carry no ADTS header
    uint32_t size;
    int64_t pts;      // 1024 samples per packet
    uint32_t sample_rate;
    uint8_t channels, profile;
};

// adts.rs:278-309 over a resident buffer.
class AdtsIndexer {
  public:
    AdtsIndexer(const uint8_t* data, size_t n) : d_(data), n_(n) {}

    // EndOfStream at the end of the bytes (truncated() tells a clean end from a cut payload); DecodeError /
    // Unsupported leave the cursor behind the offending header, as the reference's reader does, so that a caller
    // that chooses to carry on resynchronises from there.
    Status next(AdtsPacket& p) {
        size_t q = pos_;
        for (;; ++q) {
            if (q + 2 > n_) return pos_ = n_, Status::EndOfStream;
            if (d_[q] != 0xff) {
                const void* hit = std::memchr(d_ + q, 0xff, n_ - q);
                if (!hit) return pos_ = n_, Status::EndOfStream;
                q = size_t(static_cast<const uint8_t*>(hit) - d_);
                if (q + 2 > n_) return pos_ = n_, Status::EndOfStream;
            }
            if (adts_is_sync(detail::be16(d_ + q))) break;
        }
        AdtsHeader h;
        const Status s = adts_parse_header(d_ + q, n_ - q, h);
        if (s == Status::EndOfStream) return pos_ = n_, s;
        pos_ = q + h.header_len;
        if (s != Status::Ok) return s;
        if (pos_ + h.payload_len() > n_) return truncated_ = true, pos_ = n_, Status::EndOfStream;
        p.offset = pos_, p.size = h.payload_len(), p.pts = ts_, p.sample_rate = h.sample_rate, p.channels = h.channels, p.profile = h.profile;
        pos_ += h.payload_len();
        ts_ += 1024;
        return Status::Ok;
    }
    bool truncated() const { return truncated_; }

    // Up to the first non-Ok status, which is returned.
    static Status index(const uint8_t* data, size_t n, std::vector<AdtsPacket>& out, bool* truncated = nullptr) {
        AdtsIndexer ix(data, n);
        AdtsPacket p;
        Status s;
        while ((s = ix.next(p)) == Status::Ok) out.push_back(p);
        if (truncated) *truncated = ix.truncated();
        return s;
    }

  private:
    const uint8_t* d_;
    size_t n_;
    size_t pos_ = 0;
    int64_t ts_ = 0;
    bool truncated_ = false;
};

// =====================================================================================================================
// Ogg
// =====================================================================================================================

constexpr size_t kOggHeaderSize = 27;
constexpr size_t kOggMaxPageSize = kOggHeaderSize + 255 + 255 * 255;  // page.rs:17
constexpr uint64_t kOggMaxPacketLen = 16u * 1024 * 1024;              // logical.rs:61

struct OggPage {
    uint64_t offset;       // of the capture pattern
    uint64_t absgp;        // granule position
    uint32_t serial, sequence, crc;
    uint8_t n_segments;
    bool continuation, first, last;
    uint64_t body_offset;  // first body byte
    uint32_t body_len;
    uint16_t n_packets;    // packets that END on this page
    uint16_t packet_len[255];
    uint32_t partial_len() const {  // body bytes after the last packet end: a packet continued on a later page
        uint32_t used = 0;
        for (unsigned i = 0; i < n_packets; ++i) used += packet_len[i];
        return body_len - used;
    }
};

// page.rs:166-271 over a resident buffer.
class OggPageReader {
  public:
    OggPageReader(const uint8_t* data, size_t n) : d_(data), n_(n) {}

    // One attempt (try_next_page): DecodeError for a bad version / flag byte (the search resumes after that header)
    // or a checksum mismatch (the search resumes right after the capture pattern that led here).
    Status try_next(OggPage& pg) {
        size_t q = pos_;
        for (;; ++q) {
            if (q + 4 > n_) return pos_ = n_, Status::EndOfStream;
            if (d_[q] != 'O') {
                const void* hit = std::memchr(d_ + q, 'O', n_ - q);
                if (!hit) return pos_ = n_, Status::EndOfStream;
                q = size_t(static_cast<const uint8_t*>(hit) - d_);
                if (q + 4 > n_) return pos_ = n_, Status::EndOfStream;
            }
            if (std::memcmp(d_ + q, "OggS", 4) == 0) break;
        }
        if (q + kOggHeaderSize > n_) return pos_ = n_, Status::EndOfStream;
        const uint8_t* h = d_ + q;
        pos_ = q + kOggHeaderSize;
        if (h[4] != 0 || (h[5] & 0xf8)) return Status::DecodeError;
        pg.offset = q;
        pg.continuation = h[5] & 1, pg.first = (h[5] & 2) != 0, pg.last = (h[5] & 4) != 0;
        pg.absgp = detail::le64(h + 6);
        pg.serial = detail::le32(h + 14), pg.sequence = detail::le32(h + 18), pg.crc = detail::le32(h + 22);
        pg.n_segments = h[26];
        if (pos_ + pg.n_segments > n_) return pos_ = n_, Status::EndOfStream;
        const uint8_t* lacing = d_ + pos_;
        uint32_t body = 0, run = 0;
        pg.n_packets = 0;
        for (unsigned i = 0; i < pg.n_segments; ++i) {
            body += lacing[i], run += lacing[i];
            if (lacing[i] < 255) pg.packet_len[pg.n_packets++] = uint16_t(run), run = 0;  // a short segment closes a packet
        }
        pos_ += pg.n_segments;
        if (pos_ + body > n_) return pos_ = n_, Status::EndOfStream;
        pg.body_offset = pos_, pg.body_len = body;
        // checksum over the page with its own checksum field read as zero
        static const uint8_t zero[4] = {0, 0, 0, 0};
        uint32_t crc = crc32_update(0, h, 22);
        crc = crc32_update(crc, zero, 4);
        crc = crc32_update(crc, h + 26, 1 + size_t(pg.n_segments) + body);
        if (crc != pg.crc) {
            pos_ = q + 4;
            return Status::DecodeError;
        }
        pos_ += body;
        return Status::Ok;
    }

    // next_page: skip whatever does not verify.
    Status next(OggPage& pg) {
        for (;;) {
            const Status s = try_next(pg);
            if (s == Status::Ok || s == Status::EndOfStream) return s;
            ++n_rejected_;
        }
    }
    size_t position() const { return pos_; }
    size_t rejected() const { return n_rejected_; }

  private:
    const uint8_t* d_;
    size_t n_;
    size_t pos_ = 0;
    size_t n_rejected_ = 0;
};

// A packet of a logical stream: pieces [first_piece, first_piece + n_pieces) of the stream's piece list.
struct OggPacket {
    uint32_t first_piece;
    uint32_t n_pieces;
    uint64_t len;
    uint32_t page_sequence;  // of the page it ended on
    uint64_t page_absgp;     // granule position of that page: the end time of its LAST completed packet
    bool last_on_page;
};

// logical.rs:104-205, 577-620 without the codec mapper: reassembles the packets of one serial number.
class OggLogicalStream {
  public:
    // DecodeError when an open packet would pass the reference's 16 MiB cap; the page's completed packets are kept.
    Status read_page(const OggPage& pg) {
        if (have_prev_ && (pg.sequence < prev_seq_ || pg.sequence - prev_seq_ > 1)) drop_partial();  // lost or re-ordered pages
        have_prev_ = true, prev_seq_ = pg.sequence;
        if (!pg.continuation && part_len_ > 0) drop_partial();  // the continuation never came
        unsigned i = 0;
        uint64_t at = pg.body_offset;
        if (pg.continuation && part_len_ == 0) {
            // the head of this packet was never seen: drop its tail, or the whole page when nothing else ends here
            if (pg.n_packets == 0) return Status::Ok;
            at += pg.packet_len[i++];
        }
        const size_t before = packets_.size();
        for (; i < pg.n_packets; ++i) {
            const uint32_t n = pg.packet_len[i];
            pieces_.push_back(Piece{at, n});
            OggPacket p;
            p.first_piece = part_first_, p.n_pieces = uint32_t(pieces_.size()) - part_first_, p.len = part_len_ + n;
            p.page_sequence = pg.sequence, p.page_absgp = pg.absgp, p.last_on_page = false;
            packets_.push_back(p);
            part_first_ = uint32_t(pieces_.size()), part_len_ = 0;
            at += n;
        }
        if (packets_.size() > before) packets_.back().last_on_page = true;
        const uint64_t rest = pg.body_offset + pg.body_len - at;
        if (rest > 0) {
            if (part_len_ + rest > kOggMaxPacketLen) return Status::DecodeError;
            pieces_.push_back(Piece{at, uint32_t(rest)});
            part_len_ += rest;
        }
        return Status::Ok;
    }

    const std::vector<OggPacket>& packets() const { return packets_; }
    const std::vector<Piece>& pieces() const { return pieces_; }
    uint64_t open_len() const { return part_len_; }  // bytes of a packet still waiting for its continuation

    // Copy a packet out of the source buffer (tests, host-side header parsing); the device path gathers instead.
    void gather(const uint8_t* src, const OggPacket& p, uint8_t* dst) const {
        for (uint32_t k = 0; k < p.n_pieces; ++k) {
            const Piece& pc = pieces_[p.first_piece + k];
            std::memcpy(dst, src + pc.offset, pc.len);
            dst += pc.len;
        }
    }

  private:
    void drop_partial() {
        pieces_.resize(part_first_);
        part_len_ = 0;
    }
    std::vector<OggPacket> packets_;
    std::vector<Piece> pieces_;
    uint32_t part_first_ = 0;  // pieces_[part_first_..] belong to the packet still open
    uint64_t part_len_ = 0;
    bool have_prev_ = false;
    uint32_t prev_seq_ = 0;
};

// A physical stream: every page that verifies, routed by serial number.  A logical stream exists from its
// beginning-of-stream page on (demuxer.rs:320-345); pages of serials never announced are counted and skipped.
struct OggIndex {
    std::vector<OggPage> pages;
    std::map<uint32_t, OggLogicalStream> streams;
    size_t rejected = 0;  // capture patterns that did not lead to a valid page
    size_t orphans = 0;   // valid pages of unannounced serials

    static Status build(const uint8_t* data, size_t n, OggIndex& ix, bool keep_pages = true) {
        OggPageReader rd(data, n);
        OggPage pg;
        Status worst = Status::Ok;
        while (rd.next(pg) == Status::Ok) {
            if (keep_pages) ix.pages.push_back(pg);
            auto it = ix.streams.find(pg.serial);
            if (it == ix.streams.end()) {
                if (!pg.first) {
                    ++ix.orphans;
                    continue;
                }
                it = ix.streams.emplace(pg.serial, OggLogicalStream{}).first;
            }
            if (it->second.read_page(pg) != Status::Ok) worst = Status::DecodeError;
        }
        ix.rejected = rd.rejected();
        return worst;
    }
};

// =====================================================================================================================
// Vorbis in Ogg: what the container layer has to know about the codec
// =====================================================================================================================
// The mapper (symphonia-format-ogg/src/mappings/vorbis.rs) sorts a logical stream's packets into identification /
// comment / setup headers and audio, hands the decoder its `extra_data` (identification packet followed by the
// setup packet, consumed at symphonia-codec-vorbis/src/lib.rs:75-89) and derives every audio packet's duration
// from its first bits: the mode number selects a short or long block, and a packet yields a quarter of the
// previous block plus a quarter of its own.  For that it must walk the WHOLE setup header -- codebooks, floors,
// residues, mappings -- just to reach the mode list at its end.

// Bits least-significant first, as Vorbis packs them (symphonia-core/src/io/bit.rs:941-1027).
class BitReaderRtl {
  public:
    BitReaderRtl(const uint8_t* p, size_t n) : p_(p), n_bits_(uint64_t(n) * 8) {}
    bool ok() const { return ok_; }
    uint64_t bits_left() const { return n_bits_ - at_; }
    // Past the end: returns 0 and latches !ok() (every caller checks once per structure, not per field).
    uint32_t read(unsigned width) {
        if (width > bits_left()) return ok_ = false, at_ = n_bits_, 0u;
        uint64_t v = 0;
        const uint64_t byte = at_ >> 3;
        const unsigned shift = unsigned(at_ & 7), need = (shift + width + 7) >> 3;
        for (unsigned k = 0; k < need; ++k) v |= uint64_t(p_[byte + k]) << (8 * k);
        at_ += width;
        return uint32_t((v >> shift) & ((uint64_t(1) << width) - 1));
    }
    bool read_bool() { return read(1) != 0; }
    void ignore(uint64_t width) {
        if (width > bits_left()) ok_ = false, at_ = n_bits_;
        else at_ += width;
    }

  private:
    const uint8_t* p_;
    uint64_t n_bits_, at_ = 0;
    bool ok_ = true;
};

inline uint32_t vorbis_ilog(uint32_t x) {
    uint32_t n = 0;
    for (; x; x >>= 1) ++n;
    return n;
}

struct VorbisIdent {
    uint8_t n_channels;
    uint32_t sample_rate;
    uint8_t bs0_exp, bs1_exp;  // block sizes as powers of two, 6..13, short <= long
};

// mappings/vorbis.rs:293-360.  The Ogg mapper only accepts an identification packet of exactly 30 bytes (:113-117).
inline Status vorbis_read_ident(const uint8_t* p, size_t n, VorbisIdent& id) {
    if (n < 30) return Status::EndOfStream;
    if (p[0] != 1 || std::memcmp(p + 1, "vorbis", 6) != 0) return Status::DecodeError;
    if (detail::le32(p + 7) != 0) return Status::Unsupported;
    id.n_channels = p[11];
    id.sample_rate = detail::le32(p + 12);
    if (id.n_channels == 0 || id.sample_rate == 0) return Status::DecodeError;
    id.bs0_exp = p[28] & 15, id.bs1_exp = p[28] >> 4;
    if (id.bs0_exp < 6 || id.bs0_exp > 13 || id.bs1_exp < 6 || id.bs1_exp > 13 || id.bs0_exp > id.bs1_exp) return Status::DecodeError;
    if (p[29] != 1) return Status::DecodeError;  // framing
    return Status::Ok;
}

namespace detail {
// The largest v with v^dims <= entries (the reference computes it in f32 and asserts exactly this, :717-730).
inline uint32_t vorbis_lookup1_values(uint32_t entries, uint32_t dims) {
    if (dims == 1) return entries;  // (untrusted input: keep the search short -- for dims >= 2 the root of 2^24 is <= 4096)
    uint32_t v = 0;
    for (;;) {
        uint64_t pw = 1;
        bool over = false;
        for (uint32_t k = 0; k < dims && !over; ++k) pw *= uint64_t(v) + 1, over = pw > entries;
        if (over) return v;
        ++v;
    }
}

// mappings/vorbis.rs:426-500.
inline bool vorbis_skip_codebook(BitReaderRtl& bs) {
    if (bs.read(24) != 0x564342 || !bs.ok()) return false;
    const uint32_t dims = bs.read(16), entries = bs.read(24);
    if (!bs.read_bool()) {        // lengths in entry order
        if (bs.read_bool()) {     // sparse: a used flag in front of every length
            for (uint32_t i = 0; i < entries && bs.ok(); ++i)
                if (bs.read_bool()) bs.read(5);
        } else {
            bs.ignore(uint64_t(entries) * 5);
        }
    } else {                      // lengths as run lengths of ascending code length
        bs.read(5);
        for (uint32_t cur = 0;;) {
            cur += bs.read(entries > cur ? vorbis_ilog(entries - cur) : 0);
            if (!bs.ok() || cur > entries) return false;
            if (cur == entries) break;
        }
    }
    const uint32_t lookup = bs.read(4);
    if (!bs.ok()) return false;
    if (lookup == 0) return true;
    if (lookup > 2) return false;
    bs.ignore(64);
    const uint32_t value_bits = bs.read(4) + 1;
    bs.read_bool();
    if (!bs.ok()) return false;
    if (lookup == 1 && dims == 0) return false;  // the reference's float root is meaningless here (it asserts)
    const uint64_t values = lookup == 1 ? vorbis_lookup1_values(entries, dims) : uint64_t(entries) * dims;
    bs.ignore(values * value_bits);
    return bs.ok();
}

// :528-590
inline bool vorbis_skip_floor(BitReaderRtl& bs) {
    const uint32_t type = bs.read(16);
    if (!bs.ok() || type > 1) return false;
    if (type == 0) {
        bs.ignore(8 + 16 + 16 + 6 + 8);
        bs.ignore((uint64_t(bs.read(4)) + 1) * 8);
        return bs.ok();
    }
    const uint32_t partitions = bs.read(5);
    uint8_t cls[32] = {0}, dims[16] = {0};
    if (partitions > 0) {
        uint32_t max_class = 0;
        for (uint32_t i = 0; i < partitions; ++i) {
            cls[i] = uint8_t(bs.read(4));
            if (cls[i] > max_class) max_class = cls[i];
        }
        for (uint32_t c = 0; c <= max_class; ++c) {
            dims[c] = uint8_t(bs.read(3) + 1);
            const uint32_t sub = bs.read(2);
            if (sub) bs.read(8);
            bs.ignore((uint64_t(1) << sub) * 8);
        }
    }
    bs.read(2);
    const uint32_t rangebits = bs.read(4);
    for (uint32_t i = 0; i < partitions; ++i) bs.ignore(uint64_t(dims[cls[i]]) * rangebits);
    return bs.ok();
}

// :592-620
inline bool vorbis_skip_residue(BitReaderRtl& bs) {
    bs.read(16);
    bs.ignore(24 + 24 + 24);
    const uint32_t classes = bs.read(6) + 1;
    bs.ignore(8);
    uint32_t books = 0;
    for (uint32_t i = 0; i < classes && bs.ok(); ++i) {
        uint32_t used = bs.read(3);
        if (bs.read_bool()) used |= bs.read(5) << 3;
        for (; used; used &= used - 1) ++books;
    }
    bs.ignore(uint64_t(books) * 8);
    return bs.ok();
}

// :622-673
inline bool vorbis_skip_mapping(BitReaderRtl& bs, uint8_t channels) {
    if (bs.read(16) != 0 || !bs.ok()) return false;
    const uint32_t submaps = bs.read_bool() ? bs.read(4) + 1 : 1;
    if (bs.read_bool()) {
        const uint32_t steps = bs.read(8) + 1, width = vorbis_ilog(uint32_t(channels) - 1);
        bs.ignore(uint64_t(steps) * 2 * width);
    }
    if (bs.read(2) != 0 || !bs.ok()) return false;
    if (submaps > 1) bs.ignore(uint64_t(channels) * 4);
    bs.ignore(uint64_t(submaps) * 24);
    return bs.ok();
}
}  // namespace detail

// mappings/vorbis.rs:362-405, 675-708: the mode list of a setup packet, as a bit mask of "long block" flags.
inline Status vorbis_read_setup_modes(const uint8_t* p, size_t n, const VorbisIdent& id, uint8_t& num_modes, uint64_t& long_block_mask) {
    if (n < 7) return Status::EndOfStream;
    if (p[0] != 5 || std::memcmp(p + 1, "vorbis", 6) != 0) return Status::DecodeError;
    BitReaderRtl bs(p + 7, n - 7);
    for (uint32_t i = 0, count = bs.read(8) + 1; i < count; ++i)
        if (!detail::vorbis_skip_codebook(bs)) return Status::DecodeError;
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i)
        if (bs.read(16) != 0 || !bs.ok()) return Status::DecodeError;  // time-domain transforms: placeholders
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i)
        if (!detail::vorbis_skip_floor(bs)) return Status::DecodeError;
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i)
        if (!detail::vorbis_skip_residue(bs)) return Status::DecodeError;
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i)
        if (!detail::vorbis_skip_mapping(bs, id.n_channels)) return Status::DecodeError;
    const uint32_t count = bs.read(6) + 1;
    uint64_t mask = 0;
    for (uint32_t i = 0; i < count; ++i) {
        if (bs.read_bool()) mask |= uint64_t(1) << i;
        const uint32_t window = bs.read(16), transform = bs.read(16);
        bs.read(8);
        if (!bs.ok() || window != 0 || transform != 0) return Status::DecodeError;
    }
    if (!bs.read_bool() || !bs.ok()) return Status::DecodeError;  // framing
    num_modes = uint8_t(count), long_block_mask = mask;
    return Status::Ok;
}

// ---- the decoder's view of the setup header ---------------------------------------------------------------------------
// symphonia-codec-vorbis/src/lib.rs:490-770 (read_setup: floors, residues, mappings, modes with their cross checks),
// floor.rs:160-201 (floor 0), :455-560 (floor 1: classes, X list, neighbours, sort order), residue.rs:73-140.
// Everything the synthesis configuration needs EXCEPT the codebooks' contents, which are walked over with the same
// syntax checks as above but not built (their Huffman trees and VQ tables belong to the packet decoder).
struct VorbisFloor1Setup {
    uint8_t multiplier;  // 1..4
    uint8_t n_posts;     // 2..65
    uint16_t x_list[65];
    uint8_t low[65], high[65];   // neighbours among the earlier posts (0, 0 for the first two, as find_neighbors leaves them)
    uint8_t sort_order[65];      // posts by ascending x (stable)
    uint8_t partitions;
    uint8_t partition_class[32];
    struct Class {
        uint8_t dimensions, subclass_bits, mainbook, subbook_used;
        uint8_t subbooks[8];
    } classes[16];
};
struct VorbisResidueSetup {
    uint16_t type;
    uint32_t begin, end, partition_size;
    uint8_t classifications, classbook, max_pass;
    uint8_t used[64];
    uint8_t books[64][8];
};
struct VorbisMappingSetup {
    uint8_t n_submaps;
    std::vector<std::pair<uint8_t, uint8_t>> couplings;  // (magnitude channel, angle channel)
    std::vector<uint8_t> multiplex;                      // sub-map of each channel
    uint8_t submap_floor[16], submap_residue[16];
};
struct VorbisSetup {
    uint32_t n_codebooks = 0;
    std::vector<uint8_t> floor_type;          // 0 or 1 per floor
    std::vector<VorbisFloor1Setup> floor1;    // per floor; zeroed for type-0 floors
    std::vector<VorbisResidueSetup> residues;
    std::vector<VorbisMappingSetup> mappings;
    std::vector<std::pair<bool, uint8_t>> modes;  // (long block, mapping)
};

inline Status vorbis_read_setup(const uint8_t* p, size_t n, const VorbisIdent& id, VorbisSetup& out) {
    if (n < 7) return Status::EndOfStream;
    if (p[0] != 5 || std::memcmp(p + 1, "vorbis", 6) != 0) return Status::DecodeError;
    BitReaderRtl bs(p + 7, n - 7);
    out = VorbisSetup{};
    out.n_codebooks = bs.read(8) + 1;
    for (uint32_t i = 0; i < out.n_codebooks; ++i)
        if (!detail::vorbis_skip_codebook(bs)) return Status::DecodeError;
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i)
        if (bs.read(16) != 0 || !bs.ok()) return Status::DecodeError;
    const uint8_t max_book = uint8_t(out.n_codebooks);  // `codebooks.len() as u8` (lib.rs:517): 256 codebooks wrap to 0
    // ---- floors
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i) {
        const uint32_t type = bs.read(16);
        if (!bs.ok() || type > 1) return Status::DecodeError;
        VorbisFloor1Setup f{};
        if (type == 0) {
            bs.ignore(8 + 16 + 16 + 6 + 8);
            for (uint32_t k = 0, books = bs.read(4) + 1; k < books; ++k)
                if (bs.read(8) >= max_book || !bs.ok()) return Status::DecodeError;
        } else {
            f.partitions = uint8_t(bs.read(5));
            if (f.partitions) {
                uint8_t max_class = 0;
                for (int k = 0; k < f.partitions; ++k) f.partition_class[k] = uint8_t(bs.read(4)), max_class = std::max(max_class, f.partition_class[k]);
                for (int c = 0; c <= max_class; ++c) {
                    auto& cl = f.classes[c];
                    cl.dimensions = uint8_t(bs.read(3) + 1), cl.subclass_bits = uint8_t(bs.read(2));
                    if (cl.subclass_bits) {
                        cl.mainbook = uint8_t(bs.read(8));
                        if (cl.mainbook >= max_book) return Status::DecodeError;
                    }
                    for (int k = 0; k < (1 << cl.subclass_bits); ++k) {
                        uint8_t book = uint8_t(bs.read(8));
                        if (book > 0) {  // 0 = no codebook for this sub-class; otherwise the number minus one
                            if (--book >= max_book) return Status::DecodeError;
                            cl.subbook_used |= uint8_t(1 << k);
                        }
                        cl.subbooks[k] = book;
                    }
                }
            }
            f.multiplier = uint8_t(bs.read(2) + 1);
            const uint32_t rangebits = bs.read(4);
            if (!bs.ok()) return Status::DecodeError;
            int np = 0;
            f.x_list[np++] = 0, f.x_list[np++] = uint16_t(1u << rangebits);
            for (int k = 0; k < f.partitions; ++k) {
                const int dims = f.classes[f.partition_class[k]].dimensions;
                if (np + dims > 65) return Status::DecodeError;
                for (int d = 0; d < dims; ++d) {
                    const uint32_t x = bs.read(rangebits);
                    // every READ element must be new among the read ones; the two implied ends are not in that set
                    // (floor.rs:519-536 inserts only what it reads)
                    for (int e = 2; e < np; ++e)
                        if (f.x_list[e] == x) return Status::DecodeError;
                    f.x_list[np++] = uint16_t(x);
                }
            }
            if (!bs.ok()) return Status::DecodeError;
            f.n_posts = uint8_t(np);
            for (int i2 = 0; i2 < np; ++i2) {  // floor.rs:748-773
                uint32_t lo = 0, hi = 0xffffffffu;
                for (int e = 0; e < i2; ++e) {
                    const uint32_t xv = f.x_list[e];
                    if (xv > lo && xv < f.x_list[i2]) lo = xv, f.low[i2] = uint8_t(e);
                    if (xv < hi && xv > f.x_list[i2]) hi = xv, f.high[i2] = uint8_t(e);
                }
                f.sort_order[i2] = uint8_t(i2);
            }
            std::stable_sort(f.sort_order, f.sort_order + np, [&](uint8_t a, uint8_t b) { return f.x_list[a] < f.x_list[b]; });
        }
        out.floor_type.push_back(uint8_t(type)), out.floor1.push_back(f);
    }
    // ---- residues
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i) {
        VorbisResidueSetup r{};
        r.type = uint16_t(bs.read(16));
        if (!bs.ok() || r.type > 2) return Status::DecodeError;
        r.begin = bs.read(24), r.end = bs.read(24), r.partition_size = bs.read(24) + 1;
        r.classifications = uint8_t(bs.read(6) + 1), r.classbook = uint8_t(bs.read(8));
        if (!bs.ok() || r.classbook >= max_book || r.end < r.begin) return Status::DecodeError;
        for (int c = 0; c < r.classifications; ++c) {
            const uint32_t low = bs.read(3);
            r.used[c] = uint8_t((bs.read_bool() ? bs.read(5) << 3 : 0) | low);
        }
        for (int c = 0; c < r.classifications; ++c)
            for (int j = 0; j < 8; ++j)
                if (r.used[c] & (1 << j)) {
                    r.books[c][j] = uint8_t(bs.read(8));
                    if (!bs.ok() || r.books[c][j] == 0 || r.books[c][j] >= max_book) return Status::DecodeError;
                    r.max_pass = std::max<uint8_t>(r.max_pass, uint8_t(j));
                }
        if (!bs.ok()) return Status::DecodeError;
        out.residues.push_back(r);
    }
    // ---- mappings
    const uint8_t max_floor = uint8_t(out.floor_type.size()), max_residue = uint8_t(out.residues.size());
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i) {
        if (bs.read(16) != 0 || !bs.ok()) return Status::DecodeError;
        VorbisMappingSetup m{};
        m.n_submaps = uint8_t(bs.read_bool() ? bs.read(4) + 1 : 1);
        if (bs.read_bool()) {
            const uint32_t steps = bs.read(8) + 1, width = vorbis_ilog(uint32_t(id.n_channels) - 1), max_ch = uint32_t(id.n_channels) - 1;
            for (uint32_t k = 0; k < steps; ++k) {
                const uint32_t mag = bs.read(width) & 0xff, ang = bs.read(width) & 0xff;
                if (!bs.ok() || mag == ang || mag > max_ch || ang > max_ch) return Status::DecodeError;
                m.couplings.emplace_back(uint8_t(mag), uint8_t(ang));
            }
        }
        if (bs.read(2) != 0 || !bs.ok()) return Status::DecodeError;
        m.multiplex.assign(id.n_channels, 0);
        if (m.n_submaps > 1)
            for (int c = 0; c < id.n_channels; ++c) {
                m.multiplex[c] = uint8_t(bs.read(4));
                if (!bs.ok() || m.multiplex[c] >= m.n_submaps) return Status::DecodeError;
            }
        for (int k = 0; k < m.n_submaps; ++k) {
            bs.read(8);
            m.submap_floor[k] = uint8_t(bs.read(8)), m.submap_residue[k] = uint8_t(bs.read(8));
            if (!bs.ok() || m.submap_floor[k] >= max_floor) return Status::DecodeError;
            if (m.submap_residue[k] >= max_residue) return Status::DecodeError;
        }
        out.mappings.push_back(std::move(m));
    }
    // ---- modes
    const uint8_t max_mapping = uint8_t(out.mappings.size());
    for (uint32_t i = 0, count = bs.read(6) + 1; i < count; ++i) {
        const bool flag = bs.read_bool();
        const uint32_t window = bs.read(16), transform = bs.read(16), mapping = bs.read(8);
        if (!bs.ok() || window != 0 || transform != 0 || mapping >= max_mapping) return Status::DecodeError;
        out.modes.emplace_back(flag, uint8_t(mapping));
    }
    if (!bs.read_bool() || !bs.ok()) return Status::DecodeError;
    return Status::Ok;
}

// mappings/vorbis.rs:45-107: (duration, leading samples to discard) of each audio packet in turn.
class VorbisPacketTimer {
  public:
    VorbisPacketTimer() = default;
    VorbisPacketTimer(const VorbisIdent& id, uint8_t num_modes, uint64_t long_block_mask)
        : mask_(long_block_mask), num_modes_(num_modes), bs0_(id.bs0_exp), bs1_(id.bs1_exp) {}
    void reset() { prev_exp_ = 0; }
    // New headers, same overlap state (a chained stream restarts its modes but a caller batching one stream in
    // several calls carries the previous block across them).
    void rebind(const VorbisIdent& id, uint8_t num_modes, uint64_t long_block_mask) {
        mask_ = long_block_mask, num_modes_ = num_modes, bs0_ = id.bs0_exp, bs1_ = id.bs1_exp;
    }
    uint8_t prev_exp() const { return prev_exp_; }  // 0: no previous block
    // A packet that is not audio, names no valid mode or is cut short takes no time and leaves the state alone.
    void next(const uint8_t* p, size_t n, uint64_t& dur, uint64_t& discard) {
        dur = discard = 0;
        BitReaderRtl bs(p, n);
        if (bs.read_bool() || !bs.ok()) return;
        const uint32_t mode = bs.read(vorbis_ilog(uint32_t(num_modes_) - 1)) & 0xff;
        if (!bs.ok() || mode >= num_modes_) return;
        const unsigned exp = (mask_ >> mode) & 1 ? bs1_ : bs0_;
        const uint64_t cur = uint64_t(1) << exp;
        if (prev_exp_) dur = ((uint64_t(1) << prev_exp_) >> 2) + (cur >> 2);
        else dur = discard = cur >> 1;  // nothing to overlap with: the lapped half is thrown away
        prev_exp_ = uint8_t(exp);
    }

  private:
    uint64_t mask_ = 0;
    uint8_t num_modes_ = 1, bs0_ = 6, bs1_ = 6, prev_exp_ = 0;
};

// symphonia-common/src/xiph/audio/vorbis/mod.rs:66-118: Matroska / WebM carry the three Vorbis headers in one
// Xiph-laced blob (count byte 2, two lacing-coded lengths, then identification, comment and setup back to back).
inline Status vorbis_unpack_xiph_laced(const uint8_t* p, size_t n, Piece& ident, Piece& setup) {
    if (n == 0 || p[0] != 2) return Status::DecodeError;
    size_t at = 1;
    uint64_t len[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
        for (;;) {
            if (at >= n) return Status::DecodeError;
            const uint8_t v = p[at++];
            len[k] += v;
            if (v < 255) break;
        }
    }
    const uint64_t rest = n - at;
    if (rest == 0 || len[0] + len[1] > rest) return Status::DecodeError;
    ident = Piece{at, uint32_t(len[0])};
    setup = Piece{at + len[0] + len[1], uint32_t(rest - len[0] - len[1])};
    return Status::Ok;
}

// symphonia-format-ogg/src/logical.rs:164-302: end trims of the stream packets of one logical stream, page by page.  A page's
// granule position is the time stamp one past the last valid sample of the last packet that ends on it; a packet whose end
// (the running sum of decoded durations from the page's start) lies beyond that loses the excess, never more than it has left
// after its leading discard.  A page starts where the previous page ended when that page also completed a stream packet (its
// sequence number is the previous one's + 1), otherwise at end - total duration; a stream whose packets all end on one page
// starts at -discard when that leaves padding at the end (the reference's single-page rule, applied here to "every stream
// packet ends on the same page").
inline void ogg_page_end_trims(const uint32_t* page_sequence, const uint64_t* page_absgp, const uint32_t* dur, const uint32_t* discard,
                               size_t n, uint32_t* trim_end) {
    bool single_page = true;
    for (size_t i = 1; i < n; ++i) single_page = single_page && page_sequence[i] == page_sequence[0];
    bool have_prev = false;
    uint32_t prev_seq = 0;
    int64_t prev_end = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i;
        int64_t tot = 0, disc = 0;
        while (j < n && page_sequence[j] == page_sequence[i]) tot += dur[j], disc += discard[j], ++j;
        const int64_t page_end = int64_t(page_absgp[i]);
        int64_t start;
        if (have_prev && prev_seq + 1 == page_sequence[i]) start = prev_end;
        else if (single_page && tot >= disc + page_end) start = -disc;
        else start = page_end - tot;
        int64_t next = start;
        for (size_t k = i; k < j; ++k) {
            next += dur[k];
            const int64_t left = int64_t(dur[k]) - int64_t(discard[k]);
            trim_end[k] = next > page_end ? uint32_t(std::min<int64_t>(next - page_end, left < 0 ? 0 : left)) : 0u;
        }
        have_prev = true, prev_seq = page_sequence[i], prev_end = page_end, i = j;
    }
}

// mappings/vorbis.rs:109-285: the per-stream state machine.  detect() on the first packet of the first page,
// map() on every later packet.
class OggVorbisMapper {
  public:
    enum class Kind : uint8_t { Audio, Comment, Setup, Unknown, Error };
    struct Mapped {
        Kind kind;
        uint64_t dur, discard;  // Audio only
    };

    // False: not a Vorbis stream (the packet is not a well-formed 30-byte identification header).
    bool detect(const uint8_t* p, size_t n) {
        if (n != 30 || vorbis_read_ident(p, n, ident_) != Status::Ok) return false;
        extra_.assign(p, p + n);
        return true;
    }

    Mapped map(const uint8_t* p, size_t n) {
        Mapped m{Kind::Error, 0, 0};
        if (n == 0) return m;
        if ((p[0] & 1) == 0) {  // even packet types are audio; before the setup header they take no time
            m.kind = Kind::Audio;
            if (have_timer_) timer_.next(p, n, m.dur, m.discard);
            return m;
        }
        if (n < 7 || std::memcmp(p + 1, "vorbis", 6) != 0) return m;
        if (p[0] == 3) return m.kind = Kind::Comment, m;
        if (p[0] != 5) return m.kind = Kind::Unknown, m;
        extra_.insert(extra_.end(), p, p + n);  // appended whether or not it parses, as in the reference
        uint8_t modes;
        uint64_t mask;
        if (vorbis_read_setup_modes(p, n, ident_, modes, mask) == Status::Ok) timer_ = VorbisPacketTimer(ident_, modes, mask), have_timer_ = true;
        ready_ = true;
        return m.kind = Kind::Setup, m;
    }

    void reset() { timer_.reset(); }
    bool ready() const { return ready_; }  // a setup header was seen
    const VorbisIdent& ident() const { return ident_; }
    // Identification packet + setup packet(s): what the decoder is constructed from.
    const std::vector<uint8_t>& extra_data() const { return extra_; }
    // Largest gap between points a decoder can start from: half the long block (:156-164).
    uint64_t max_rap_period() const { return have_timer_ ? (uint64_t(1) << ident_.bs1_exp) >> 1 : 0; }

  private:
    VorbisIdent ident_{};
    VorbisPacketTimer timer_;
    std::vector<uint8_t> extra_;
    bool have_timer_ = false, ready_ = false;
};

// =====================================================================================================================
// FLAC (native container)
// =====================================================================================================================
// "fLaC", metadata blocks (STREAMINFO first), then frames back to back.  A frame carries no length: its end is where
// the next frame starts, and the only proof of that is the CRC-16 in front of it.  The reference finds packets with a
// fragment-merging parser driven by moving averages of the frame size (symphonia-bundle-flac/src/parser.rs:149-560);
// this is a plain CRC-validated splitter instead: a frame starts at a sync code whose header parses, checks out (CRC-8)
// and fits the stream (parser.rs:586-648: rate, bit depth, channel count, block size, blocking strategy, monotonic
// sequence number), and ends at the first later such header -- or the end of the data -- in front of which the CRC-16
// of everything since the start matches.  On a well-formed file both give the same packets, time stamps and durations;
// on a damaged one this splitter drops exactly the frames whose checksum fails, which is not promised to be what the
// reference's heuristics do.

struct FlacStreamInfo {  // symphonia-common/src/xiph/audio/flac/mod.rs:78-186
    uint16_t block_min, block_max;
    uint32_t frame_min, frame_max;  // bytes, 0 = unknown
    uint32_t sample_rate;
    uint8_t channels, bits_per_sample;
    uint64_t n_samples;             // 0 = unknown
    uint8_t md5[16];
    bool has_md5;
};

inline Status flac_read_stream_info(const uint8_t* p, size_t n, FlacStreamInfo& si) {
    if (n < 34) return Status::EndOfStream;
    si.block_min = uint16_t(detail::be16(p)), si.block_max = uint16_t(detail::be16(p + 2));
    if (si.block_min < 16 || si.block_max < 16 || si.block_max < si.block_min) return Status::DecodeError;
    si.frame_min = detail::be24(p + 4), si.frame_max = detail::be24(p + 7);
    if (si.frame_min && si.frame_max && si.frame_max < si.frame_min) return Status::DecodeError;
    const uint64_t bits = uint64_t(detail::be32(p + 10)) << 32 | detail::be32(p + 14);  // 20 + 3 + 5 + 36 bits
    si.sample_rate = uint32_t(bits >> 44);
    if (si.sample_rate < 1 || si.sample_rate > 655350) return Status::DecodeError;
    si.channels = uint8_t(((bits >> 41) & 7) + 1);
    si.bits_per_sample = uint8_t(((bits >> 36) & 31) + 1);
    if (si.bits_per_sample < 4) return Status::DecodeError;
    si.n_samples = bits & 0xfffffffffull;
    std::memcpy(si.md5, p + 18, 16);
    si.has_md5 = false;
    for (int k = 0; k < 16; ++k) si.has_md5 |= si.md5[k] != 0;
    return Status::Ok;
}

inline uint8_t crc8_ccitt(const uint8_t* p, size_t n) {  // polynomial 0x07 (symphonia-core/src/checksum/crc8.rs:32-65)
    uint8_t c = 0;
    for (size_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = uint8_t(c & 0x80 ? (c << 1) ^ 0x07 : c << 1);
    }
    return c;
}
namespace detail {
struct Crc16Msb {
    uint16_t t[256];
    constexpr Crc16Msb() : t() {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i << 8;
            for (int k = 0; k < 8; ++k) c = (c & 0x8000) ? (c << 1) ^ 0x8005 : c << 1;
            t[i] = uint16_t(c);
        }
    }
};
}  // namespace detail
inline uint16_t crc16_ansi_update(uint16_t state, const uint8_t* p, size_t n) {  // polynomial 0x8005, most-significant bit first
    static constexpr detail::Crc16Msb tab{};
    for (size_t i = 0; i < n; ++i) state = uint16_t((state << 8) ^ tab.t[(state >> 8) ^ p[i]]);
    return state;
}

struct FlacFrameHeader {
    uint64_t sequence;
    bool by_sample;
    uint32_t block, sample_rate, bits_per_sample;  // rate / depth 0: not in the header
    uint8_t channels;
    uint8_t size;  // bytes, sync code to CRC-8
};

// frame.rs:81-233 at p[0..n): false unless a complete header with a matching CRC-8 starts here.
inline bool flac_parse_frame_header(const uint8_t* p, size_t n, FlacFrameHeader& h) {
    if (n < 6 || p[0] != 0xff || (p[1] & 0xfc) != 0xf8 || (p[3] & 1)) return false;
    size_t at = 4;
    h.by_sample = p[1] & 1;
    const unsigned bs = p[2] >> 4, sr = p[2] & 15, ch = p[3] >> 4, bd = (p[3] >> 1) & 7;
    uint64_t v = p[at++];
    int more;
    if (v < 0x80) more = 0;
    else if (v >= 0xc0 && v <= 0xdf) more = 1, v &= 0x1f;
    else if (v >= 0xe0 && v <= 0xef) more = 2, v &= 0x0f;
    else if (v >= 0xf0 && v <= 0xf7) more = 3, v &= 0x07;
    else if (v >= 0xf8 && v <= 0xfb) more = 4, v &= 0x03;
    else if (v >= 0xfc && v <= 0xfd) more = 5, v &= 0x01;
    else if (v == 0xfe) more = 6, v = 0;
    else return false;
    for (int k = 0; k < more; ++k) {
        if (at >= n) return false;
        v = v << 6 | (p[at++] & 0x3f);
    }
    if (v > (h.by_sample ? 0xfffffffffull : 0x7fffffffull)) return false;
    h.sequence = v;
    if (bs == 0) return false;
    if (bs == 1) h.block = 192;
    else if (bs <= 5) h.block = 576u << (bs - 2);
    else if (bs == 6) {
        if (at + 1 > n) return false;
        h.block = uint32_t(p[at++]) + 1;
    } else if (bs == 7) {
        if (at + 2 > n) return false;
        const uint32_t x = detail::be16(p + at);
        at += 2;
        if (x == 0xffff) return false;
        h.block = x + 1;
    } else h.block = 256u << (bs - 8);
    static constexpr uint32_t rates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    if (sr < 12) h.sample_rate = rates[sr];
    else if (sr == 12) {
        if (at + 1 > n) return false;
        h.sample_rate = uint32_t(p[at++]) * 1000;
    } else if (sr == 15) return false;
    else {
        if (at + 2 > n) return false;
        h.sample_rate = detail::be16(p + at) * (sr == 14 ? 10u : 1u);
        at += 2;
    }
    if (sr != 0 && (h.sample_rate < 1 || h.sample_rate > 655350)) return false;
    static constexpr uint8_t widths[8] = {0, 8, 12, 255, 16, 20, 24, 32};
    if (widths[bd] == 255) return false;
    h.bits_per_sample = widths[bd];
    if (ch <= 7) h.channels = uint8_t(ch + 1);
    else if (ch <= 10) h.channels = 2;
    else return false;
    if (at + 1 > n || p[at] != crc8_ccitt(p, at)) return false;
    h.size = uint8_t(at + 1);
    return true;
}

struct FlacPacket {
    uint64_t offset;
    uint32_t size;
    uint64_t ts;   // first sample (parser.rs:566-584)
    uint32_t dur;  // block size
};

class FlacIndexer {
  public:
    FlacIndexer(const uint8_t* data, size_t n) : d_(data), n_(n) {}

    // demuxer.rs:60-170: the stream marker, then metadata blocks up to the one flagged last; the first must be STREAMINFO.
    Status open() {
        if (n_ < 4 || std::memcmp(d_, "fLaC", 4) != 0) return Status::Unsupported;
        size_t at = 4;
        bool first = true;
        for (;;) {
            if (at + 4 > n_) return Status::EndOfStream;
            const bool last = d_[at] & 0x80;
            const unsigned type = d_[at] & 0x7f;
            const size_t len = detail::be24(d_ + at + 1);
            at += 4;
            if (at + len > n_) return Status::EndOfStream;
            if (first) {
                if (type != 0 || len != 34) return Status::DecodeError;
                const Status s = flac_read_stream_info(d_ + at, len, info_);
                if (s != Status::Ok) return s;
                first = false;
            }
            at += len;
            if (last) break;
        }
        pos_ = first_frame_ = at;
        have_last_ = false;
        return Status::Ok;
    }
    const FlacStreamInfo& info() const { return info_; }
    size_t first_frame_pos() const { return first_frame_; }
    size_t skipped_bytes() const { return skipped_; }  // bytes between packets that belonged to no valid frame

    Status next(FlacPacket& pk) {
        for (size_t start = pos_; start + 2 <= n_;) {
            FlacFrameHeader h;
            if (!candidate(start, h)) {
                const size_t q = next_sync(start + 1);
                skipped_ += q - start;
                start = pos_ = q;
                continue;
            }
            // the end: the next plausible header (or the end of the data) that the CRC-16 vouches for
            uint16_t crc = 0;
            size_t done = start;  // crc covers [start, done)
            for (size_t q = next_sync(start + h.size);; q = next_sync(q + 1)) {
                FlacFrameHeader nh;
                const bool at_end = q >= n_;
                if (at_end) q = n_;
                if (q - start >= 2 + size_t(h.size) && (at_end || candidate(q, nh, &h))) {
                    crc = crc16_ansi_update(crc, d_ + done, q - 2 - done), done = q - 2;
                    if (crc == detail::be16(d_ + q - 2)) {
                        pk.offset = start, pk.size = uint32_t(q - start), pk.dur = h.block;
                        const bool fixed = info_.block_min == info_.block_max;
                        pk.ts = h.by_sample ? h.sequence : h.sequence * (fixed ? info_.block_min : h.block);
                        last_ = h, have_last_ = true, pos_ = q;
                        return Status::Ok;
                    }
                }
                if (at_end || q - start > kMaxFrame) break;
            }
            const size_t q = next_sync(start + 1);  // no end vouched for: this was not a frame
            skipped_ += q - start;
            start = pos_ = q;
        }
        return Status::EndOfStream;
    }

    static Status index(const uint8_t* data, size_t n, FlacStreamInfo& info, std::vector<FlacPacket>& out) {
        FlacIndexer ix(data, n);
        const Status s = ix.open();
        if (s != Status::Ok) return s;
        info = ix.info();
        FlacPacket p;
        while (ix.next(p) == Status::Ok) out.push_back(p);
        return Status::Ok;
    }

  private:
    static constexpr size_t kMaxFrame = 16u * 1024 * 1024;  // frame.rs:17

    size_t next_sync(size_t from) const {
        for (size_t q = from; q + 2 <= n_;) {
            const void* hit = std::memchr(d_ + q, 0xff, n_ - q);
            if (!hit) break;
            q = size_t(static_cast<const uint8_t*>(hit) - d_);
            if (q + 2 > n_) break;
            if ((d_[q + 1] & 0xfc) == 0xf8) return q;
            ++q;
        }
        return n_;
    }
    // A header at `at` that fits the stream (parser.rs:586-648) and follows `prev` (default: the last accepted frame).
    bool candidate(size_t at, FlacFrameHeader& h, const FlacFrameHeader* prev = nullptr) const {
        if (at + 6 > n_ || !flac_parse_frame_header(d_ + at, n_ - at, h)) return false;
        if (h.sample_rate && h.sample_rate != info_.sample_rate) return false;
        if (h.bits_per_sample && h.bits_per_sample != info_.bits_per_sample) return false;
        if (h.block > info_.block_max || h.channels != info_.channels) return false;
        const bool fixed = info_.block_min == info_.block_max;
        if (h.by_sample == fixed) return false;
        const FlacFrameHeader* before = prev ? prev : (have_last_ ? &last_ : nullptr);
        const uint64_t last_seq = before ? before->sequence : 0;
        return h.sequence > last_seq || h.sequence == 0;
    }

    const uint8_t* d_;
    size_t n_;
    size_t pos_ = 0, first_frame_ = 0, skipped_ = 0;
    FlacStreamInfo info_{};
    FlacFrameHeader last_{};
    bool have_last_ = false;
};

}  // namespace packet
}  // namespace symgpu
