"""Packetiser oracle: a CPU restatement of the byte rules by which the reference cuts MPEG audio,
ADTS and Ogg byte streams into codec packets.  TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product
(include/symgpu/packetizer.hpp is the product and shares no code with this file).

The reference drives a consuming byte reader (MediaSourceStream); this file keeps that shape on purpose -- a cursor
that reads, fails at end of data and seeks back -- so that every rule can be checked against the cited lines:

  MPEG audio   symphonia-bundle-mp3/src/header.rs:49-237 (header word), demuxer.rs:160-218 (next_packet),
               demuxer.rs:414-487 (open: strict first frame, Xing/Info/VBRI tag, duration estimate),
               demuxer.rs:585-681 (frame read, strict check, main_data_begin), demuxer.rs:683-733 (estimate),
               demuxer.rs:761-1047 (tags), common.rs:155-212 (derived sizes), symphonia-core/src/packet.rs:318-343 (trim)
  ADTS         symphonia-codec-aac/src/adts.rs:130-246 (header), 278-309 (next_packet),
               symphonia-common/src/mpeg/audio/mod.rs:178-214 (rate / channel tables)
  Ogg          symphonia-format-ogg/src/page.rs:14-300 (page sync, header, lacing, CRC),
               logical.rs:104-205, 577-620 (packet assembly across pages), symphonia-core/src/checksum/crc32.rs:543-600

Pinned by (tests/test_packetizer.py): the reference's own CRC-32 known answers (crc32.rs:602-640), its tag-heuristic
unit tests (demuxer.rs:1054-1112), the CRC catalogue check values of the two polynomials, and frame sizes every MPEG
audio text quotes (128 kbit/s at 44.1 kHz = 417 / 418 bytes)."""

EOF = "eof"          # the reader ran out of bytes (IoError UnexpectedEof in the reference)
DECODE = "decode"    # Error::DecodeError
UNSUPPORTED = "unsupported"


class ReaderError(Exception):
    def __init__(self, kind, msg=""):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind


class Reader:
    """MediaSourceStream / BufReader stand-in: a cursor over bytes."""

    def __init__(self, data, pos=0):
        self.data = bytes(data)
        self.pos = pos

    def available(self):
        return len(self.data) - self.pos

    def read_u8(self):
        if self.pos >= len(self.data):
            raise ReaderError(EOF)
        b = self.data[self.pos]
        self.pos += 1
        return b

    def read_exact(self, n):
        if self.pos + n > len(self.data):
            self.pos = len(self.data)
            raise ReaderError(EOF)
        b = self.data[self.pos:self.pos + n]
        self.pos += n
        return b

    def read_be(self, n):
        return int.from_bytes(self.read_exact(n), "big")

    def read_le(self, n):
        return int.from_bytes(self.read_exact(n), "little")


# ------------------------------------------------------------------------------------------------ checksums

def _crc32_byte(c):
    for _ in range(8):
        c = ((c << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if c & 0x80000000 else (c << 1) & 0xFFFFFFFF
    return c


_CRC32 = [_crc32_byte(i << 24) for i in range(256)]


def crc32_update(state, buf):
    """crc32.rs:543-570: polynomial 0x04c11db7, no reflection, no final xor, caller-chosen initial state."""
    for b in buf:
        state = ((state << 8) & 0xFFFFFFFF) ^ _CRC32[(state >> 24) ^ b]
    return state


def crc16_ansi_le_update(state, buf):
    """crc16.rs:377-404: polynomial 0x8005 processed least-significant bit first (table entry 1 = 0xc0c1)."""
    for b in buf:
        state ^= b
        for _ in range(8):
            state = (state >> 1) ^ 0xA001 if state & 1 else state >> 1
    return state


# ------------------------------------------------------------------------------------------------ MPEG audio

_BITRATES = {  # header.rs:19-47, kbit/s
    ("1", 1): [0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448],
    ("1", 2): [0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384],
    ("1", 3): [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320],
    ("2", 1): [0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256],
    ("2", 23): [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160],
}
_RATES = {"1": [44100, 48000, 32000], "2": [22050, 24000, 16000], "2.5": [11025, 12000, 8000]}


def mpa_check_header(w):
    """header.rs:49-69."""
    return not ((w >> 19) & 3 == 1 or (w >> 17) & 3 == 0 or (w >> 12) & 15 == 15 or (w >> 10) & 3 == 3)


def mpa_is_synced(w):
    """header.rs:71-75."""
    return (w & 0xFFE00000) == 0xFFE00000


def mpa_parse_header(w):
    """header.rs:107-233.  Returns a dict or raises ReaderError(DECODE / UNSUPPORTED)."""
    v = (w >> 19) & 3
    if v == 1:
        raise ReaderError(DECODE, "version")
    version = {0: "2.5", 2: "2", 3: "1"}[v]
    l = (w >> 17) & 3
    if l == 0:
        raise ReaderError(DECODE, "layer")
    layer = 4 - l
    bi = (w >> 12) & 15
    if bi == 0:
        raise ReaderError(UNSUPPORTED, "free bit-rate")
    if bi == 15:
        raise ReaderError(DECODE, "bit-rate")
    if version == "1":
        bitrate = _BITRATES[("1", layer)][bi] * 1000
    else:
        bitrate = _BITRATES[("2", 1 if layer == 1 else 23)][bi] * 1000
    ri = (w >> 10) & 3
    if ri == 3:
        raise ReaderError(DECODE, "sample rate")
    sample_rate = _RATES[version][ri]
    sample_rate_idx = ri + {"1": 0, "2": 3, "2.5": 6}[version]
    m = (w >> 6) & 3
    mode = {0: "stereo", 1: "joint", 2: "dual", 3: "mono"}[m]
    mid_side = intensity = False
    bound = 32
    if mode == "joint":
        if layer == 3:
            mid_side, intensity = bool(w & 0x20), bool(w & 0x10)
        else:
            bound = (1 + ((w & 0x30) >> 4)) << 2
    if layer == 2:  # header.rs:176-187
        if mode == "mono":
            if bitrate in (224000, 256000, 320000, 384000):
                raise ReaderError(DECODE, "layer 2 mono bit-rate")
        elif bitrate in (32000, 48000, 56000, 80000):
            raise ReaderError(DECODE, "layer 2 stereo bit-rate")
    padding = bool(w & 0x200)
    factor = {1: 12, 2: 144, 3: 144 if version == "1" else 72}[layer]
    slot = 4 if layer == 1 else 1
    slots = factor * bitrate // sample_rate + int(padding)
    n_ch = 1 if mode == "mono" else 2
    return dict(version=version, layer=layer, bitrate=bitrate, sample_rate=sample_rate, sample_rate_idx=sample_rate_idx,
                mode=mode, mid_side=mid_side, intensity=intensity, bound=bound, emphasis={1: 1, 3: 3}.get(w & 3, 0),
                copyrighted=bool(w & 8), original=bool(w & 4), padding=padding, crc=(w & 0x10000) == 0,
                frame_size=slots * slot - 4, n_channels=n_ch,
                samples={1: 384, 2: 1152, 3: 1152 if version == "1" else 576}[layer],
                side_info_len=(17 if n_ch == 1 else 32) if version == "1" else (9 if n_ch == 1 else 17),
                header_size=4 + (2 if (w & 0x10000) == 0 else 0))


def mpa_sync_frame(r):
    """header.rs:77-103."""
    sync = 0
    while True:
        while not mpa_is_synced(sync):
            sync = ((sync << 8) | r.read_u8()) & 0xFFFFFFFF
        if mpa_check_header(sync):
            return sync
        sync = ((sync << 8) | r.read_u8()) & 0xFFFFFFFF


def mpa_read_frame(r):
    """demuxer.rs:585-607.  Returns (header, word, start offset, packet bytes)."""
    while True:
        w = mpa_sync_frame(r)
        try:
            h = mpa_parse_header(w)
            break
        except ReaderError:
            continue
    start = r.pos - 4
    body = r.read_exact(h["frame_size"])
    return h, w, start, w.to_bytes(4, "big") + body


def _similar(h, w):
    """demuxer.rs:643-656."""
    try:
        c = mpa_parse_header(w)
    except ReaderError:
        return False
    return (h["version"], h["layer"], h["sample_rate"], h["n_channels"]) == \
           (c["version"], c["layer"], c["sample_rate"], c["n_channels"])


def mpa_read_frame_strict(r):
    """demuxer.rs:610-640."""
    while True:
        h, w, start, pkt = mpa_read_frame(r)
        pos = r.pos
        try:
            nxt = r.read_be(4)
        except ReaderError:
            nxt = None
        if nxt is not None and (not mpa_is_synced(nxt) or not _similar(h, nxt)):
            r.pos = r.pos - (len(pkt) + 4 - 1)
            continue
        r.pos = pos
        return h, w, start, pkt


def mpa_main_data_begin(pkt, h):
    """demuxer.rs:664-680 (applied to the bytes after the header word)."""
    r = Reader(pkt, 4)
    if h["crc"]:
        r.read_be(2)
    return r.read_be(2) >> 7 if h["version"] == "1" else r.read_u8()


def mpa_is_maybe_info_tag(buf, h):
    """demuxer.rs:942-968."""
    if h["layer"] != 3:
        return False
    off = 4 + h["side_info_len"]
    if len(buf) < off + 8:
        return False
    if buf[off:off + 4] not in (b"Xing", b"Info"):
        return False
    return not any(buf[h["header_size"]:off])


def mpa_read_info_tag(buf, h):
    """demuxer.rs:761-925.  None when the frame is not a (readable) tag."""
    if not mpa_is_maybe_info_tag(buf, h):
        return None
    off = 4 + h["side_info_len"]
    crc = crc16_ansi_le_update(0, buf[:off])
    r = Reader(buf[off:])

    def take(n):  # MonitorStream: bytes read through it enter the CRC
        nonlocal crc
        b = r.read_exact(n)
        crc = crc16_ansi_le_update(crc, b)
        return b

    try:
        ident = take(4)
        if ident not in (b"Xing", b"Info"):
            return None
        flags = int.from_bytes(take(4), "big")
        num_frames = int.from_bytes(take(4), "big") if flags & 1 else None
        num_bytes = int.from_bytes(take(4), "big") if flags & 2 else None
        toc = take(100) if flags & 4 else None
        quality = int.from_bytes(take(4), "big") if flags & 8 else None
        lame = None
        if r.available() >= 24:
            enc = take(9)
            take(1), take(1)
            peak = int.from_bytes(take(4), "big")
            take(2), take(2)
            take(1), take(1)
            trim = int.from_bytes(take(3), "big")
            if enc[:4] in (b"LAME", b"Lavf", b"Lavc"):
                delay, padding = 528 + 1 + (trim >> 12), max((trim & 0xFFF) - 529, 0)
            else:
                delay, padding = 0, 0
            written = None
            if r.available() >= 12:
                take(1), take(1), take(2), take(4), take(2)
                if h["crc"] or enc[:4] == b"LAME":
                    written = r.read_be(2)  # read past the monitor: not part of the sum
            if written is None or written == 0 or written == crc:
                lame = dict(encoder=enc, delay=delay, padding=padding, peak=peak)
        return dict(num_frames=num_frames, num_bytes=num_bytes, has_toc=toc is not None, quality=quality,
                    is_cbr=ident == b"Info", lame=lame)
    except ReaderError:
        return None


def mpa_is_maybe_vbri_tag(buf, h):
    """demuxer.rs:1023-1047."""
    if h["layer"] != 3 or len(buf) < 36 + 26 or buf[36:40] != b"VBRI":
        return False
    return not any(buf[h["header_size"]:36])


def mpa_read_vbri_tag(buf, h):
    """demuxer.rs:980-1019."""
    if not mpa_is_maybe_vbri_tag(buf, h):
        return None
    r = Reader(buf, 36)
    try:
        if r.read_exact(4) != b"VBRI" or r.read_be(2) != 1:
            return None
        r.read_be(2), r.read_be(2)
        return dict(num_bytes=r.read_be(4), num_mpeg_frames=r.read_be(4))
    except ReaderError:
        return None


def mpa_estimate_frames(r):
    """demuxer.rs:683-733 (byte length known)."""
    start = r.pos
    total_len = len(r.data) - start
    total_frame_len = total_frames = 0
    result = None
    while True:
        try:
            h = mpa_parse_header(r.read_be(4))
            total_frame_len += 4 + h["frame_size"]
            total_frames += 1
            r.read_exact(h["frame_size"])
        except ReaderError:
            break
        if total_frames > 16 or total_frame_len > 16 * 1024:
            result = int(float(total_len) / (float(total_frame_len) / float(total_frames)))
            break
    r.pos = start
    return result


def mpa_index(data, seekable=True):
    """Open + read to the end (demuxer.rs:414-487, 160-218).  Returns (track, packets); `track` is None when no first
    frame exists.  A packet is (offset, size, header word, pts, dur, trim_start, trim_end)."""
    r = Reader(data)
    try:
        h, w, start, pkt = mpa_read_frame_strict(r)
    except ReaderError:
        return None, []
    track = dict(header=h, word=w, delay=None, padding=None, num_frames=None, tag=None)
    info = mpa_read_info_tag(pkt, h)
    vbri = None if info is not None else mpa_read_vbri_tag(pkt, h)
    if info is not None:
        track["tag"] = "info" if info["is_cbr"] else "xing"
        if info["lame"] is not None:
            track["delay"], track["padding"] = info["lame"]["delay"], info["lame"]["padding"]
        if info["num_frames"] is not None:
            total = info["num_frames"] * h["samples"]
            track["num_frames"] = max(total - ((track["delay"] or 0) + (track["padding"] or 0)), 0)
    elif vbri is not None:
        track["tag"] = "vbri"
        track["num_frames"] = vbri["num_mpeg_frames"] * h["samples"]
    else:
        r.pos -= 4 + h["frame_size"]
        if seekable:
            n = mpa_estimate_frames(r)
            if n is not None:
                track["num_frames"] = n * h["samples"]
    track["first_packet_pos"] = r.pos
    ts = -(track["delay"] or 0)
    packets = []
    while True:
        try:
            h, w, start, pkt = mpa_read_frame(r)
        except ReaderError:
            break
        if mpa_is_maybe_info_tag(pkt, h):
            if mpa_read_info_tag(pkt, h) is not None:
                continue
        elif mpa_is_maybe_vbri_tag(pkt, h) and mpa_read_vbri_tag(pkt, h) is not None:
            continue
        dur = h["samples"]
        trim_start = min(max(-ts, 0), dur)   # packet.rs:327-330
        trim_end = 0
        if track["num_frames"] is not None:  # packet.rs:334-338
            trim_end = max(ts + dur - track["num_frames"], 0)
        packets.append((start, len(pkt), w, ts, dur, trim_start, trim_end))
        ts += dur
    return track, packets


# ------------------------------------------------------------------------------------------------ ADTS

_MPEG4_RATES = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350]
_MPEG4_CHANNELS = {1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 8}  # audio/mod.rs:201-213: channel counts of the layouts


def adts_read_header(r):
    """adts.rs:230-236 + 137-198: resync, then the header body.  Returns a dict with the header's offset."""
    sync = 0
    while (sync & 0xFFF6) != 0xFFF0:
        sync = ((sync << 8) | r.read_u8()) & 0xFFFF
    start = r.pos - 2
    has_crc = sync & 1 == 0
    n = 9 if has_crc else 7
    bits = int.from_bytes(r.read_exact(n - 2), "big")
    width = (n - 2) * 8

    def field(at, length):
        return (bits >> (width - at - length)) & ((1 << length) - 1)

    aot = field(0, 2) + 1  # 1..4 are all defined object types (Main, LC, SSR, LTP)
    ri = field(2, 4)
    if ri == 15:
        raise ReaderError(DECODE, "forbidden sample rate")
    if ri > 12:
        raise ReaderError(DECODE, "invalid sample rate")
    ci = field(7, 3)
    channels = _MPEG4_CHANNELS.get(ci, 0)  # 0: defined in-band
    frame_len = field(14, 13)
    if frame_len < n:
        raise ReaderError(DECODE, "frame length")
    if field(38, 2) + 1 > 1:
        raise ReaderError(UNSUPPORTED, "raw data blocks")
    return dict(offset=start, header_len=n, frame_len=frame_len, profile=aot, sample_rate=_MPEG4_RATES[ri],
                channels=channels, crc=field(40, 16) if has_crc else None)


def adts_index(data):
    """adts.rs:278-309 until the first error.  Returns (packets, stop) with packets = (payload offset, payload length,
    ts, sample rate, channels, profile) and stop in {"eof", "truncated", "decode", "unsupported"}."""
    r = Reader(data)
    out = []
    ts = 0
    while True:
        try:
            h = adts_read_header(r)
        except ReaderError as e:
            return out, e.kind
        payload = h["frame_len"] - h["header_len"]
        at = r.pos
        try:
            r.read_exact(payload)
        except ReaderError:
            return out, "truncated"
        out.append((at, payload, ts, h["sample_rate"], h["channels"], h["profile"]))
        ts += 1024


# ------------------------------------------------------------------------------------------------ Ogg

def ogg_try_next_page(r):
    """page.rs:166-255.  Returns (header dict, packet lengths, body offset, body length) or raises ReaderError."""
    marker = r.read_be(4)
    while marker != 0x4F676753:  # "OggS"
        marker = ((marker << 8) | r.read_u8()) & 0xFFFFFFFF
    sync_pos = r.pos
    rest = r.read_exact(23)
    hdr = b"OggS" + rest
    if hdr[4] != 0:
        raise ReaderError(DECODE, "version")
    flags = hdr[5]
    if flags & 0xF8:
        raise ReaderError(DECODE, "flags")
    header = dict(absgp=int.from_bytes(hdr[6:14], "little"), serial=int.from_bytes(hdr[14:18], "little"),
                  sequence=int.from_bytes(hdr[18:22], "little"), crc=int.from_bytes(hdr[22:26], "little"),
                  n_segments=hdr[26], continuation=bool(flags & 1), first=bool(flags & 2), last=bool(flags & 4),
                  offset=sync_pos - 4)
    crc = crc32_update(0, hdr[:22] + bytes(4) + hdr[26:])
    lacing = r.read_exact(header["n_segments"])
    crc = crc32_update(crc, lacing)
    lens, run = [], 0
    for seg in lacing:
        run += seg
        if seg < 255:
            lens.append(run)
            run = 0
    body_at = r.pos
    body = r.read_exact(sum(lacing))
    crc = crc32_update(crc, body)
    if crc != header["crc"]:
        r.pos = sync_pos
        raise ReaderError(DECODE, "crc")
    return header, lens, body_at, len(body)


class OggLogical:
    """logical.rs:104-205 + 577-620 without the codec mapper: packets as lists of (offset, length) pieces."""
    MAX_PACKET_LEN = 16 * 1024 * 1024

    def __init__(self):
        self.part = []       # pieces of the packet still open
        self.part_len = 0
        self.prev_seq = None
        self.packets = []    # (pieces, page sequence, absgp of the page it completed on, last-on-page flag)

    def read_page(self, header, lens, body_at, body_len):
        if self.prev_seq is not None:
            if header["sequence"] < self.prev_seq or header["sequence"] - self.prev_seq > 1:
                self.part, self.part_len = [], 0
        self.prev_seq = header["sequence"]
        if not header["continuation"] and self.part_len > 0:
            self.part, self.part_len = [], 0
        lens = list(lens)
        at = body_at
        if header["continuation"] and self.part_len == 0:
            if not lens:
                return
            at += lens.pop(0)
        first_new = len(self.packets)
        for n in lens:
            pieces = self.part + [(at, n)]
            self.part, self.part_len = [], 0
            self.packets.append([pieces, header["sequence"], header["absgp"], False])
            at += n
        rest = body_at + body_len - at
        if rest > 0:
            if self.part_len + rest > self.MAX_PACKET_LEN:
                raise ReaderError(DECODE, "packet too large")
            self.part.append((at, rest))
            self.part_len += rest
        if len(self.packets) > first_new:
            self.packets[-1][3] = True


def ogg_index(data):
    """Every page to the end of the data (page.rs:259-271: corrupt pages are skipped), routed by serial; a logical
    stream exists from its first-page (BOS) flag on (demuxer.rs:320-345).  Returns (pages, {serial: packets})."""
    r = Reader(data)
    pages, streams = [], {}
    while True:
        try:
            header, lens, body_at, body_len = ogg_try_next_page(r)
        except ReaderError as e:
            if e.kind == EOF:
                break
            continue
        pages.append((header["offset"], header["serial"], header["sequence"], header["absgp"], len(lens), body_len))
        if header["first"] and header["serial"] not in streams:
            streams[header["serial"]] = OggLogical()
        s = streams.get(header["serial"])
        if s is not None:
            s.read_page(header, lens, body_at, body_len)
    return pages, {k: v.packets for k, v in streams.items()}


# ------------------------------------------------------------------------------------------------ Vorbis in Ogg
# symphonia-format-ogg/src/mappings/vorbis.rs (identification header :293-360, setup walk :362-708, packet timing
# :45-107, mapper :109-285) and symphonia-common/src/xiph/audio/vorbis/mod.rs:66-118 (Xiph-laced extra data).

class BitsRtl:
    """BitReaderRtl (symphonia-core/src/io/bit.rs:941-1027): least-significant bit first; reading or skipping past
    the end is an error."""

    def __init__(self, data):
        self.v = int.from_bytes(bytes(data), "little")
        self.n = len(data) * 8
        self.at = 0

    def read(self, width):
        if self.at + width > self.n:
            raise ReaderError(EOF, "bits")
        out = (self.v >> self.at) & ((1 << width) - 1)
        self.at += width
        return out

    def read_bool(self):
        return self.read(1) == 1

    def ignore(self, width):
        if self.at + width > self.n:
            raise ReaderError(EOF, "bits")
        self.at += width


def ilog(x):
    return x.bit_length()


def vorbis_read_ident(buf):
    r = Reader(buf)
    if r.read_u8() != 1:
        raise ReaderError(DECODE, "packet type")
    if r.read_exact(6) != b"vorbis":
        raise ReaderError(DECODE, "signature")
    if r.read_le(4) != 0:
        raise ReaderError(UNSUPPORTED, "version")
    ch = r.read_u8()
    if ch == 0:
        raise ReaderError(DECODE, "channels")
    rate = r.read_le(4)
    if rate == 0:
        raise ReaderError(DECODE, "rate")
    r.read_le(4), r.read_le(4), r.read_le(4)
    bs = r.read_u8()
    bs0, bs1 = bs & 15, bs >> 4
    if not 6 <= bs0 <= 13 or not 6 <= bs1 <= 13 or bs0 > bs1:
        raise ReaderError(DECODE, "block sizes")
    if r.read_u8() != 1:
        raise ReaderError(DECODE, "framing")
    return dict(n_channels=ch, sample_rate=rate, bs0_exp=bs0, bs1_exp=bs1)


def _lookup1_values(entries, dims):
    """:717-730 -- the reference takes the float root and asserts v^dims <= entries < (v+1)^dims; that IS the definition."""
    v = 0
    while (v + 1) ** dims <= entries:
        v += 1
    return v


def _skip_codebook(bs):
    if bs.read(24) != 0x564342:
        raise ReaderError(DECODE, "codebook sync")
    dims, entries = bs.read(16), bs.read(24)
    if not bs.read_bool():
        if bs.read_bool():
            for _ in range(entries):
                if bs.read_bool():
                    bs.read(5)
        else:
            bs.ignore(entries * 5)
    else:
        cur = 0
        bs.read(5)
        while True:
            cur += bs.read(ilog(entries - cur) if entries > cur else 0)
            if cur > entries:
                raise ReaderError(DECODE, "codebook")
            if cur == entries:
                break
    lookup = bs.read(4)
    if lookup == 0:
        return
    if lookup > 2:
        raise ReaderError(DECODE, "lookup type")
    bs.read(32), bs.read(32)
    value_bits = bs.read(4) + 1
    bs.read_bool()
    if lookup == 1 and dims == 0:
        raise ReaderError(DECODE, "lookup 1 with no dimensions")  # the reference's assertion territory
    bs.ignore((_lookup1_values(entries, dims) if lookup == 1 else entries * dims) * value_bits)


def _skip_floor(bs):
    kind = bs.read(16)
    if kind == 0:
        bs.ignore(8 + 16 + 16 + 6 + 8)
        bs.ignore((bs.read(4) + 1) * 8)
    elif kind == 1:
        parts = bs.read(5)
        classes = [bs.read(4) for _ in range(parts)]
        dims = {}
        if parts:
            for c in range(max(classes) + 1):
                dims[c] = bs.read(3) + 1
                sub = bs.read(2)
                if sub:
                    bs.read(8)
                bs.ignore((1 << sub) * 8)
        bs.read(2)
        rangebits = bs.read(4)
        for c in classes:
            bs.ignore(dims[c] * rangebits)
    else:
        raise ReaderError(DECODE, "floor type")


def _skip_residue(bs):
    bs.read(16)
    bs.ignore(72)
    classes = bs.read(6) + 1
    bs.ignore(8)
    books = 0
    for _ in range(classes):
        low = bs.read(3)
        high = bs.read(5) if bs.read_bool() else 0
        books += bin((high << 3) | low).count("1")
    bs.ignore(books * 8)


def _skip_mapping(bs, channels):
    if bs.read(16) != 0:
        raise ReaderError(DECODE, "mapping type")
    submaps = bs.read(4) + 1 if bs.read_bool() else 1
    if bs.read_bool():
        steps = bs.read(8) + 1
        width = ilog(channels - 1)
        for _ in range(steps):
            bs.read(width), bs.read(width)
    if bs.read(2) != 0:
        raise ReaderError(DECODE, "reserved")
    if submaps > 1:
        bs.ignore(channels * 4)
    bs.ignore(submaps * 24)


def vorbis_read_setup_modes(buf, ident):
    """:362-405.  Returns the list of block flags."""
    r = Reader(buf)
    if r.read_u8() != 5:
        raise ReaderError(DECODE, "packet type")
    if r.read_exact(6) != b"vorbis":
        raise ReaderError(DECODE, "signature")
    bs = BitsRtl(buf[7:])
    for _ in range(bs.read(8) + 1):
        _skip_codebook(bs)
    for _ in range(bs.read(6) + 1):
        if bs.read(16) != 0:
            raise ReaderError(DECODE, "time domain transform")
    for _ in range(bs.read(6) + 1):
        _skip_floor(bs)
    for _ in range(bs.read(6) + 1):
        _skip_residue(bs)
    for _ in range(bs.read(6) + 1):
        _skip_mapping(bs, ident["n_channels"])
    modes = []
    for _ in range(bs.read(6) + 1):
        flag = bs.read_bool()
        window, transform = bs.read(16), bs.read(16)
        bs.read(8)
        if window != 0 or transform != 0:
            raise ReaderError(DECODE, "mode")
        modes.append(flag)
    if not bs.read_bool():
        raise ReaderError(DECODE, "framing")
    return modes


class VorbisTimer:
    """:45-107."""

    def __init__(self, ident, modes):
        self.ident, self.modes, self.prev = ident, modes, None

    def next(self, packet):
        bs = BitsRtl(packet)
        try:
            if bs.read_bool():
                return 0, 0
            mode = bs.read(ilog(len(self.modes) - 1))
        except ReaderError:
            return 0, 0
        if mode >= len(self.modes):
            return 0, 0
        cur = 1 << (self.ident["bs1_exp"] if self.modes[mode] else self.ident["bs0_exp"])
        if self.prev is not None:
            out = ((self.prev >> 2) + (cur >> 2), 0)
        else:
            out = (cur >> 1, cur >> 1)
        self.prev = cur
        return out


class VorbisMapper:
    """:109-285 (comment contents are metadata and not looked at)."""

    def __init__(self):
        self.ident = self.timer = None
        self.extra = b""
        self.ready = False

    def detect(self, packet):
        if len(packet) != 30:
            return False
        try:
            self.ident = vorbis_read_ident(packet)
        except ReaderError:
            return False
        self.extra = bytes(packet)
        return True

    def map(self, packet):
        if len(packet) == 0:
            return ("error", 0, 0)
        if packet[0] & 1 == 0:
            dur, discard = self.timer.next(packet) if self.timer else (0, 0)
            return ("audio", dur, discard)
        if len(packet) < 7 or packet[1:7] != b"vorbis":
            return ("error", 0, 0)
        if packet[0] == 3:
            return ("comment", 0, 0)
        if packet[0] != 5:
            return ("unknown", 0, 0)
        self.extra += bytes(packet)
        try:
            self.timer = VorbisTimer(self.ident, vorbis_read_setup_modes(packet, self.ident))
        except ReaderError:
            pass
        self.ready = True
        return ("setup", 0, 0)


def vorbis_unpack_xiph_laced(extradata):
    """xiph/audio/vorbis/mod.rs:66-118.  Returns (identification packet, setup packet)."""
    if len(extradata) == 0 or extradata[0] != 2:
        raise ReaderError(DECODE, "lacing count")
    at = 1
    lengths = []
    for _ in range(2):
        total = 0
        while True:
            if at >= len(extradata):
                raise ReaderError(DECODE, "truncated lacing")
            v = extradata[at]
            at += 1
            total += v
            if v < 255:
                break
        lengths.append(total)
    rest = extradata[at:]
    if len(rest) == 0 or lengths[0] + lengths[1] > len(rest):
        raise ReaderError(DECODE, "lengths")
    return rest[:lengths[0]], rest[lengths[0] + lengths[1]:]


# ------------------------------------------------------------------------------------------------ Vorbis setup, decoder view
# symphonia-codec-vorbis/src/lib.rs:490-770, floor.rs:160-201 / :455-560 / :748-773, residue.rs:73-140.  Codebook contents are
# walked with the syntax checks of the mapper (above), not built.

def _find_neighbors(vec, x):
    """floor.rs:748-773."""
    bound = vec[x]
    low, high = 0, 0xFFFFFFFF
    res = [0, 0]
    for i, xv in enumerate(vec[:x]):
        if low < xv < bound:
            low, res[0] = xv, i
        if bound < xv < high:
            high, res[1] = xv, i
    return res


def vorbis_read_setup(buf, ident):
    """Returns dict(n_codebooks, floors [dict], n_residues, mappings [dict], modes [(flag, mapping)]) or raises ReaderError."""
    r = Reader(buf)
    if r.read_u8() != 5:
        raise ReaderError(DECODE, "packet type")
    if r.read_exact(6) != b"vorbis":
        raise ReaderError(DECODE, "signature")
    bs = BitsRtl(buf[7:])
    n_books = bs.read(8) + 1
    for _ in range(n_books):
        _skip_codebook(bs)
    for _ in range(bs.read(6) + 1):
        if bs.read(16) != 0:
            raise ReaderError(DECODE, "time domain transform")
    max_book = n_books & 0xFF
    floors = []
    for _ in range(bs.read(6) + 1):
        kind = bs.read(16)
        if kind == 0:
            bs.ignore(8 + 16 + 16 + 6 + 8)
            for _ in range(bs.read(4) + 1):
                if bs.read(8) >= max_book:
                    raise ReaderError(DECODE, "floor0 book")
            floors.append(dict(type=0))
        elif kind == 1:
            parts = bs.read(5)
            classes = [bs.read(4) for _ in range(parts)]
            dims = {}
            class_info = {}
            if parts:
                for c in range(max(classes) + 1):
                    dims[c] = bs.read(3) + 1
                    sub = bs.read(2)
                    mainbook = 0
                    if sub:
                        mainbook = bs.read(8)
                        if mainbook >= max_book:
                            raise ReaderError(DECODE, "floor1 master book")
                    subbooks, used = [0] * 8, 0
                    for k in range(1 << sub):
                        book = bs.read(8)
                        if book > 0:
                            book -= 1
                            if book >= max_book:
                                raise ReaderError(DECODE, "floor1 sub book")
                            used |= 1 << k
                        subbooks[k] = book
                    class_info[c] = dict(dimensions=dims[c], subclass_bits=sub, mainbook=mainbook, subbooks=subbooks, used=used)
            mult = bs.read(2) + 1
            rangebits = bs.read(4)
            x_list = [0, 1 << rangebits]
            seen = set()
            for c in classes:
                if len(x_list) + dims[c] > 65:
                    raise ReaderError(DECODE, "x_list too long")
                for _ in range(dims[c]):
                    x = bs.read(rangebits)
                    if x in seen:
                        raise ReaderError(DECODE, "x_list not unique")
                    seen.add(x)
                    x_list.append(x)
            nb = [_find_neighbors(x_list, i) for i in range(len(x_list))]
            order = sorted(range(len(x_list)), key=lambda i: x_list[i])
            floors.append(dict(type=1, multiplier=mult, x_list=x_list, low=[a for a, _ in nb], high=[b for _, b in nb], sort_order=order,
                               partition_class=classes, classes=class_info))
        else:
            raise ReaderError(DECODE, "floor type")
    n_res = bs.read(6) + 1
    residues = []
    for _ in range(n_res):
        rtype = bs.read(16)
        if rtype > 2:
            raise ReaderError(DECODE, "residue type")
        begin, end = bs.read(24), bs.read(24)
        part_size = bs.read(24) + 1
        ncls = bs.read(6) + 1
        classbook = bs.read(8)
        if classbook >= max_book:
            raise ReaderError(DECODE, "classbook")
        if end < begin:
            raise ReaderError(DECODE, "residue range")
        used = []
        for _ in range(ncls):
            low = bs.read(3)
            high = bs.read(5) if bs.read_bool() else 0
            used.append(((high << 3) & 0xFF) | low)
        books = [[0] * 8 for _ in used]
        max_pass = 0
        for ci, u in enumerate(used):
            for j in range(8):
                if u >> j & 1:
                    book = bs.read(8)
                    if book == 0 or book >= max_book:
                        raise ReaderError(DECODE, "residue book")
                    books[ci][j] = book
                    max_pass = max(max_pass, j)
        residues.append(dict(type=rtype, begin=begin, end=end, partition_size=part_size, classifications=ncls, classbook=classbook, used=used, books=books,
                             max_pass=max_pass))
    mappings = []
    n_floors = len(floors) & 0xFF
    for _ in range(bs.read(6) + 1):
        if bs.read(16) != 0:
            raise ReaderError(DECODE, "mapping type")
        submaps = bs.read(4) + 1 if bs.read_bool() else 1
        couplings = []
        if bs.read_bool():
            steps = bs.read(8) + 1
            width = ilog(ident["n_channels"] - 1)
            for _ in range(steps):
                mag, ang = bs.read(width), bs.read(width)
                if mag == ang or mag > ident["n_channels"] - 1 or ang > ident["n_channels"] - 1:
                    raise ReaderError(DECODE, "coupling")
                couplings.append((mag, ang))
        if bs.read(2) != 0:
            raise ReaderError(DECODE, "reserved")
        mux = [0] * ident["n_channels"]
        if submaps > 1:
            for c in range(ident["n_channels"]):
                mux[c] = bs.read(4)
                if mux[c] >= submaps:
                    raise ReaderError(DECODE, "multiplex")
        sm = []
        for _ in range(submaps):
            bs.read(8)
            fl = bs.read(8)
            if fl >= n_floors:
                raise ReaderError(DECODE, "submap floor")
            rs = bs.read(8)
            if rs >= n_res:
                raise ReaderError(DECODE, "submap residue")
            sm.append((fl, rs))
        mappings.append(dict(couplings=couplings, multiplex=mux, submaps=sm))
    modes = []
    for _ in range(bs.read(6) + 1):
        flag = bs.read_bool()
        window, transform, mapping = bs.read(16), bs.read(16), bs.read(8)
        if window != 0 or transform != 0 or mapping >= len(mappings):
            raise ReaderError(DECODE, "mode")
        modes.append((flag, mapping))
    if not bs.read_bool():
        raise ReaderError(DECODE, "framing")
    return dict(n_codebooks=n_books, floors=floors, n_residues=n_res, residues=residues, mappings=mappings, modes=modes)
