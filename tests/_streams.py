"""Synthetic byte streams for the packetiser tests: MPEG audio frames (with Xing / Info / LAME / VBRI tag frames),
ADTS frames and Ogg pages, with the damage real files show -- junk between frames, false sync words, truncated
tails, bad checksums, lost pages.  Builders only: nothing here parses, so the C++ index builders and the oracle are
both checked against bytes neither of them produced."""
import numpy as np

from oracle import packetizer_oracle as po

_KBPS = {("1", 1): [32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448],
         ("1", 2): [32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384],
         ("1", 3): [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320],
         ("2", 1): [32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256],
         ("2", 23): [8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]}
_RATES = {"1": [44100, 48000, 32000], "2": [22050, 24000, 16000], "2.5": [11025, 12000, 8000]}


def mpa_word(version="1", layer=3, bitrate_idx=9, rate_idx=0, mode=0, mode_ext=0, padding=0, protected=False, private=0,
             copyright=0, original=0, emphasis=0):
    """The 32-bit header word from its fields (ISO 11172-3 2.4.1.3 bit layout)."""
    v = {"1": 3, "2": 2, "2.5": 0}[version]
    return (0x7FF << 21) | (v << 19) | ((4 - layer) << 17) | ((0 if protected else 1) << 16) | (bitrate_idx << 12) | (rate_idx << 10) | \
        (padding << 9) | (private << 8) | (mode << 6) | (mode_ext << 4) | (copyright << 3) | (original << 2) | emphasis


def mpa_frame_len(version, layer, bitrate_idx, rate_idx, padding):
    """Whole frame in bytes, by the standard's slot formula (ISO 11172-3 2.4.3.1 / 13818-3 2.4.3.1)."""
    key = ("1", layer) if version == "1" else ("2", 1 if layer == 1 else 23)
    bitrate = _KBPS[key][bitrate_idx - 1] * 1000
    rate = _RATES[version][rate_idx]
    if layer == 1:
        return (12 * bitrate // rate + padding) * 4
    if layer == 3 and version != "1":
        return 72 * bitrate // rate + padding
    return 144 * bitrate // rate + padding


def _layer2_allowed(bitrate_idx, mono):
    k = _KBPS[("1", 2)][bitrate_idx - 1]
    return k not in (224, 256, 320, 384) if mono else k not in (32, 48, 56, 80)


def mpa_random_params(rng, version=None, layer=None, mode=None):
    version = version or ["1", "2", "2.5"][rng.integers(3)]
    layer = layer or int(rng.integers(1, 4))
    mode = int(rng.integers(4)) if mode is None else mode
    while True:
        bi = int(rng.integers(1, 15))
        if layer == 2 and version == "1" and not _layer2_allowed(bi, mode == 3):
            continue
        if layer == 2 and version != "1":
            # the reference applies the Layer II table of forbidden rates to the halved-rate versions too, by VALUE
            k = _KBPS[("2", 23)][bi - 1]
            if (k in (224, 256, 320, 384)) if mode == 3 else (k in (32, 48, 56, 80)):
                continue
        break
    return dict(version=version, layer=layer, bitrate_idx=bi, rate_idx=int(rng.integers(3)), mode=mode)


def mpa_frame(rng, params, padding=None, protected=None, body=None, mode_ext=None):
    """One frame with a random body (which may well contain bytes that look like sync words)."""
    padding = int(rng.integers(2)) if padding is None else padding
    protected = bool(rng.integers(4) == 0) if protected is None else protected
    w = mpa_word(padding=padding, protected=protected, mode_ext=int(rng.integers(4)) if mode_ext is None else mode_ext,
                 copyright=int(rng.integers(2)), original=int(rng.integers(2)), emphasis=int(rng.integers(4)), **params)
    n = mpa_frame_len(params["version"], params["layer"], params["bitrate_idx"], params["rate_idx"], padding)
    if body is None:
        body = rng.integers(0, 256, n - 4, dtype=np.uint8).tobytes()
    assert len(body) == n - 4
    return w.to_bytes(4, "big") + body


def mpa_junk(rng, n):
    """Random bytes salted with 0xff runs and almost-headers (bad version / layer / bit-rate / rate fields, free format)."""
    b = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
    for _ in range(n // 24):
        at = int(rng.integers(0, max(n - 4, 1)))
        kind = int(rng.integers(6))
        w = mpa_word(**mpa_random_params(rng))
        if kind == 0:
            w = (w & ~(3 << 19)) | (1 << 19)          # reserved version
        elif kind == 1:
            w &= ~(3 << 17)                             # reserved layer
        elif kind == 2:
            w |= 15 << 12                               # forbidden bit-rate
        elif kind == 3:
            w |= 3 << 10                                # reserved sample rate
        elif kind == 4:
            w &= ~(15 << 12)                            # free format: passes the quick check, fails the parse
        else:
            w = 0xFFFFFFFF
        b[at:at + 4] = w.to_bytes(4, "big")[:max(0, min(4, n - at))]
    return bytes(b)


def mpa_tag_frame(rng, params, kind="Xing", flags=0xF, num_frames=1000, num_bytes=400000, quality=57, lame=b"LAME3.100",
                  lame_ext=36, delay=576, padding=1000, crc="good", protected=False, side_info_noise=False, vbri_version=1):
    """A Layer III frame holding a Xing / Info (+ LAME extension) or VBRI tag.  `crc`: good | bad | zero;
    `lame_ext`: how many bytes of the 36-byte extension to write (0 = none; < 36 = truncated)."""
    assert params["layer"] == 3
    w = mpa_word(padding=0, protected=protected, **params)
    n = mpa_frame_len(params["version"], 3, params["bitrate_idx"], params["rate_idx"], 0)
    f = bytearray(n)
    f[0:4] = w.to_bytes(4, "big")
    mono = params["mode"] == 3
    side = (17 if mono else 32) if params["version"] == "1" else (9 if mono else 17)
    if protected:
        f[4:6] = bytes(rng.integers(1, 256, 2, dtype=np.uint8))  # the CRC bytes are not part of the "zero side info" test
    if side_info_noise:
        f[4 + (2 if protected else 0) + 1] = 0x5A
    if kind == "VBRI":
        at = 36
        body = b"VBRI" + int(vbri_version).to_bytes(2, "big") + (0).to_bytes(2, "big") + (75).to_bytes(2, "big") + \
            int(num_bytes).to_bytes(4, "big") + int(num_frames).to_bytes(4, "big") + bytes(8)
        f[at:at + len(body)] = body[:max(0, n - at)]
        return bytes(f)
    at = 4 + side
    tag = bytearray(kind.encode() + int(flags).to_bytes(4, "big"))
    if flags & 1:
        tag += int(num_frames).to_bytes(4, "big")
    if flags & 2:
        tag += int(num_bytes).to_bytes(4, "big")
    if flags & 4:
        tag += bytes(int(x) for x in np.linspace(0, 255, 100))
    if flags & 8:
        tag += int(quality).to_bytes(4, "big")
    if lame_ext:
        ext = bytearray(36)
        ext[0:9] = lame[:9].ljust(9, b" ")
        ext[9], ext[10] = 0x04, 190
        ext[11:15] = (0x00400000).to_bytes(4, "big")
        ext[15:17], ext[17:19] = (0x2C3C).to_bytes(2, "big"), (0x4A14).to_bytes(2, "big")
        ext[19], ext[20] = 0x24, 128
        ext[21:24] = ((delay << 12) | padding).to_bytes(3, "big")
        ext[24], ext[25] = 0x01, 0
        ext[26:28] = (0).to_bytes(2, "big")
        ext[28:32] = int(num_bytes).to_bytes(4, "big")
        ext[32:34] = (0xBEEF).to_bytes(2, "big")
        tag += ext[:lame_ext]
    if at + len(tag) > n:
        tag = tag[:n - at]
    f[at:at + len(tag)] = tag
    if lame_ext >= 36 and at + len(tag) <= n:
        end = at + len(tag) - 2
        good = po.crc16_ansi_le_update(0, bytes(f[:end]))
        val = {"good": good, "bad": good ^ 0x1234 or 1, "zero": 0}[crc]
        f[end:end + 2] = val.to_bytes(2, "big")
    return bytes(f)


# ---------------------------------------------------------------------------------------------------- ADTS

def adts_frame(rng, payload_len, rate_idx=4, channels=2, profile=1, protected=False, mpeg2=False, blocks=0, frame_len=None, payload=None):
    """ISO 13818-7 6.2 adts_fixed_header + adts_variable_header (+ crc), then a random payload (or `payload`)."""
    if payload is not None:
        payload_len = len(payload)
    hlen = 9 if protected else 7
    total = hlen + payload_len if frame_len is None else frame_len
    bits = (0xFFF << 44) | (int(mpeg2) << 43) | (0 << 41) | ((0 if protected else 1) << 40) | (profile << 38) | (rate_idx << 34) | \
        (int(rng.integers(2)) << 33) | (channels << 30) | (int(rng.integers(16)) << 26) | (total << 13) | (int(rng.integers(0x800)) << 2) | blocks
    h = bits.to_bytes(7, "big")
    if protected:
        h += bytes(rng.integers(0, 256, 2, dtype=np.uint8))
    return h + (rng.integers(0, 256, payload_len, dtype=np.uint8).tobytes() if payload is None else bytes(payload))


# ---------------------------------------------------------------------------------------------------- Ogg

def ogg_page(serial, sequence, absgp, lacing, body, continuation=False, first=False, last=False, version=0, flag_noise=0, bad_crc=False):
    """RFC 3533 section 6 page; the checksum is computed with the oracle's bitwise CRC."""
    assert len(lacing) <= 255 and sum(lacing) == len(body)
    flags = int(continuation) | (int(first) << 1) | (int(last) << 2) | flag_noise
    h = b"OggS" + bytes([version, flags]) + int(absgp).to_bytes(8, "little") + int(serial).to_bytes(4, "little") + \
        int(sequence).to_bytes(4, "little") + bytes(4) + bytes([len(lacing)]) + bytes(lacing)
    crc = po.crc32_update(0, h + body)
    if bad_crc:
        crc ^= 0x00010000
    return h[:22] + crc.to_bytes(4, "little") + h[26:] + body


def ogg_paginate(serial, packets, rng, max_segments=255, first_sequence=0, granule_step=1024, bos=True, eos=True, granule_of=None):
    """Pack packets into pages the way a muxer does: lacing values of 255 continue, a packet may spill over any number
    of pages, pages close at random fill levels.  Returns the list of page byte strings.  granule_of[k] (optional): the granule
    position once packet k is complete; a page then carries that of the last packet ending on it."""
    segs = []  # (lacing value, bytes, closes-a-packet)
    for p in packets:
        n = len(p)
        at = 0
        while n - at >= 255:
            segs.append((255, p[at:at + 255], False))
            at += 255
        segs.append((n - at, p[at:], True))
    pages, seq, absgp, i = [], first_sequence, 0, 0
    open_packet = False
    done = 0
    while i < len(segs):
        take = int(rng.integers(1, max_segments + 1))
        chunk = segs[i:i + take]
        i += len(chunk)
        ends = sum(1 for s in chunk if s[2])
        absgp += ends * granule_step
        done += ends
        if granule_of is not None and ends:
            absgp = granule_of[done - 1]
        pages.append(ogg_page(serial, seq, absgp if ends else 0xFFFFFFFFFFFFFFFF, [s[0] for s in chunk], b"".join(s[1] for s in chunk),
                              continuation=open_packet, first=bos and seq == first_sequence, last=eos and i >= len(segs)))
        open_packet = not chunk[-1][2]
        seq += 1
    return pages


# ---------------------------------------------------------------------------------------------------- Vorbis headers

class BitWriterRtl:
    """Vorbis bit packing (Vorbis I specification section 2): values go in least-significant bit first."""

    def __init__(self):
        self.v = 0
        self.n = 0

    def put(self, value, width):
        assert 0 <= value < (1 << width) or width == 0
        self.v |= value << self.n
        self.n += width

    def bytes(self):
        return self.v.to_bytes((self.n + 7) // 8, "little")


def vorbis_ident(channels=2, rate=44100, bs0=8, bs1=11, version=0, framing=1, sig=b"vorbis", ptype=1):
    """Vorbis I 4.2.2."""
    return bytes([ptype]) + sig + version.to_bytes(4, "little") + bytes([channels]) + rate.to_bytes(4, "little") + \
        (0).to_bytes(4, "little") + (128000).to_bytes(4, "little") + (0).to_bytes(4, "little") + bytes([(bs1 << 4) | bs0, framing])


def _ilog(x):
    return x.bit_length()


def _put_codebook(w, rng, style=None, lookup=None, sync=0x564342):
    """Vorbis I 3.2.1.  `style`: plain | sparse | ordered."""
    style = style or ["plain", "sparse", "ordered"][int(rng.integers(3))]
    dims = int(rng.integers(1, 9))
    entries = int(rng.integers(1, 400))
    w.put(sync, 24), w.put(dims, 16), w.put(entries, 24)
    if style == "ordered":
        w.put(1, 1)
        w.put(int(rng.integers(32)), 5)
        cur = 0
        while cur < entries:
            num = int(rng.integers(0, entries - cur + 1))
            w.put(num, _ilog(entries - cur))
            cur += num
    else:
        w.put(0, 1)
        w.put(int(style == "sparse"), 1)
        for _ in range(entries):
            if style == "sparse":
                used = int(rng.integers(2))
                w.put(used, 1)
                if used:
                    w.put(int(rng.integers(32)), 5)
            else:
                w.put(int(rng.integers(32)), 5)
    lookup = int(rng.integers(3)) if lookup is None else lookup
    w.put(lookup, 4)
    if lookup in (1, 2):
        w.put(int(rng.integers(1 << 32)), 32), w.put(int(rng.integers(1 << 32)), 32)
        value_bits = int(rng.integers(1, 17))
        w.put(value_bits - 1, 4), w.put(int(rng.integers(2)), 1)
        if lookup == 1:
            n = 0
            while (n + 1) ** dims <= entries:
                n += 1
        else:
            n = entries * dims
        for _ in range(n):
            w.put(int(rng.integers(1 << value_bits)), value_bits)


def _put_floor(w, rng, kind=None):
    """Vorbis I 6.2.1 / 7.2.2."""
    kind = int(rng.integers(2)) if kind is None else kind
    w.put(kind, 16)
    if kind == 0:
        w.put(int(rng.integers(256)), 8), w.put(int(rng.integers(65536)), 16), w.put(int(rng.integers(65536)), 16)
        w.put(int(rng.integers(64)), 6), w.put(int(rng.integers(256)), 8)
        books = int(rng.integers(1, 17))
        w.put(books - 1, 4)
        for _ in range(books):
            w.put(int(rng.integers(256)), 8)
        return
    if kind != 1:
        return
    parts = int(rng.integers(0, 32))
    classes = [int(rng.integers(0, 16)) for _ in range(parts)]
    w.put(parts, 5)
    for c in classes:
        w.put(c, 4)
    dims = {}
    if parts:
        for c in range(max(classes) + 1):
            dims[c] = int(rng.integers(1, 9))
            sub = int(rng.integers(4))
            w.put(dims[c] - 1, 3), w.put(sub, 2)
            if sub:
                w.put(int(rng.integers(256)), 8)
            for _ in range(1 << sub):
                w.put(int(rng.integers(256)), 8)
    w.put(int(rng.integers(4)), 2)
    rangebits = int(rng.integers(16))
    w.put(rangebits, 4)
    for c in classes:
        for _ in range(dims[c]):
            w.put(int(rng.integers(1 << rangebits)) if rangebits else 0, rangebits)


def _put_residue(w, rng):
    """Vorbis I 8.6.1."""
    w.put(int(rng.integers(3)), 16)
    w.put(int(rng.integers(1 << 24)), 24), w.put(int(rng.integers(1 << 24)), 24), w.put(int(rng.integers(1 << 24)), 24)
    classes = int(rng.integers(1, 65))
    w.put(classes - 1, 6), w.put(int(rng.integers(256)), 8)
    books = 0
    for _ in range(classes):
        low = int(rng.integers(8))
        w.put(low, 3)
        high = 0
        if rng.integers(2):
            high = int(rng.integers(32))
            w.put(1, 1), w.put(high, 5)
        else:
            w.put(0, 1)
        books += bin((high << 3) | low).count("1")
    for _ in range(books):
        w.put(int(rng.integers(256)), 8)


def _put_mapping(w, rng, channels, kind=0, reserved=0):
    """Vorbis I 4.2.4 (mappings)."""
    w.put(kind, 16)
    submaps = 1
    if rng.integers(2):
        submaps = int(rng.integers(1, 17))
        w.put(1, 1), w.put(submaps - 1, 4)
    else:
        w.put(0, 1)
    if rng.integers(2):
        steps = int(rng.integers(1, 257))
        w.put(1, 1), w.put(steps - 1, 8)
        width = _ilog(channels - 1)
        for _ in range(steps):
            w.put(int(rng.integers(1 << width)) if width else 0, width), w.put(int(rng.integers(1 << width)) if width else 0, width)
    else:
        w.put(0, 1)
    w.put(reserved, 2)
    if submaps > 1:
        for _ in range(channels):
            w.put(int(rng.integers(16)), 4)
    for _ in range(submaps):
        w.put(0, 8), w.put(int(rng.integers(256)), 8), w.put(int(rng.integers(256)), 8)


def vorbis_setup(rng, channels=2, modes=None, fault=None, n_codebooks=None):
    """A random but well-formed setup header (Vorbis I 4.2.4) ending in the given mode block flags, or, with `fault`,
    one broken in a named place.  Returns (packet, block flags)."""
    modes = [bool(rng.integers(2)) for _ in range(int(rng.integers(1, 9)))] if modes is None else modes
    w = BitWriterRtl()
    n_codebooks = int(rng.integers(1, 12)) if n_codebooks is None else n_codebooks
    w.put(n_codebooks - 1, 8)
    for k in range(n_codebooks):
        _put_codebook(w, rng, sync=0x564343 if fault == "codebook_sync" and k == n_codebooks - 1 else 0x564342,
                      lookup=3 if fault == "lookup_type" and k == 0 else None)
    n = int(rng.integers(1, 4))
    w.put(n - 1, 6)
    for k in range(n):
        w.put(1 if fault == "time_domain" and k == n - 1 else 0, 16)
    n = int(rng.integers(1, 5))
    w.put(n - 1, 6)
    for k in range(n):
        _put_floor(w, rng, kind=2 if fault == "floor_type" and k == 0 else None)
    n = int(rng.integers(1, 5))
    w.put(n - 1, 6)
    for _ in range(n):
        _put_residue(w, rng)
    n = int(rng.integers(1, 4))
    w.put(n - 1, 6)
    for k in range(n):
        _put_mapping(w, rng, channels, kind=1 if fault == "mapping_type" and k == 0 else 0, reserved=2 if fault == "mapping_reserved" and k == n - 1 else 0)
    w.put(len(modes) - 1, 6)
    for k, flag in enumerate(modes):
        w.put(int(flag), 1)
        w.put(1 if fault == "window" and k == 0 else 0, 16)
        w.put(1 if fault == "transform" and k == len(modes) - 1 else 0, 16)
        w.put(int(rng.integers(256)), 8)
    w.put(0 if fault == "framing" else 1, 1)
    body = w.bytes()
    if fault == "truncated":
        body = body[:len(body) * 2 // 3]
    return (b"\x05vorbis" if fault != "signature" else b"\x05vorbiz") + body, modes


def vorbis_audio_packet(rng, n_modes, mode=None, n=None):
    """An audio packet: type bit 0, the mode number, then opaque bits."""
    w = BitWriterRtl()
    w.put(0, 1)
    mode = int(rng.integers(n_modes)) if mode is None else mode
    w.put(mode, _ilog(n_modes - 1))
    w.put(int(rng.integers(1 << 30)), 30)
    body = w.bytes()
    return body + rng.integers(0, 256, int(rng.integers(0, 400)) if n is None else n, dtype=np.uint8).tobytes(), mode


def vorbis_setup_valid(rng, channels=2, bs_exp=(8, 11), n_codebooks=None, floor_types=None, fault=None):
    """A setup header a DECODER accepts (Vorbis I 4.2.4 with every cross reference in range: codebook numbers, floor /
    residue / mapping indices, distinct X positions, distinct coupled channels), and what it says.  `fault` breaks one
    cross reference.  Returns (packet, truth dict)."""
    w = BitWriterRtl()
    n_books = int(rng.integers(2, 10)) if n_codebooks is None else n_codebooks
    w.put(n_books - 1, 8)
    for _ in range(n_books):
        _put_codebook(w, rng)
    w.put(0, 6), w.put(0, 16)
    n_floors = int(rng.integers(1, 5)) if floor_types is None else len(floor_types)
    w.put(n_floors - 1, 6)
    floors = []
    for fi in range(n_floors):
        kind = (int(rng.integers(5) > 0) if floor_types is None else floor_types[fi])
        w.put(kind, 16)
        if kind == 0:
            w.put(int(rng.integers(256)), 8), w.put(int(rng.integers(65536)), 16), w.put(int(rng.integers(65536)), 16)
            w.put(int(rng.integers(64)), 6), w.put(int(rng.integers(256)), 8)
            books = int(rng.integers(1, 17))
            w.put(books - 1, 4)
            for k in range(books):
                w.put(n_books if (fault == "floor0_book" and fi == 0 and k == 0) else int(rng.integers(n_books)), 8)
            floors.append(dict(type=0))
            continue
        rangebits = int(rng.integers(6, 13))
        parts = int(rng.integers(0, 12))
        classes = [int(rng.integers(0, 5)) for _ in range(parts)]
        dims = {c: int(rng.integers(1, 5)) for c in range(max(classes) + 1)} if parts else {}
        while 2 + sum(dims[c] for c in classes) > min(65, (1 << rangebits) - 1):
            classes.pop()
            parts -= 1
        w.put(parts, 5)
        for c in classes:
            w.put(c, 4)
        if parts:
            for c in range(max(classes) + 1):
                sub = int(rng.integers(4))
                w.put(dims[c] - 1, 3), w.put(sub, 2)
                if sub:
                    w.put(n_books if (fault == "floor1_mainbook" and fi == 0 and c == 0) else int(rng.integers(n_books)), 8)
                for _ in range(1 << sub):
                    w.put(int(rng.integers(0, n_books + 1)), 8)  # 0 = none, else book + 1
        mult = int(rng.integers(1, 5))
        w.put(mult - 1, 2), w.put(rangebits, 4)
        n_x = sum(dims[c] for c in classes)
        xs = [int(v) for v in rng.choice(np.arange(1, 1 << rangebits), size=n_x, replace=False)]
        if fault == "floor1_duplicate_x" and fi == 0 and n_x >= 2:
            xs[-1] = xs[0]
        for x in xs:
            w.put(x, rangebits)
        floors.append(dict(type=1, multiplier=mult, x_list=[0, 1 << rangebits] + xs, fault_applicable=n_x >= 2))
    n_res = int(rng.integers(1, 4))
    w.put(n_res - 1, 6)
    for ri in range(n_res):
        w.put(int(rng.integers(3)) if not (fault == "residue_type" and ri == 0) else 3, 16)
        begin = int(rng.integers(0, 1000))
        w.put(begin, 24), w.put(begin + int(rng.integers(0, 2000)) if not (fault == "residue_range" and ri == 0) else max(begin - 1, 0), 24)
        w.put(int(rng.integers(1, 64)), 24)
        ncls = int(rng.integers(1, 9))
        w.put(ncls - 1, 6), w.put(int(rng.integers(n_books)), 8)
        used = []
        for _ in range(ncls):
            low, high = int(rng.integers(8)), (int(rng.integers(32)) if rng.integers(2) else None)
            w.put(low, 3)
            if high is None:
                w.put(0, 1)
            else:
                w.put(1, 1), w.put(high, 5)
            used.append(((high or 0) << 3) | low)
        first = True
        for u in used:
            for j in range(8):
                if u >> j & 1:
                    w.put(0 if (fault == "residue_book_zero" and ri == 0 and first) else int(rng.integers(1, n_books)), 8)
                    first = False
    n_map = int(rng.integers(1, 4))
    w.put(n_map - 1, 6)
    mappings = []
    for mi in range(n_map):
        w.put(0, 16)
        submaps = int(rng.integers(1, 4)) if channels > 1 else 1
        if submaps > 1:
            w.put(1, 1), w.put(submaps - 1, 4)
        else:
            w.put(0, 1)
        couplings = []
        if channels > 1 and rng.integers(2):
            steps = int(rng.integers(1, 4))
            w.put(1, 1), w.put(steps - 1, 8)
            width = _ilog(channels - 1)
            for k in range(steps):
                a, b = (int(v) for v in rng.choice(channels, size=2, replace=False))
                if fault == "coupling_same" and mi == 0 and k == 0:
                    b = a
                w.put(a, width), w.put(b, width)
                couplings.append((a, b))
        else:
            w.put(0, 1)
        w.put(0, 2)
        mux = [0] * channels
        if submaps > 1:
            for c in range(channels):
                mux[c] = int(rng.integers(submaps)) if not (fault == "mux" and mi == 0 and c == 0) else submaps
                w.put(mux[c], 4)
        sm = []
        for k in range(submaps):
            fl = int(rng.integers(n_floors)) if not (fault == "submap_floor" and mi == 0 and k == 0) else n_floors
            rs = int(rng.integers(n_res)) if not (fault == "submap_residue" and mi == 0 and k == 0) else n_res
            w.put(int(rng.integers(256)), 8), w.put(fl, 8), w.put(rs, 8)
            sm.append((fl, rs))
        mappings.append(dict(couplings=couplings, multiplex=mux, submaps=sm))
    n_modes = int(rng.integers(1, 6))
    w.put(n_modes - 1, 6)
    modes = []
    for k in range(n_modes):
        flag, mp = int(rng.integers(2)), (int(rng.integers(n_map)) if not (fault == "mode_mapping" and k == 0) else n_map)
        w.put(flag, 1), w.put(0, 16), w.put(0, 16), w.put(mp, 8)
        modes.append((bool(flag), mp))
    w.put(1, 1)
    return b"\x05vorbis" + w.bytes(), dict(n_codebooks=n_books, floors=floors, n_residues=n_res, mappings=mappings, modes=modes)
