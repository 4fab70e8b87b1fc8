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

def adts_frame(rng, payload_len, rate_idx=4, channels=2, profile=1, protected=False, mpeg2=False, blocks=0, frame_len=None):
    """ISO 13818-7 6.2 adts_fixed_header + adts_variable_header (+ crc), then a random payload."""
    hlen = 9 if protected else 7
    total = hlen + payload_len if frame_len is None else frame_len
    bits = (0xFFF << 44) | (int(mpeg2) << 43) | (0 << 41) | ((0 if protected else 1) << 40) | (profile << 38) | (rate_idx << 34) | \
        (int(rng.integers(2)) << 33) | (channels << 30) | (int(rng.integers(16)) << 26) | (total << 13) | (int(rng.integers(0x800)) << 2) | blocks
    h = bits.to_bytes(7, "big")
    if protected:
        h += bytes(rng.integers(0, 256, 2, dtype=np.uint8))
    return h + rng.integers(0, 256, payload_len, dtype=np.uint8).tobytes()


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


def ogg_paginate(serial, packets, rng, max_segments=255, first_sequence=0, granule_step=1024, bos=True, eos=True):
    """Pack packets into pages the way a muxer does: lacing values of 255 continue, a packet may spill over any number
    of pages, pages close at random fill levels.  Returns the list of page byte strings."""
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
    while i < len(segs):
        take = int(rng.integers(1, max_segments + 1))
        chunk = segs[i:i + take]
        i += len(chunk)
        ends = sum(1 for s in chunk if s[2])
        absgp += ends * granule_step
        pages.append(ogg_page(serial, seq, absgp if ends else 0xFFFFFFFFFFFFFFFF, [s[0] for s in chunk], b"".join(s[1] for s in chunk),
                              continuation=open_packet, first=bos and seq == first_sequence, last=eos and i >= len(segs)))
        open_packet = not chunk[-1][2]
        seq += 1
    return pages
