"""A Layer III bitstream WRITER for the front-end tests: packs chosen side information, scale factors and quantised
spectral values into MPEG audio frames with a working bit reservoir (ISO/IEC 11172-3 2.4.1.7 / 2.4.2.7 and
13818-3 2.4.3.2), and keeps what it packed as ground truth.  It is not an encoder -- values are drawn at random --
but every stream it emits is one a decoder must take, and it shares no code with either reader.  Huffman (code,
length) pairs come from tests/golden/mp3_huffman.json."""
import json
import os

import numpy as np

from tests import _streams as st

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mp3_huffman.json")) as _f:
    HUFF = json.load(_f)
LINBITS = [0] * 16 + [1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13]
TABLE_OF = {**{t: str(t) for t in (1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15)}, **{t: "16" for t in range(16, 24)}, **{t: "24" for t in range(24, 32)}}
# scale factor band edges of long blocks, ISO 11172-3 Table B.8 / 13818-3 Table B.2 (+ the MPEG-2.5 extension), by
# sampling frequency 44.1 48 32 22.05 24 16 11.025 12 8 kHz
SFB_LONG = [
    [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 52, 62, 74, 90, 110, 134, 162, 196, 238, 288, 342, 418, 576],
    [0, 4, 8, 12, 16, 20, 24, 30, 36, 42, 50, 60, 72, 88, 106, 128, 156, 190, 230, 276, 330, 384, 576],
    [0, 4, 8, 12, 16, 20, 24, 30, 36, 44, 54, 66, 82, 102, 126, 156, 194, 240, 296, 364, 448, 550, 576],
    [0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576],
    [0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 114, 136, 162, 194, 232, 278, 332, 394, 464, 540, 576],
    [0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576],
    [0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576],
    [0, 6, 12, 18, 24, 30, 36, 44, 54, 66, 80, 96, 116, 140, 168, 200, 238, 284, 336, 396, 464, 522, 576],
    [0, 12, 24, 36, 48, 60, 72, 88, 108, 132, 160, 192, 232, 280, 336, 400, 476, 566, 568, 570, 572, 574, 576],
]
SLEN_MPEG1 = [(0, 0), (0, 1), (0, 2), (0, 3), (3, 0), (1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 3), (4, 2), (4, 3)]


class BitWriterMsb:
    def __init__(self):
        self.v = 0
        self.n = 0

    def put(self, value, width):
        assert 0 <= value < (1 << width) or width == 0, (value, width)
        self.v = (self.v << width) | value
        self.n += width

    def extend(self, other):
        self.v = (self.v << other.n) | other.v
        self.n += other.n

    def bytes(self):
        pad = (-self.n) % 8
        return ((self.v << pad).to_bytes((self.n + pad) // 8, "big")) if self.n else b""


def _mpeg2_partition(sfc, intensity_channel, block):
    """slen[4] and band counts[4] of an MPEG-2 scalefac_compress value, ISO 13818-3 2.4.3.2 (the two tables of
    `nr_of_sfb_block`), block = 0 long, 1 short, 2 mixed."""
    if intensity_channel:
        half = sfc >> 1
        if half < 180:
            slen, row = (half // 36, (half % 36) // 6, half % 6, 0), ((7, 7, 7, 0), (12, 12, 12, 0), (6, 15, 12, 0))
        elif half < 244:
            k = half - 180
            slen, row = ((k >> 4) & 3, (k >> 2) & 3, k & 3, 0), ((6, 6, 6, 3), (12, 9, 9, 6), (6, 12, 9, 6))
        else:
            k = half - 244
            slen, row = (k // 3, k % 3, 0, 0), ((8, 8, 5, 0), (15, 12, 9, 0), (6, 18, 9, 0))
        preflag = False
    else:
        if sfc < 400:
            slen, row = ((sfc >> 4) // 5, (sfc >> 4) % 5, (sfc & 15) >> 2, sfc & 3), ((6, 5, 5, 5), (9, 9, 9, 9), (6, 9, 9, 9))
        elif sfc < 500:
            k = sfc - 400
            slen, row = ((k >> 2) // 5, (k >> 2) % 5, k & 3, 0), ((6, 5, 7, 3), (9, 9, 12, 6), (6, 9, 12, 6))
        else:
            k = sfc - 500
            slen, row = (k // 3, k % 3, 0, 0), ((11, 10, 0, 0), (18, 18, 0, 0), (15, 18, 0, 0))
        preflag = sfc >= 500
    return slen, row[block], preflag


def gen_granule_channel(rng, version, rate_idx9, budget_bits, gr, scfsi, gr0_scalefacs, intensity_channel, rich=True, like=None, force_sfc=None):
    """One granule-channel: returns a dict with the side-information fields, `bits` (a BitWriterMsb holding part 2 +
    part 3 + stuffing) and the ground truth (`scalefacs`, `quant`, `rzero`, `preflag`)."""
    mpeg1 = version == "1"
    g = dict(global_gain=int(rng.integers(256)), scalefac_scale=int(rng.integers(2)), count1table=int(rng.integers(2)),
             subblock_gain=[0, 0, 0], window_switching=int(rng.integers(5) < 2), block_type=0, mixed_bit=0)
    if g["window_switching"]:
        g["block_type"] = int(rng.integers(1, 4))
        g["mixed_bit"] = int(rng.integers(2))
        g["subblock_gain"] = [int(x) for x in rng.integers(0, 8, 3)]
    if like is not None:  # joint stereo wants one window sequence for the pair (ISO 11172-3 2.4.3.4.10)
        g.update(window_switching=like["window_switching"], block_type=like["block_type"], mixed_bit=like["mixed_bit"])
        g["subblock_gain"] = [int(x) for x in rng.integers(0, 8, 3)] if g["window_switching"] else [0, 0, 0]
    short = g["block_type"] == 2
    mixed = short and g["mixed_bit"] == 1
    bits = BitWriterMsb()
    scalefacs = [0] * 39
    # ---- part 2
    if mpeg1:
        g["scalefac_compress"] = int(rng.integers(16))
        g["preflag"] = int(rng.integers(2))
        s1, s2 = SLEN_MPEG1[g["scalefac_compress"]]
        if short:
            n1 = 17 if mixed else 18
            for i in range(n1 + 18):
                s = s1 if i < n1 else s2
                if s:
                    scalefacs[i] = int(rng.integers(1 << s))
                    bits.put(scalefacs[i], s)
        else:
            for grp, (a, b) in enumerate(((0, 6), (6, 11), (11, 16), (16, 21))):
                s = s1 if grp < 2 else s2
                if gr == 1 and scfsi[grp]:
                    scalefacs[a:b] = gr0_scalefacs[a:b]
                elif s:
                    for i in range(a, b):
                        scalefacs[i] = int(rng.integers(1 << s))
                        bits.put(scalefacs[i], s)
    else:
        g["scalefac_compress"] = int(rng.integers(512)) if force_sfc is None else int(force_sfc)
        slen, counts, preflag = _mpeg2_partition(g["scalefac_compress"], intensity_channel, 2 if mixed else 1 if short else 0)
        g["preflag"] = int(preflag)
        at = 0
        for s, n in zip(slen, counts):
            for i in range(at, at + n):
                if s:
                    scalefacs[i] = int(rng.integers(1 << s))
                    bits.put(scalefacs[i], s)
            at += n
    part2 = bits.n
    # ---- regions
    if g["window_switching"]:
        g["table_select"] = [int(rng.integers(32)), int(rng.integers(32)), 0]
        if version == "2.5":
            r1 = SFB_LONG[rate_idx9][6 if (short and not mixed) else 8]
        elif mpeg1 or short:
            r1 = 36
        else:
            r1 = 54
        r2 = 576
        g["region0_count"] = g["region1_count"] = 0
    else:
        g["table_select"] = [int(rng.integers(32)) for _ in range(3)]
        g["region0_count"], g["region1_count"] = int(rng.integers(16)), int(rng.integers(8))
        r1 = SFB_LONG[rate_idx9][g["region0_count"] + 1]
        k = g["region0_count"] + g["region1_count"] + 2
        r2 = SFB_LONG[rate_idx9][k] if k <= 22 else 576
    if not rich:  # favour small tables: short codes, many values per bit
        g["table_select"] = [int(rng.choice([1, 2, 3, 5, 6, 7, 0])) for _ in range(3)]
    # ---- part 3
    quant = [0] * 576
    room = max(min(budget_bits, 4095) - part2, 0)
    target_pairs = int(rng.integers(0, 289)) if room > 0 else 0
    pairs = 0
    coded = BitWriterMsb()
    while pairs < target_pairs:
        i = 2 * pairs
        sel = g["table_select"][0 if i < r1 else 1 if i < r2 else 2]
        name = TABLE_OF.get(sel)
        if name is None:  # tables 0, 4, 14: the pair is (0, 0) and takes no bits
            pairs += 1
            continue
        t = HUFF[name]
        wrap, lin = t["wrap"], LINBITS[sel]
        xy = [int(min(rng.geometric(0.35) - 1, wrap - 1)) if rng.integers(4) else int(rng.integers(wrap)) for _ in range(2)]
        one = BitWriterMsb()
        idx = xy[0] * wrap + xy[1]
        one.put(t["codes"][idx], t["lens"][idx])
        for k in range(2):
            x = xy[k]
            if x == 0:
                continue
            if x == 15 and lin:
                extra = int(rng.integers(1 << lin)) if rng.integers(3) else (1 << lin) - 1
                one.put(extra, lin)
                x += extra
            sign = int(rng.integers(2))
            one.put(sign, 1)
            quant[i + k] = -x if sign else x
        if coded.n + one.n > room:
            quant[i] = quant[i + 1] = 0
            break
        coded.extend(one)
        pairs += 1
    g["big_values"] = pairs
    i = 2 * pairs
    quads = 0
    t = HUFF["quadB" if g["count1table"] else "quadA"]
    want_quads = int(rng.integers(0, (576 - i) // 4 + 1)) if room - coded.n > 0 else 0
    while quads < want_quads and i <= 572:
        flags = int(rng.integers(16)) if rng.integers(3) else int(rng.choice([0, 1, 2, 4, 8]))
        one = BitWriterMsb()
        one.put(t["codes"][flags], t["lens"][flags])
        vals = [0, 0, 0, 0]
        for k in range(4):  # v, w, x, y = bits 3..0; their sign bits follow in that order
            if flags & (8 >> k):
                sign = int(rng.integers(2))
                one.put(sign, 1)
                vals[k] = -1 if sign else 1
        if coded.n + one.n > room:
            break
        coded.extend(one)
        quant[i:i + 4] = vals
        i += 4
        quads += 1
    rzero = i if coded.n > 0 else 0  # no part-3 bits at all: the reader returns before looking at big_values
    if coded.n == 0:
        quant = [0] * 576
    # stuffing is only safe where the quad loop cannot run any more
    stuffing = 0
    if coded.n > 0 and i > 572 and room - coded.n > 0 and rng.integers(2):
        stuffing = int(rng.integers(1, min(room - coded.n, 40) + 1))
        coded.put(int(rng.integers(1 << stuffing)), stuffing)
    bits.extend(coded)
    g["part2_3_length"] = bits.n
    g.update(bits=bits, scalefacs=scalefacs, quant=quant, rzero=rzero, short=short, mixed=mixed, stuffing=stuffing)
    return g


def side_info_bytes(version, n_ch, main_data_begin, scfsi, granules):
    """ISO 11172-3 2.4.1.7 / 13818-3 2.4.1.7."""
    w = BitWriterMsb()
    mpeg1 = version == "1"
    if mpeg1:
        w.put(main_data_begin, 9)
        w.put(0, 5 if n_ch == 1 else 3)
        for ch in range(n_ch):
            for b in range(4):
                w.put(int(scfsi[ch][b]), 1)
    else:
        w.put(main_data_begin, 8)
        w.put(0, 1 if n_ch == 1 else 2)
    for gr in granules:
        for g in gr:
            w.put(g["part2_3_length"], 12), w.put(g["big_values"], 9), w.put(g["global_gain"], 8)
            w.put(g["scalefac_compress"], 4 if mpeg1 else 9)
            w.put(g["window_switching"], 1)
            if g["window_switching"]:
                w.put(g["block_type"], 2), w.put(g["mixed_bit"], 1)
                w.put(g["table_select"][0], 5), w.put(g["table_select"][1], 5)
                for k in range(3):
                    w.put(g["subblock_gain"][k], 3)
            else:
                for k in range(3):
                    w.put(g["table_select"][k], 5)
                w.put(g["region0_count"], 4), w.put(g["region1_count"], 3)
            if mpeg1:
                w.put(g["preflag"], 1)
            w.put(g["scalefac_scale"], 1), w.put(g["count1table"], 1)
    out = w.bytes()
    want = (17 if n_ch == 1 else 32) if mpeg1 else (9 if n_ch == 1 else 17)
    assert len(out) == want, (len(out), want)
    return out


def gen_stream(rng, n_frames, version="1", mode=1, rate_idx=0, bitrate_idx=9, protected=False, fill=(0.3, 1.0), padding=None, rich=True,
               pair_blocks=False, force_sfc=None, force_mode_ext=None):
    """A stream of frames with a bit reservoir.  Returns (list of frame bytes, list of per-frame truth dicts)."""
    n_ch = 1 if mode == 3 else 2
    n_gr = 2 if version == "1" else 1
    rate_idx9 = rate_idx + {"1": 0, "2": 3, "2.5": 6}[version]
    side_len = (17 if n_ch == 1 else 32) if version == "1" else (9 if n_ch == 1 else 17)
    limit = 511 if version == "1" else 255
    frames, truth = [], []
    slack = 0            # bytes before this frame's slot that its main data may use
    stream_slots = []    # per frame: (header + CRC + side information, slot size)
    payload = bytearray()  # the concatenated slots
    for k in range(n_frames):
        pad = int(rng.integers(2)) if padding is None else padding
        mode_ext = int(rng.integers(4)) if force_mode_ext is None else int(force_mode_ext(k))
        word = st.mpa_word(version=version, layer=3, bitrate_idx=bitrate_idx, rate_idx=rate_idx, mode=mode, mode_ext=mode_ext, padding=pad,
                           protected=protected, copyright=int(rng.integers(2)), original=int(rng.integers(2)), emphasis=int(rng.integers(4)))
        total = st.mpa_frame_len(version, 3, bitrate_idx, rate_idx, pad)
        slot = total - 4 - (2 if protected else 0) - side_len
        begin = min(slack, limit)
        stuffing_bytes = slack - begin  # reservoir the frame may not reach back to: ancillary bytes nobody reads
        budget = 8 * (begin + slot)
        intensity = mode == 1 and bool(mode_ext & 1)
        scfsi = [[int(rng.integers(2)) for _ in range(4)] for _ in range(2)]
        use = int(budget * rng.uniform(*fill))
        while True:  # part 2 is written whatever the budget says: shrink until the frame's main data fits
            granules = []
            shares = rng.dirichlet(np.ones(n_gr * n_ch)) * use
            for gr in range(n_gr):
                row = []
                for ch in range(n_ch):
                    row.append(gen_granule_channel(rng, version, rate_idx9, int(shares[gr * n_ch + ch]), gr, scfsi[ch],
                                                   granules[0][ch]["scalefacs"] if gr == 1 else None, ch == 1 and intensity, rich=rich,
                                                   like=row[0] if (pair_blocks and ch == 1) else None,
                                                   force_sfc=None if force_sfc is None else force_sfc(k, gr, ch)))
                granules.append(row)
            md = BitWriterMsb()
            for row in granules:
                for g in row:
                    md.extend(g["bits"])
            md_bytes = md.bytes()
            if len(md_bytes) <= begin + slot:
                break
            use //= 2
        payload += bytes(rng.integers(0, 256, stuffing_bytes, dtype=np.uint8)) + md_bytes
        slack = begin + slot - len(md_bytes)
        side = side_info_bytes(version, n_ch, begin, scfsi, granules)
        stream_slots.append((word.to_bytes(4, "big") + (bytes(rng.integers(0, 256, 2, dtype=np.uint8)) if protected else b"") + side, slot))
        truth.append(dict(word=word, version=version, n_ch=n_ch, n_gr=n_gr, rate_idx9=rate_idx9, mode=mode, mode_ext=mode_ext, granules=granules, scfsi=scfsi,
                          main_data_begin=begin, main_data_bytes=len(md_bytes)))
    # pour the payload into the slots
    payload = bytes(payload) + bytes(sum(s for _, s in stream_slots) - len(payload))
    at = 0
    for head, slot in stream_slots:
        frames.append(head + payload[at:at + slot])
        at += slot
    return frames, truth
