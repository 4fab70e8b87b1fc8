"""MP3 entropy front-end oracle: restates, in the reference's own sequence, how a Layer III packet
becomes side information, scale factors and Huffman-decoded spectral values.  TEST INFRASTRUCTURE ONLY.

  MpaDecoder::decode_inner        symphonia-bundle-mp3/src/decoder.rs:84-131
  Layer3::decode                  layer3/mod.rs:373-418
  BitResevoir                     layer3/mod.rs:34-108
  read_main_data                  layer3/mod.rs:272-370
  read_side_info & co.            layer3/bitstream.rs:57-236
  read_scale_factors_mpeg1/2      layer3/bitstream.rs:240-427
  read_huffman_samples            layer3/requantize.rs:47-237  (the value written is sign * x; the reference writes
                                  sign * POW43[x], a pure table lookup on x)
  BitReaderLtr / read_codebook    symphonia-core/src/io/bit.rs:557-808

The Huffman codes are matched bit by bit against the standard's (code, length) lists (tests/golden/mp3_huffman.json,
checked complete and prefix-free when generated) -- no lookup tables, nothing shared with the C++ front-end.
Pinned by: a frame assembled by hand from ISO/IEC 11172-3 field tables in tests/test_mp3_frontend.py, and round trips
through an independent bitstream writer (tests/_mp3_bitstream.py)."""
import json
import os

from oracle import packetizer_oracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(os.path.dirname(_HERE), "tests", "golden", "mp3_huffman.json")) as _f:
    _RAW = json.load(_f)

LINBITS = [0] * 16 + [1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13]  # codebooks.rs:553-556


def _codebook(name):
    t = _RAW[name]
    return {(l, c): ((i // t["wrap"]) << 4) | (i % t["wrap"]) for i, (c, l) in enumerate(zip(t["codes"], t["lens"]))}


_BIG = {}
for _t in (1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15):
    _BIG[_t] = _codebook(str(_t))
for _t in range(16, 24):
    _BIG[_t] = _codebook("16")
for _t in range(24, 32):
    _BIG[_t] = _codebook("24")
_QUAD = [_codebook("quadA"), _codebook("quadB")]

SFB_LONG = [  # layer3/common.rs:9-55
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


class DecodeError(Exception):
    pass


class BitsLtr:
    """BitReaderLtr: most-significant bit first; reads past the end fail; a codeword is matched against the data
    padded with zeros and then must fit in what is left (bit.rs:771-808)."""

    def __init__(self, data):
        self.data = bytes(data)
        self.n = len(self.data) * 8
        self.at = 0

    def _bit(self, i):
        return (self.data[i >> 3] >> (7 - (i & 7))) & 1 if i < self.n else 0

    def read(self, width):
        if self.at + width > self.n:
            raise DecodeError("end of bitstream")
        v = 0
        for _ in range(width):
            v = (v << 1) | self._bit(self.at)
            self.at += 1
        return v

    def ignore(self, width):
        if self.at + width > self.n:
            raise DecodeError("end of bitstream")
        self.at += width

    def read_codebook(self, book):
        code = 0
        for length in range(1, 20):
            code = (code << 1) | self._bit(self.at + length - 1)
            if (length, code) in book:
                if length > self.n - self.at:
                    raise DecodeError("end of bitstream")
                self.at += length
                return book[(length, code)], length
        raise AssertionError("complete prefix codes always match")


class Reservoir:
    """layer3/mod.rs:34-108."""

    def __init__(self):
        self.buf = bytearray(2048)
        self.len = self.consumed = 0

    def fill(self, main_data, begin):
        end = begin + len(main_data)
        if end > len(self.buf):
            raise DecodeError("main_data length")
        unread = self.len - self.consumed
        if begin <= unread:
            self.buf[0:begin] = self.buf[self.len - begin:self.len]
            self.buf[begin:end] = main_data
            self.len = end
            underflow = 0
        else:
            self.buf[0:unread] = self.buf[self.len - unread:self.len]
            self.buf[unread:unread + len(main_data)] = main_data
            self.len = unread + len(main_data)
            underflow = begin - unread
        self.consumed = 0
        return underflow

    def consume(self, n):
        self.consumed = min(self.len, self.consumed + n)

    def bytes_ref(self):
        return bytes(self.buf[self.consumed:self.len])

    def clear(self):
        self.len = self.consumed = 0


def _new_channel():
    return dict(part2_3_length=0, big_values=0, global_gain=0, scalefac_compress=0, block_type="long", mixed=False, subblock_gain=[0, 0, 0],
                table_select=[0, 0, 0], region1_start=0, region2_start=0, preflag=False, scalefac_scale=False, count1table_select=0,
                scalefacs=[0] * 39, rzero=0)


def read_side_info(bs, h):
    """bitstream.rs:57-236.  Returns (main_data_begin, scfsi, granules[gr][ch])."""
    mpeg1 = h["version"] == "1"
    n_ch = h["n_channels"]
    scfsi = [[False] * 4, [False] * 4]
    if mpeg1:
        begin = bs.read(9)
        bs.ignore(5 if n_ch == 1 else 3)
        for ch in range(n_ch):
            for b in range(4):
                scfsi[ch][b] = bs.read(1) == 1
    else:
        begin = bs.read(8)
        bs.ignore(1 if n_ch == 1 else 2)
    granules = [[_new_channel(), _new_channel()], [_new_channel(), _new_channel()]]
    bands = SFB_LONG[h["sample_rate_idx"]]
    for gr in range(2 if mpeg1 else 1):
        for ch in range(n_ch):
            c = granules[gr][ch]
            c["part2_3_length"] = bs.read(12)
            c["big_values"] = bs.read(9)
            if c["big_values"] > 288:
                raise DecodeError("big_values")
            c["global_gain"] = bs.read(8)
            c["scalefac_compress"] = bs.read(4 if mpeg1 else 9)
            if bs.read(1):
                enc = bs.read(2)
                mixed = bs.read(1) == 1
                if enc == 0:
                    raise DecodeError("block_type")
                c["block_type"] = {1: "start", 2: "short", 3: "end"}[enc]
                c["mixed"] = mixed and enc == 2
                for i in range(2):
                    c["table_select"][i] = bs.read(5)
                for i in range(3):
                    c["subblock_gain"][i] = bs.read(3)
                if h["version"] == "2.5":
                    c["region1_start"] = bands[6 if (enc == 2 and not mixed) else 8]
                elif mpeg1 or enc == 2:
                    c["region1_start"] = 36
                else:
                    c["region1_start"] = 54
                c["region2_start"] = 576
            else:
                for i in range(3):
                    c["table_select"][i] = bs.read(5)
                r0 = bs.read(4) + 1
                r01 = bs.read(3) + r0 + 1
                c["region1_start"] = bands[r0]
                c["region2_start"] = bands[r01] if r01 <= 22 else 576
            c["preflag"] = bs.read(1) == 1 if mpeg1 else False
            c["scalefac_scale"] = bs.read(1) == 1
            c["count1table_select"] = bs.read(1)
    return begin, scfsi, granules


_SLEN = [(0, 0), (0, 1), (0, 2), (0, 3), (3, 0), (1, 1), (1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 3), (4, 2), (4, 3)]
_NSFB = [[[7, 7, 7, 0], [12, 12, 12, 0], [6, 15, 12, 0]], [[6, 6, 6, 3], [12, 9, 9, 6], [6, 12, 9, 6]], [[8, 8, 5, 0], [15, 12, 9, 0], [6, 18, 9, 0]],
         [[6, 5, 5, 5], [9, 9, 9, 9], [6, 9, 9, 9]], [[6, 5, 7, 3], [9, 9, 12, 6], [6, 9, 12, 6]], [[11, 10, 0, 0], [18, 18, 0, 0], [15, 18, 0, 0]]]


def read_scale_factors_mpeg1(bs, gr, ch, scfsi, granules):
    """bitstream.rs:240-318."""
    c = granules[gr][ch]
    s1, s2 = _SLEN[c["scalefac_compress"]]
    bits = 0
    if c["block_type"] == "short":
        n = 17 if c["mixed"] else 18
        if s1:
            for i in range(n):
                c["scalefacs"][i] = bs.read(s1)
            bits += n * s1
        if s2:
            for i in range(n, n + 18):
                c["scalefacs"][i] = bs.read(s2)
            bits += 18 * s2
        return bits
    for g, (a, b) in enumerate(((0, 6), (6, 11), (11, 16), (16, 21))):
        s = s1 if g < 2 else s2
        if gr > 0 and scfsi[ch][g]:
            c["scalefacs"][a:b] = granules[0][ch]["scalefacs"][a:b]
        elif s:
            for i in range(a, b):
                c["scalefacs"][i] = bs.read(s)
            bits += s * (b - a)
    return bits


def read_scale_factors_mpeg2(bs, intensity_channel, c):
    """bitstream.rs:320-427."""
    block = {True: 2, False: 1}[c["mixed"]] if c["block_type"] == "short" else 0
    if intensity_channel:
        sfc = c["scalefac_compress"] >> 1
        if sfc < 180:
            slen, row = [sfc // 36, (sfc % 36) // 6, (sfc % 36) % 6, 0], 0
        elif sfc < 244:
            slen, row = [((sfc - 180) % 64) >> 4, ((sfc - 180) % 16) >> 2, (sfc - 180) % 4, 0], 1
        else:
            slen, row = [(sfc - 244) // 3, (sfc - 244) % 3, 0, 0], 2
    else:
        sfc = c["scalefac_compress"]
        c["preflag"] = sfc >= 500
        if sfc < 400:
            slen, row = [(sfc >> 4) // 5, (sfc >> 4) % 5, (sfc % 16) >> 2, sfc % 4], 3
        elif sfc < 500:
            slen, row = [((sfc - 400) >> 2) // 5, ((sfc - 400) >> 2) % 5, (sfc - 400) % 4, 0], 4
        else:
            slen, row = [(sfc - 500) // 3, (sfc - 500) % 3, 0, 0], 5
    bits = start = 0
    for s, n in zip(slen, _NSFB[row][block]):
        if s:
            for i in range(start, start + n):
                c["scalefacs"][i] = bs.read(s)
            bits += s * n
        start += n
    return bits


def read_huffman_samples(bs, c, part3_bits):
    """requantize.rs:47-237.  Returns (rzero, 576 signed integers)."""
    buf = [0] * 576
    if part3_bits == 0:
        return 0, buf
    bits_read = 0
    i = 0
    big_len = 2 * c["big_values"]
    regions = [min(c["region1_start"], big_len), min(c["region2_start"], big_len), min(576, big_len)]
    for r, region_end in enumerate(regions):
        sel = c["table_select"][r]
        book = _BIG.get(sel)
        linbits = LINBITS[sel]
        if book is None:  # tables 0, 4 and 14 hold no codes: the region is silent
            while i < region_end:
                buf[i] = buf[i + 1] = 0
                i += 2
            continue
        while i < region_end and bits_read < part3_bits:
            value, n = bs.read_codebook(book)
            bits_read += n
            for k, x in enumerate((value >> 4, value & 15)):
                if x > 0:
                    if x == 15 and linbits > 0:
                        x += bs.read(linbits)
                        bits_read += linbits
                    buf[i + k] = -x if bs.read(1) else x
                    bits_read += 1
                else:
                    buf[i + k] = 0
            i += 2
    book = _QUAD[c["count1table_select"]]
    while i <= 572 and bits_read < part3_bits:
        value, n = bs.read_codebook(book)
        bits_read += n
        ones = bin(value & 15).count("1")
        signs = bs.read(ones)
        bits_read += ones
        for k, mask in ((3, 1), (2, 2), (1, 4), (0, 8)):
            if value & mask:
                buf[i + k] = -1 if signs & 1 else 1
                signs >>= 1
            else:
                buf[i + k] = 0
        i += 4
    if bits_read < part3_bits:
        bs.ignore(part3_bits - bits_read)
    elif bits_read > part3_bits and i > big_len:
        i -= 4
    for k in range(i, 576):
        buf[k] = 0
    return i, buf


class Mp3Frontend:
    """One stream: MpaDecoder::decode_inner + Layer3::decode up to the synthesis seam."""

    def __init__(self):
        self.reservoir = Reservoir()
        self.spec = None

    def reset(self):
        self.__init__()

    def decode(self, packet):
        """Returns (header, granules[gr][ch] dicts incl. rzero, quant[gr][ch] lists, underflow bytes, consumed bytes) or raises
        DecodeError.  Absent channels / granules keep their defaults."""
        r = po.Reader(packet)
        try:
            word = po.mpa_sync_frame(r)
            h = po.mpa_parse_header(word)
        except po.ReaderError as e:
            raise DecodeError(str(e))
        if h["frame_size"] != r.available():
            raise DecodeError("packet length")
        spec = (h["sample_rate"], h["n_channels"])
        if self.spec is None:
            self.spec = spec
        elif self.spec != spec:
            raise DecodeError("signal spec")
        if h["layer"] != 3:
            raise DecodeError("layer")
        body = packet[r.pos:]
        if h["crc"]:
            if len(body) < 2:
                raise DecodeError("crc")
            body = body[2:]
        bs = BitsLtr(body)
        try:
            begin, scfsi, granules = read_side_info(bs, h)
        except DecodeError:
            self.reservoir.clear()
            raise
        underflow = self.reservoir.fill(body[h["side_info_len"]:], begin)  # an error here leaves the reservoir as it was
        try:
            used, quant = self._read_main_data(h, 8 * underflow, scfsi, granules)
        except DecodeError:
            self.reservoir.clear()
            raise
        self.reservoir.consume(used)
        return h, granules, quant, underflow, used

    def _read_main_data(self, h, underflow_bits, scfsi, granules):
        """layer3/mod.rs:272-370."""
        main_data = self.reservoir.bytes_ref()
        quant = [[[0] * 576, [0] * 576], [[0] * 576, [0] * 576]]
        begin = skipped = 0
        mpeg1 = h["version"] == "1"
        intensity = h["mode"] == "joint" and h["intensity"]
        for gr in range(2 if mpeg1 else 1):
            if skipped < underflow_bits:
                for ch in range(h["n_channels"]):
                    skipped += granules[gr][ch]["part2_3_length"]
                if skipped > underflow_bits:
                    begin = skipped - underflow_bits
                continue
            for ch in range(h["n_channels"]):
                byte_index = begin >> 3
                if byte_index > len(main_data):
                    raise DecodeError("main_data offset")
                bs = BitsLtr(main_data[byte_index:])
                if begin & 7:
                    bs.ignore(begin & 7)
                c = granules[gr][ch]
                if mpeg1:
                    part2 = read_scale_factors_mpeg1(bs, gr, ch, scfsi, granules)
                else:
                    part2 = read_scale_factors_mpeg2(bs, ch > 0 and intensity, c)
                if part2 > c["part2_3_length"]:
                    raise DecodeError("part2_3_length")
                c["rzero"], quant[gr][ch] = read_huffman_samples(bs, c, c["part2_3_length"] - part2)
                begin += c["part2_3_length"]
        return (begin + 7) >> 3, quant
