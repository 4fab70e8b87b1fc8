"""Vorbis entropy front-end oracle: codebooks, floor-1 packet decode, residue decode and the packet-level steps
of VorbisDecoder::decode_inner up to inverse coupling, in the reference's sequence.  TEST INFRASTRUCTURE ONLY.

  symphonia-codec-vorbis/src/codebook.rs:16-400   float32_unpack, lookup1_values, VQ unpack, synthesize_codewords, read
  floor.rs:655-722 (Floor1::read_channel), residue.rs:142-543, lib.rs:146-250
  symphonia-core/src/io/bit.rs:941-1027, :1211-1250, :1305-1370   BitReaderRtl -- followed state for state (cache refills),
      because a Vorbis packet may legally end early and decoding continues with whatever the failed read left behind

numpy float32 scalars carry the arithmetic (one IEEE operation per reference operation).  Pinned by the reference's own unit
tests: verify_synthesize_codewords (+ the over-specified cases), verify_lookup1_values, verify_ilog (codebook.rs:402-485)."""
import numpy as np

from oracle import packetizer_oracle as po

f32 = np.float32


class End(Exception):
    """end_of_bitstream_error (io::ErrorKind::Other)."""


class PacketBits:
    """BitReaderRtl with its 64-bit cache."""

    def __init__(self, data):
        self.buf = bytes(data)
        self.pos = 0
        self.bits = 0
        self.left = 0

    def fetch(self):
        k = min(len(self.buf) - self.pos, 8)
        if k == 0:
            raise End()
        self.bits = int.from_bytes(self.buf[self.pos:self.pos + k], "little")
        self.pos += k
        self.left = 8 * k

    def top_up(self):
        k = min((64 - self.left) >> 3, len(self.buf) - self.pos)
        for i in range(k):
            self.bits |= self.buf[self.pos + i] << self.left
            self.left += 8
        self.pos += k

    def consume(self, w):
        self.left -= w
        self.bits >>= w

    def read(self, width):
        acc, needed = self.bits, width
        while needed > self.left:
            needed -= self.left
            self.fetch()
            acc |= (self.bits << (width - needed)) & 0xFFFFFFFFFFFFFFFF
        self.consume(needed)
        return acc & ((1 << width) - 1)

    def read_bool(self):
        if self.left < 1:
            self.fetch()
        b = self.bits & 1
        self.consume(1)
        return b == 1


def powi2(b):
    """compiler-rt __powisf2(2.0, b): what f32::powi lowers to."""
    recip = b < 0
    a, r = f32(2.0), f32(1.0)
    b = abs(b)  # (b /= 2 truncates toward zero: the parity of |b| sequence is the same)
    with np.errstate(over="ignore", divide="ignore"):
        while True:
            if b & 1:
                r = f32(r * a)
            b //= 2
            if b == 0:
                break
            a = f32(a * a)
        return f32(f32(1.0) / r) if recip else r


def float32_unpack(x):
    """codebook.rs:16-27."""
    with np.errstate(over="ignore", invalid="ignore"):
        value = f32(f32(x & 0x1FFFFF) * powi2(((x & 0x7FE00000) >> 21) - 788))
    return f32(-value) if x & 0x80000000 else value


def lookup1_values(entries, dims):
    """codebook.rs:33-49: floor(entries^(1/dims)) in f32, asserted to satisfy v^dims <= entries."""
    if dims == 0:
        return 0
    v = 0
    while (v + 1) ** dims <= entries:
        v += 1
    return v


def synthesize_codewords(lens):
    """codebook.rs:112-210, the reference's next-codeword table."""
    words = []
    nxt = [0] * 33
    for n in lens:
        if n == 0:
            continue
        code = nxt[n]
        if n < 32 and (code >> n) > 0:
            raise po.ReaderError(po.DECODE, "overspecified")
        for i in range(n, 0, -1):
            if nxt[i] & 1 == 1:
                if i == 1:
                    nxt[1] += 1
                else:
                    nxt[i] = (nxt[i - 1] << 1) & 0xFFFFFFFF
                break
            nxt[i] += 1
        branch = nxt[n]
        for i in range(1, 33 - n):
            if nxt[n + i] == (code << i) & 0xFFFFFFFF:
                nxt[n + i] = (branch << i) & 0xFFFFFFFF
            else:
                break
        words.append(code)
    if any(nxt[i] & (0xFFFFFFFF >> (32 - i)) for i in range(1, 33)):
        raise po.ReaderError(po.DECODE, "underspecified")
    return words


class Codebook:
    def __init__(self, bs):
        """codebook.rs:214-372 from the setup header's bit reader (errors are errors there)."""
        if bs.read(24) != 0x564342:
            raise po.ReaderError(po.DECODE, "sync")
        self.dims, entries = bs.read(16), bs.read(24)
        if self.dims == 0 or self.dims > 32 or entries > 128 * 1024:
            raise po.ReaderError(po.DECODE, "limits")
        lens, values = [], []
        if not bs.read_bool():
            if bs.read_bool():
                for e in range(entries):
                    if bs.read_bool():
                        lens.append(bs.read(5) + 1)
                        values.append(e)
            else:
                for e in range(entries):
                    lens.append(bs.read(5) + 1)
                values = list(range(entries))
        else:
            cur, n = 0, bs.read(5) + 1
            while True:
                num = bs.read(po.ilog(entries - cur) if entries > cur else 0)
                lens += [n] * num
                n += 1
                cur += num
                if cur > entries:
                    raise po.ReaderError(po.DECODE, "codebook")
                if cur == entries:
                    break
            values = list(range(cur))
            if any(x > 32 for x in lens):
                raise po.ReaderError(po.DECODE, "length beyond 32")  # the reference's 33-entry table would be indexed out of range
        if len(lens) == 1 and lens[0] == 1:
            lens.append(1)
            values.append(values[0])
        lookup = bs.read(4)
        self.vq = None
        if lookup in (1, 2):
            lo, delta = float32_unpack(bs.read(32)), float32_unpack(bs.read(32))
            value_bits = bs.read(4) + 1
            seq = bs.read_bool()
            n_values = lookup1_values(entries, self.dims) if lookup == 1 else entries * self.dims
            mult = [bs.read(value_bits) for _ in range(n_values)]
            vq = np.zeros((entries, self.dims), dtype=np.float32)
            with np.errstate(over="ignore", invalid="ignore"):
                for e in range(entries):
                    last, div = f32(0.0), 1
                    for d in range(self.dims):
                        at = (e // div) % n_values if lookup == 1 else e * self.dims + d
                        v = f32(f32(f32(f32(mult[at]) * delta) + lo) + last)
                        vq[e, d] = v
                        if seq:
                            last = v
                        div = (div * n_values) & 0xFFFFFFFF
            self.vq = vq
        elif lookup != 0:
            raise po.ReaderError(po.DECODE, "lookup type")
        words = synthesize_codewords(lens)
        used = [n for n in lens if n]
        self.max_len = max(used) if used else 0
        self.table = {}
        k = 0
        for n, v in zip(lens, values):
            if n == 0:
                continue
            self.table[(n, words[k])] = v
            k += 1

    def read(self, pb):
        """bit.rs:1211-1250 with a complete tree: first stream bit = most significant bit of the codeword."""
        if pb.left < self.max_len:
            pb.top_up()
        code = 0
        for depth in range(1, 33):
            code = (code << 1) | ((pb.bits >> (depth - 1)) & 1)
            if (depth, code) in self.table:
                if depth > pb.left:
                    raise End()
                pb.consume(depth)
                return self.table[(depth, code)]
        raise End()


class VorbisFrontend:
    def __init__(self, ident_packet, setup_packet):
        self.ident = po.vorbis_read_ident(ident_packet)
        self.setup = po.vorbis_read_setup(setup_packet, self.ident)
        bs = po.BitsRtl(setup_packet[7:])
        self.books = [Codebook(bs) for _ in range(bs.read(8) + 1)]
        self.prev_block_flag = None
        self.part_classes = []   # residue.rs:434-441: grows, never shrinks, never cleared
        self.type2 = []

    def _floor(self, f, pb):
        """floor.rs:655-722 -> list of Y values, or None when unused."""
        y = [0] * len(f["x_list"])
        try:
            if not pb.read_bool():
                return None
            bits = po.ilog({1: 256, 2: 128, 3: 86, 4: 64}[f["multiplier"]] - 1)
            y[0], y[1] = pb.read(bits), pb.read(bits)
            offset = 2
            for cidx in f["partition_class"]:
                cl = f["classes"][cidx]
                cbits = cl["subclass_bits"]
                cval = self.books[cl["mainbook"]].read(pb) if cbits else 0
                for d in range(cl["dimensions"]):
                    sub = cval & ((1 << cbits) - 1)
                    cval >>= cbits
                    y[offset + d] = self.books[cl["subbooks"][sub]].read(pb) if cl["used"] >> sub & 1 else 0
                offset += cl["dimensions"]
        except End:
            return None
        return y

    @staticmethod
    def _decode_classes(val, per_word, classifications, out, base):
        """residue.rs:451-477 on out[base:]."""
        num = len(out) - base
        skip = 0
        if per_word > num:
            skip = per_word - num
            for _ in range(skip):
                val //= classifications
        for k in range(per_word - skip - 1, -1, -1):
            out[base + k] = val % classifications
            val //= classifications

    def _partition(self, book, pb, out, start, n, format0):
        if book.vq is None:
            raise po.ReaderError(po.DECODE, "not a vq codebook")
        dim = book.dims
        if format0:
            step = n // dim
            for i in range(step):
                v = book.vq[book.read(pb)]
                for k, o in zip(range(dim), range(i, n, step)):
                    out[start + o] = f32(out[start + o] + v[k])
        else:
            for o in range(0, n - dim + 1, dim):
                v = book.vq[book.read(pb)]
                for k in range(dim):
                    out[start + o + k] = f32(out[start + o + k] + v[k])

    def _residue(self, r, pb, bs_exp, chans, dnd, residue):
        n2 = (1 << bs_exp) >> 1
        count = len(chans)
        full = n2 * count if r["type"] == 2 else n2
        begin, end = min(r["begin"], full), min(r["end"], full)
        per_word = self.books[r["classbook"]].dims
        parts = (end - begin) // r["partition_size"]
        need = parts if r["type"] == 2 else parts * count
        if len(self.part_classes) < need:
            self.part_classes += [0] * (need - len(self.part_classes))
        if r["type"] == 2:
            buf = np.zeros(full, dtype=np.float32)
        any_ch = any(not dnd[c] for c in chans)
        if any_ch:
            try:
                with np.errstate(over="ignore", invalid="ignore"):
                    for p in range(r["max_pass"] + 1):
                        for first in range(0, parts, per_word):
                            if p == 0:
                                for i, c in enumerate(chans if r["type"] != 2 else chans[:1]):
                                    if r["type"] != 2 and dnd[c]:
                                        continue
                                    code = self.books[r["classbook"]].read(pb)
                                    self._decode_classes(code, per_word, r["classifications"], self.part_classes, first + i * parts)
                            for part in range(first, min(parts, first + per_word)):
                                for i, c in enumerate(chans if r["type"] != 2 else chans[:1]):
                                    if r["type"] != 2 and dnd[c]:
                                        continue
                                    cls = self.part_classes[part + parts * i]
                                    if r["used"][cls] >> p & 1:
                                        book = self.books[r["books"][cls][p]]
                                        start = begin + r["partition_size"] * part
                                        target = buf if r["type"] == 2 else residue[c]
                                        self._partition(book, pb, target, start, r["partition_size"], r["type"] == 0)
            except End:
                pass
        if r["type"] == 2:
            for i, c in enumerate(chans):
                residue[c][:n2] = buf[i::count][:n2]

    def decode(self, packet, slot):
        """lib.rs:146-250.  Returns dict(block_flag, prev_block_flag, do_not_decode[2], floor[2] (index or None), floor_y [2][65],
        residue [2][slot] f32) or raises ReaderError for what the reference returns as an error."""
        pb = PacketBits(packet)
        try:
            if pb.read_bool():
                raise po.ReaderError(po.DECODE, "not an audio packet")
            modes = self.setup["modes"]
            mode_number = pb.read(po.ilog(len(modes) - 1))
            if mode_number >= len(modes):
                raise po.ReaderError(po.DECODE, "mode number")
            long_block, mapping_idx = modes[mode_number]
            if long_block:
                pb.read_bool(), pb.read_bool()
        except End:
            raise po.ReaderError(po.DECODE, "packet header cut")
        mapping = self.setup["mappings"][mapping_idx]
        bs_exp = self.ident["bs1_exp"] if long_block else self.ident["bs0_exp"]
        n_ch = self.ident["n_channels"]
        floor_y = np.zeros((2, 65), dtype=np.uint16)
        residue = [np.zeros(slot, dtype=np.float32) for _ in range(2)]
        dnd, floor_idx = [True, True], [None, None]
        for ch in range(n_ch):
            fi = mapping["submaps"][mapping["multiplex"][ch]][0]
            y = self._floor(self.setup["floors"][fi], pb)
            dnd[ch] = y is None
            if y is not None:
                floor_idx[ch] = fi
                floor_y[ch, :len(y)] = y
        for mag, ang in mapping["couplings"]:
            if dnd[mag] != dnd[ang]:
                dnd[mag] = dnd[ang] = False
        for sm, (_, res_idx) in enumerate(mapping["submaps"]):
            chans = [c for c in range(n_ch) if mapping["multiplex"][c] == sm]
            if not chans:
                continue
            self._residue(self.setup["residues"][res_idx], pb, bs_exp, chans, dnd, residue)
        prev = long_block if self.prev_block_flag is None else self.prev_block_flag
        self.prev_block_flag = long_block
        return dict(block_flag=long_block, prev_block_flag=prev, do_not_decode=dnd, floor=floor_idx, floor_y=floor_y, residue=np.stack(residue))
