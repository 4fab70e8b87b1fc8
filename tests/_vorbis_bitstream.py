"""Vorbis stream WRITER for the front-end tests (Vorbis I specification sections 3, 4, 6-8): setup headers with fully specified
Huffman codebooks (all three length codings, VQ lookup types 1 and 2, sequence flag), floor-1 and residue (types 0 / 1 / 2)
configurations within what the synthesis kernel supports, and audio packets whose floor Y values and residue vectors are
chosen at random, coded, and kept as ground truth.  Builders only; nothing here reads a bitstream."""
import numpy as np

from tests._streams import BitWriterRtl, _ilog, vorbis_ident

f32 = np.float32


def canonical_codewords(lens):
    """Vorbis I 3.2.1: entries in order, each takes the lowest-valued unused leaf of its length (free sub-trees kept explicitly)."""
    free = [(0, 0)]  # (prefix, depth)
    out = []
    for n in lens:
        if n == 0:
            out.append(None)
            continue
        cands = [(p << (n - d), k) for k, (p, d) in enumerate(free) if d <= n]
        assert cands, "over-specified"
        value, k = min(cands)
        p, d = free.pop(k)
        for depth in range(d + 1, n + 1):
            free.append((((p << (depth - d)) | 1), depth))
        out.append(value)
    assert not free, "under-specified"
    return out


def random_lengths(rng, n_used, max_len=14):
    """Leaf depths of a random full binary tree with n_used leaves."""
    if n_used == 1:
        return [1]  # the single-entry special case (both one-bit codes decode to it)
    leaves = [1, 1]
    while len(leaves) < n_used:
        splittable = [k for k, d in enumerate(leaves) if d < max_len]
        k = int(rng.choice(splittable))
        d = leaves.pop(k)
        leaves += [d + 1, d + 1]
    rng.shuffle(leaves)
    return [int(x) for x in leaves]


def pack_float32(mantissa, exponent, negative=False):
    """Inverse of float32_unpack (Vorbis I 9.2.2): value = mantissa * 2^(exponent - 788)."""
    return (0x80000000 if negative else 0) | (exponent << 21) | mantissa


class Book:
    """One codebook: the bits of its setup entry and what a decoder must make of them."""

    def __init__(self, rng, entries, dims, vq=None, style=None, sparse_unused=0.3):
        self.entries, self.dims = entries, dims
        style = style or ["plain", "sparse", "ordered"][int(rng.integers(3))]
        w = BitWriterRtl()
        w.put(0x564342, 24), w.put(dims, 16), w.put(entries, 24)
        if style == "ordered":
            lens = sorted(random_lengths(rng, entries))
            w.put(1, 1), w.put(lens[0] - 1, 5)
            cur, n = 0, lens[0]
            while cur < entries:
                num = sum(1 for x in lens if x == n)
                w.put(num, _ilog(entries - cur))
                cur += num
                n += 1
            self.lens = lens
        elif style == "sparse" and entries > 2:
            used = [bool(rng.random() > sparse_unused) for _ in range(entries)]
            if sum(used) < 2:
                used[0] = used[1] = True
            depths = random_lengths(rng, sum(used))
            it = iter(depths)
            self.lens = [next(it) if u else 0 for u in used]
            w.put(0, 1), w.put(1, 1)
            for n in self.lens:
                if n:
                    w.put(1, 1), w.put(n - 1, 5)
                else:
                    w.put(0, 1)
        else:
            self.lens = random_lengths(rng, entries)
            w.put(0, 1), w.put(0, 1)
            for n in self.lens:
                w.put(n - 1, 5)
        if entries == 1:
            self.codes = [0]
        else:
            self.codes = canonical_codewords(self.lens)
        self.usable = [e for e, n in enumerate(self.lens) if n]
        self.vq = None
        if vq is None:
            w.put(0, 4)
        else:
            lookup, seq = vq
            w.put(lookup, 4)
            lo_m, lo_e, lo_neg = int(rng.integers(0, 1 << 12)), int(rng.integers(788 - 14, 788 - 6)), bool(rng.integers(2))
            d_m, d_e = int(rng.integers(1, 1 << 10)), int(rng.integers(788 - 14, 788 - 8))
            w.put(pack_float32(lo_m, lo_e, lo_neg), 32), w.put(pack_float32(d_m, d_e), 32)
            lo = f32(-(lo_m * 2.0 ** (lo_e - 788)) if lo_neg else lo_m * 2.0 ** (lo_e - 788))
            delta = f32(d_m * 2.0 ** (d_e - 788))
            value_bits = int(rng.integers(2, 9))
            w.put(value_bits - 1, 4), w.put(int(seq), 1)
            if lookup == 1:
                n_values = 0
                while (n_values + 1) ** dims <= entries:
                    n_values += 1
            else:
                n_values = entries * dims
            mult = [int(rng.integers(1 << value_bits)) for _ in range(n_values)]
            for m in mult:
                w.put(m, value_bits)
            table = np.zeros((entries, dims), dtype=np.float32)
            for e in range(entries):
                last = f32(0)
                for d in range(dims):
                    m = mult[(e // n_values ** d) % n_values] if lookup == 1 else mult[e * dims + d]
                    v = f32(f32(f32(f32(m) * delta) + lo) + last)
                    table[e, d] = v
                    if seq:
                        last = v
            self.vq = table
        self.bits = w

    def put(self, w, entry):
        """Writes the codeword of `entry`: first bit = root of the tree."""
        n, code = self.lens[entry], self.codes[entry]
        if self.entries == 1:
            w.put(int(np.random.default_rng(entry).integers(2)), 1)  # either one-bit code
            return
        for b in range(n - 1, -1, -1):
            w.put((code >> b) & 1, 1)


class Stream:
    """A whole logical stream's headers + a packet generator with ground truth."""

    def __init__(self, rng, channels=2, bs_exp=(7, 9), residue_type=None, coupled=None, per_word=None, residue_begin=None):
        self.rng, self.channels, self.bs_exp = rng, channels, bs_exp
        self.ident = vorbis_ident(channels=channels, bs0=bs_exp[0], bs1=bs_exp[1])
        books = []
        # floor books: scalar, small alphabets
        n_floor_books = int(rng.integers(2, 5))
        for _ in range(n_floor_books):
            books.append(Book(rng, int(rng.integers(1, 40)), 1))
        # residue class book: dims = partitions per class word, entries = classifications^dims exactly
        self.classifications = int(rng.integers(1, 5))
        # class words of several partitions: when the partition count is not a multiple of it, the reference lets the last word spill
        # into the next channel's classes (residue.rs:451-477 bounds the write by the vector's end) -- the truth kept here follows
        # the specification, so comparisons against it use per_word = 1; reader-vs-reader comparisons use any
        self.per_word = int(rng.integers(1, 4)) if per_word is None else per_word
        class_book = len(books)
        books.append(Book(rng, self.classifications ** self.per_word, self.per_word, style="plain"))
        vq_first = len(books)
        for _ in range(int(rng.integers(2, 5))):
            dims = int(rng.choice([1, 2, 4, 8]))
            books.append(Book(rng, int(rng.integers(2, 30)), dims, vq=(int(rng.integers(1, 3)), bool(rng.integers(2)))))
        self.books = books
        w = BitWriterRtl()
        w.put(len(books) - 1, 8)
        for b in books:
            w.v |= b.bits.v << w.n
            w.n += b.bits.n
        w.put(0, 6), w.put(0, 16)
        # floors (type 1)
        self.floors = []
        n_floors = int(rng.integers(1, 4))
        w.put(n_floors - 1, 6)
        for _ in range(n_floors):
            w.put(1, 16)
            rangebits = bs_exp[0] - 1
            parts = int(rng.integers(0, 6))
            pclass = [int(rng.integers(0, 3)) for _ in range(parts)]
            classes = {}
            w.put(parts, 5)
            for c in pclass:
                w.put(c, 4)
            if parts:
                for c in range(max(pclass) + 1):
                    dims, sub = int(rng.integers(1, 4)), int(rng.integers(0, 3))
                    w.put(dims - 1, 3), w.put(sub, 2)
                    main = int(rng.integers(n_floor_books))
                    if sub:
                        w.put(main, 8)
                    subbooks = []
                    for _k in range(1 << sub):
                        sb = int(rng.integers(0, n_floor_books + 1))  # 0 = none
                        w.put(sb, 8)
                        subbooks.append(sb - 1 if sb else None)
                    classes[c] = dict(dims=dims, sub=sub, main=main, subbooks=subbooks)
            mult = int(rng.integers(1, 5))
            w.put(mult - 1, 2), w.put(rangebits, 4)
            n_x = sum(classes[c]["dims"] for c in pclass)
            xs = [int(v) for v in rng.choice(np.arange(1, 1 << rangebits), size=n_x, replace=False)]
            for x in xs:
                w.put(x, rangebits)
            self.floors.append(dict(multiplier=mult, pclass=pclass, classes=classes, n_posts=2 + n_x))
        # residues
        self.residues = []
        n_res = int(rng.integers(1, 3))
        w.put(n_res - 1, 6)
        for _ in range(n_res):
            rtype = int(rng.integers(3)) if residue_type is None else residue_type
            n2_short = (1 << bs_exp[0]) >> 1
            part_size = int(rng.choice([8, 16]))
            begin = int(rng.choice([0, part_size])) if residue_begin is None else residue_begin
            end = int(rng.choice([n2_short, (1 << bs_exp[1]) >> 1, 1 << bs_exp[1], 3 * part_size + begin]))
            if end < begin:   # (a setup header with end < begin is refused, residue.rs:88-90)
                end = begin + 3 * part_size
            w.put(rtype, 16), w.put(begin, 24), w.put(end, 24), w.put(part_size - 1, 24)
            w.put(self.classifications - 1, 6), w.put(class_book, 8)
            used = []
            for _c in range(self.classifications):
                u = int(rng.integers(0, 8)) | (int(rng.integers(2)) << int(rng.integers(3, 8)))
                w.put(u & 7, 3)
                if u >> 3:
                    w.put(1, 1), w.put(u >> 3, 5)
                else:
                    w.put(0, 1)
                used.append(u)
            vbooks = [[None] * 8 for _ in used]
            for ci, u in enumerate(used):
                for j in range(8):
                    if u >> j & 1:
                        vbooks[ci][j] = int(rng.integers(vq_first, len(books)))
                        w.put(vbooks[ci][j], 8)
            self.residues.append(dict(type=rtype, begin=begin, end=end, part_size=part_size, used=used, books=vbooks))
        # one mapping per coupling choice would need matching flags across modes: use a single coupling choice for the stream
        self.coupled = bool(rng.integers(2)) if (coupled is None and channels == 2) else bool(coupled and channels == 2)
        n_map = int(rng.integers(1, 3))
        w.put(n_map - 1, 6)
        self.mappings = []
        for _ in range(n_map):
            w.put(0, 16)
            submaps = int(rng.integers(1, 3)) if channels == 2 else 1
            if submaps > 1:
                w.put(1, 1), w.put(submaps - 1, 4)
            else:
                w.put(0, 1)
            if self.coupled:
                w.put(1, 1), w.put(0, 8), w.put(0, 1), w.put(1, 1)  # one step: magnitude channel 0, angle channel 1 (1 bit each)
            else:
                w.put(0, 1)
            w.put(0, 2)
            mux = [0] * channels
            if submaps > 1:
                mux = [int(rng.integers(submaps)) for _ in range(channels)]
                for m in mux:
                    w.put(m, 4)
            sm = []
            for _k in range(submaps):
                fl, rs = int(rng.integers(n_floors)), int(rng.integers(n_res))
                w.put(0, 8), w.put(fl, 8), w.put(rs, 8)
                sm.append((fl, rs))
            self.mappings.append(dict(mux=mux, submaps=sm))
        n_modes = int(rng.integers(1, 5))
        w.put(n_modes - 1, 6)
        self.modes = []
        for _ in range(n_modes):
            flag, mp = int(rng.integers(2)), int(rng.integers(n_map))
            w.put(flag, 1), w.put(0, 16), w.put(0, 16), w.put(mp, 8)
            self.modes.append((bool(flag), mp))
        w.put(1, 1)
        self.setup = b"\x05vorbis" + w.bytes()
        self.prev_flag = None

    def packet(self, unused_prob=0.15):
        """One audio packet: (bytes, truth dict like the oracle's decode result)."""
        rng = self.rng
        w = BitWriterRtl()
        w.put(0, 1)
        mode = int(rng.integers(len(self.modes)))
        w.put(mode, _ilog(len(self.modes) - 1))
        long_block, mp = self.modes[mode]
        if long_block:
            w.put(int(rng.integers(2)), 1), w.put(int(rng.integers(2)), 1)
        mapping = self.mappings[mp]
        n2 = (1 << (self.bs_exp[1] if long_block else self.bs_exp[0])) >> 1
        slot = (1 << self.bs_exp[1]) >> 1
        floor_y = np.zeros((2, 65), dtype=np.uint16)
        dnd, floor_idx = [True, True], [None, None]
        for ch in range(self.channels):
            fi = mapping["submaps"][mapping["mux"][ch]][0]
            f = self.floors[fi]
            if rng.random() < unused_prob:
                w.put(0, 1)
                continue
            w.put(1, 1)
            rng_ = {1: 256, 2: 128, 3: 86, 4: 64}[f["multiplier"]]
            bits = _ilog(rng_ - 1)
            y = [int(rng.integers(rng_)), int(rng.integers(rng_))]
            w.put(y[0], bits), w.put(y[1], bits)
            for c in f["pclass"]:
                cl = f["classes"][c]
                cval = 0
                if cl["sub"]:
                    book = self.books[cl["main"]]
                    cval = int(rng.choice(book.usable))
                    book.put(w, cval)
                for _d in range(cl["dims"]):
                    sub = cval & ((1 << cl["sub"]) - 1)
                    cval >>= cl["sub"]
                    sb = cl["subbooks"][sub]
                    if sb is None:
                        y.append(0)
                    else:
                        e = int(rng.choice(self.books[sb].usable))
                        self.books[sb].put(w, e)
                        y.append(e)
            dnd[ch], floor_idx[ch] = False, fi
            floor_y[ch, :len(y)] = y
        if self.coupled and dnd[0] != dnd[1]:
            dnd = [False, False]
        residue = np.zeros((2, slot), dtype=np.float32)
        for sm, (_fl, ri) in enumerate(mapping["submaps"]):
            chans = [c for c in range(self.channels) if mapping["mux"][c] == sm]
            if not chans:
                continue
            r = self.residues[ri]
            count = len(chans)
            full = n2 * count if r["type"] == 2 else n2
            begin, end = min(r["begin"], full), min(r["end"], full)
            parts = (end - begin) // r["part_size"]
            if not any(not dnd[c] for c in chans):
                continue
            buf = np.zeros(full, dtype=np.float32)
            lanes = [None] if r["type"] == 2 else [c for c in chans]
            active = [True] if r["type"] == 2 else [not dnd[c] for c in chans]
            classes = {k: [0] * (parts + self.per_word) for k in range(len(lanes))}
            max_pass = max([j for u in r["used"] for j in range(8) if u >> j & 1], default=0)
            class_book = self.books[[k for k, b in enumerate(self.books) if b.dims == self.per_word and b.entries == self.classifications ** self.per_word and b.vq is None][-1]]
            for p in range(max_pass + 1):
                for first in range(0, parts, self.per_word):
                    if p == 0:
                        for k in range(len(lanes)):
                            if not active[k]:
                                continue
                            group = [int(rng.integers(self.classifications)) for _ in range(self.per_word)]
                            val = 0
                            for g in group:
                                val = val * self.classifications + g
                            class_book.put(w, val)
                            classes[k][first:first + self.per_word] = group
                    for part in range(first, min(parts, first + self.per_word)):
                        for k in range(len(lanes)):
                            if not active[k]:
                                continue
                            cls = classes[k][part]
                            if not r["used"][cls] >> p & 1:
                                continue
                            book = self.books[r["books"][cls][p]]
                            start = begin + r["part_size"] * part
                            target = buf if r["type"] == 2 else residue[lanes[k]]
                            n = r["part_size"]
                            if r["type"] == 0:
                                step = n // book.dims
                                for i in range(step):
                                    e = int(rng.integers(book.entries)) if len(book.usable) == book.entries else int(rng.choice(book.usable))
                                    book.put(w, e)
                                    for d, o in zip(range(book.dims), range(i, n, step)):
                                        target[start + o] = f32(target[start + o] + book.vq[e, d])
                            else:
                                for o in range(0, n - book.dims + 1, book.dims):
                                    e = int(rng.choice(book.usable))
                                    book.put(w, e)
                                    for d in range(book.dims):
                                        target[start + o + d] = f32(target[start + o + d] + book.vq[e, d])
            if r["type"] == 2:
                for i, c in enumerate(chans):
                    residue[c, :n2] = buf[i::count][:n2]
        prev = long_block if self.prev_flag is None else self.prev_flag
        self.prev_flag = long_block
        return w.bytes(), dict(block_flag=long_block, prev_block_flag=prev, do_not_decode=dnd, floor=floor_idx, floor_y=floor_y, residue=residue)
