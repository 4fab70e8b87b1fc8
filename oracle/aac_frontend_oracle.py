"""AAC-LC entropy front-end oracle (SURVEY §8f N1 for the AAC path): raw_data_block elements, ICS info, section data, scale
factors, pulse data, TNS data, spectrum (Huffman books 1-11, escapes, perceptual noise substitution), joint stereo (mid/side,
intensity), pulse restoration and the resolution of TNS filters to line ranges -- everything the reference does between
`BitReaderLtr::new(packet.data)` and `Dsp::synth`, in the reference's sequence.  TEST INFRASTRUCTURE ONLY.

  symphonia-codec-aac/src/aac/mod.rs:52-255           AacDecoder::try_new (no extra data), set_pair, decode_ga, decode_inner
  aac/cpe.rs:35-161                                    ChannelPair: decode_ga_sce / decode_ga_cpe (common window, ms mask, IS, M/S)
  aac/ics/mod.rs:103-632                               IcsInfo::decode, section data, scale factors, spectrum, noise, escapes
  aac/ics/pulse.rs:19-106, aac/ics/tns.rs:23-199       pulse read + synth, TNS read (+ the line ranges of Tns::synth)
  aac/common.rs:22-172                                 band tables, Lcg, GASubbandInfo
  symphonia-core/src/io/bit.rs:771-808                 read_codebook on a stream that ends

numpy float32 scalars carry the arithmetic, one IEEE operation per reference operation; powf / sinf are the C library's, which
is what the reference's f32::powf / f32::sin call on this platform (tests/test_aac_frontend.py shows the scale-factor tables to be the
correctly rounded powers of two whichever routine computes them, and x^(4/3) to be within one unit in the last place of the correctly
rounded value -- 10 of 8192 entries differ under glibc, so that table is a property of the platform's libm in the reference itself).
AudioSpecificConfig::read (symphonia-common/src/mpeg/audio/mod.rs:230-439) is restated at the end.  Pinned by the reference's
own unit test (decode_section_data_rejects_excess_zero_length_sections, ics/mod.rs:612-635) and by an independent stream
writer's ground truth."""
import ctypes
import ctypes.util
import json
import os

import numpy as np

f32 = np.float32
_m = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_m.powf.restype = _m.sinf.restype = ctypes.c_float
_m.powf.argtypes = [ctypes.c_float, ctypes.c_float]
_m.sinf.argtypes = [ctypes.c_float]


def powf(a, b):
    return f32(_m.powf(float(f32(a)), float(f32(b))))


def sinf(a):
    return f32(_m.sinf(float(f32(a))))


with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "aac_huffman.json")) as _f:
    _RAW = json.load(_f)
BOOKS = {k: {(l, c): i for i, (c, l) in enumerate(zip(t["codes"], t["lens"]))} for k, t in _RAW.items()}

DECODE, UNSUPPORTED = "decode", "unsupported"


class AacError(Exception):
    def __init__(self, kind, what=""):
        super().__init__(f"{kind}: {what}")
        self.kind = kind


class BitsLtr:
    """BitReaderLtr + FiniteBitStream."""

    def __init__(self, data):
        self.data, self.n, self.at = bytes(data), 8 * len(data), 0

    def _bit(self, i):
        return (self.data[i >> 3] >> (7 - (i & 7))) & 1 if i < self.n else 0

    def bits_left(self):
        return self.n - self.at

    def read(self, width):
        if self.at + width > self.n:
            raise AacError(DECODE, "end of bitstream")
        v = 0
        for _ in range(width):
            v = (v << 1) | self._bit(self.at)
            self.at += 1
        return v

    def read_bool(self):
        return self.read(1) == 1

    def ignore(self, width):
        if self.at + width > self.n:
            raise AacError(DECODE, "end of bitstream")
        self.at += width

    def realign(self):
        self.at = (self.at + 7) & ~7

    def read_unary_ones(self):
        n = 0
        while True:
            if self.at >= self.n:
                raise AacError(DECODE, "end of bitstream")
            b = self._bit(self.at)
            self.at += 1
            if not b:
                return n
            n += 1

    def read_codebook(self, book):
        code = 0
        for length in range(1, 20):
            code = (code << 1) | self._bit(self.at + length - 1)
            if (length, code) in book:
                if length > self.n - self.at:
                    raise AacError(DECODE, "end of bitstream")
                self.at += length
                return book[(length, code)]
        raise AssertionError("complete prefix codes always match")


# ---- tables (aac/common.rs:22-92, :121-172; aac/ics/tns.rs:22-24) ------------------------------------------------------------
L48 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 48, 56, 64, 72, 80, 88, 96, 108, 120, 132, 144, 160, 176, 196, 216, 240, 264, 292, 320,
       352, 384, 416, 448, 480, 512, 544, 576, 608, 640, 672, 704, 736, 768, 800, 832, 864, 896, 928, 1024]
S48 = [0, 4, 8, 12, 16, 20, 28, 36, 44, 56, 68, 80, 96, 112, 128]
L32 = L48[:49] + [960, 992, 1024]
L8 = [0, 12, 24, 36, 48, 60, 72, 84, 96, 108, 120, 132, 144, 156, 172, 188, 204, 220, 236, 252, 268, 288, 308, 328, 348, 372, 396, 420,
      448, 476, 508, 544, 580, 620, 664, 712, 764, 820, 880, 944, 1024]
S8 = [0, 4, 8, 12, 16, 20, 24, 28, 36, 44, 52, 60, 72, 88, 108, 128]
L16 = [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 100, 112, 124, 136, 148, 160, 172, 184, 196, 212, 228, 244, 260, 280, 300, 320, 344,
       368, 396, 424, 456, 492, 532, 572, 616, 664, 716, 772, 832, 896, 960, 1024]
S16 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 60, 72, 88, 108, 128]
L24 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 52, 60, 68, 76, 84, 92, 100, 108, 116, 124, 136, 148, 160, 172, 188, 204, 220, 240,
       260, 284, 308, 336, 364, 396, 432, 468, 508, 552, 600, 652, 704, 768, 832, 896, 960, 1024]
S24 = [0, 4, 8, 12, 16, 20, 24, 28, 36, 44, 52, 64, 76, 92, 108, 128]
L64 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 64, 72, 80, 88, 100, 112, 124, 140, 156, 172, 192, 216, 240, 268, 304,
       344, 384, 424, 464, 504, 544, 584, 624, 664, 704, 744, 784, 824, 864, 904, 944, 984, 1024]
S64 = [0, 4, 8, 12, 16, 20, 24, 32, 40, 48, 64, 92, 128]
L96 = [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 64, 72, 80, 88, 96, 108, 120, 132, 144, 156, 172, 188, 212, 240, 276,
       320, 384, 448, 512, 576, 640, 704, 768, 832, 896, 960, 1024]
SUBBAND_INFO = [(92017, L96, S64), (75132, L96, S64), (55426, L64, S64), (46009, L48, S48), (37566, L48, S48), (27713, L32, S48),
                (23004, L24, S24), (18783, L24, S24), (13856, L16, S16), (11502, L16, S16), (9391, L16, S16), (0, L8, S8)]
TNS_MAX_LONG = [31, 31, 34, 40, 42, 51, 46, 46, 42, 42, 42, 39]
TNS_MAX_SHORT = [9, 9, 10, 14, 14, 14, 14, 14, 14, 14, 14, 14]


def rate_index(rate):
    return next(i for i, (lo, _, _) in enumerate(SUBBAND_INFO) if rate >= lo)


POW43 = [powf(i, f32(4.0) / f32(3.0)) for i in range(8192)]                                  # ics/mod.rs:44-50
NORMAL_SCF = [powf(2.0, f32(0.25) * f32(i - 56 - 100)) for i in range(256)]                   # :58-66
INTENSITY_SCF = [powf(0.5, f32(0.25) * f32(i - 155)) for i in range(256)]                     # :74-82
QUAD_U2 = f32(2.51984209978974632953)


class Ics:
    def __init__(self, long_bands, short_bands):
        self.long_bands, self.short_bands = long_bands, short_bands
        self.reset()
        self.coeffs = np.zeros(1024, dtype=np.float32)
        self.pulse = self.tns = None
        self.sfb_cb = [[0] * 64 for _ in range(8)]
        self.scales = [[f32(0)] * 64 for _ in range(8)]
        self.max_sfb = 0

    def reset(self):
        """IcsInfo::new (ics/mod.rs:103-117)."""
        self.window_sequence = self.prev_window_sequence = 0
        self.window_shape = self.prev_window_shape = False
        self.grouping = [False] * 8
        self.group_start = [0] * 8
        self.window_groups = self.num_windows = self.max_sfb = 0
        self.long_win = True

    def bands(self):
        return self.long_bands if self.long_win else self.short_bands

    def info_fields(self):
        return (self.window_sequence, self.window_shape, list(self.grouping), list(self.group_start), self.window_groups,
                self.num_windows, self.max_sfb, self.long_win)

    def copy_from_common(self, other):
        seq, shape = self.window_sequence, self.window_shape
        (self.window_sequence, self.window_shape, self.grouping, self.group_start, self.window_groups, self.num_windows, self.max_sfb,
         self.long_win) = other.info_fields()
        self.prev_window_sequence, self.prev_window_shape = seq, shape

    def decode_info(self, bs):
        """IcsInfo::decode + Ics::decode_info (ics/mod.rs:120-177, :292-300)."""
        self.prev_window_sequence, self.prev_window_shape = self.window_sequence, self.window_shape
        if bs.read_bool():
            raise AacError(DECODE, "ics reserved bit")
        self.window_sequence = bs.read(2)
        self.window_shape = bs.read_bool()
        self.window_groups = 1
        if self.window_sequence == 2:
            self.long_win, self.num_windows = False, 8
            self.max_sfb = bs.read(4)
            for i in range(7):
                self.grouping[i] = bs.read_bool()
                if not self.grouping[i]:
                    self.group_start[self.window_groups] = i + 1
                    self.window_groups += 1
        else:
            self.long_win, self.num_windows = True, 1
            self.max_sfb = bs.read(6)
            if bs.read_bool():
                raise AacError(UNSUPPORTED, "predictor data")
        if self.max_sfb + 1 > len(self.bands()):
            raise AacError(DECODE, "max_sfb")

    def get_group_start(self, g):
        if g == 0:
            return 0
        if g >= self.window_groups:
            return 1 if self.long_win else 8
        return self.group_start[g]

    def decode_section_data(self, bs):
        bits = 5 if self.long_win else 3
        esc = (1 << bits) - 1
        for g in range(self.window_groups):
            k = l = 0
            while k < self.max_sfb:
                if l >= 64:
                    raise AacError(DECODE, "sections")
                cb = bs.read(4)
                if cb == 12:
                    raise AacError(DECODE, "band type")
                length = 0
                while True:
                    inc = bs.read(bits)
                    length += inc
                    if inc < esc:
                        break
                if k + length > self.max_sfb:
                    raise AacError(DECODE, "section length")
                for sfb in range(k, k + length):
                    self.sfb_cb[g][sfb] = cb
                k += length
                l += 1

    def decode_scale_factor_data(self, bs):
        noise_pcm, scf_int, scf_noise, scf_normal = True, 155, self.global_gain - 90 + 100, self.global_gain
        book = BOOKS["scf"]
        for g in range(self.window_groups):
            for sfb in range(self.max_sfb):
                cb = self.sfb_cb[g][sfb]
                if cb == 0:
                    v = f32(0)
                elif cb in (14, 15):
                    scf_int += bs.read_codebook(book) - 60
                    if not 0 <= scf_int < 256:
                        raise AacError(DECODE, "intensity scale")
                    v = INTENSITY_SCF[scf_int]
                elif cb == 13:
                    if noise_pcm:
                        noise_pcm = False
                        scf_noise += bs.read(9) - 256
                    else:
                        scf_noise += bs.read_codebook(book) - 60
                    if not 0 <= scf_noise < 256:
                        raise AacError(DECODE, "noise scale")
                    v = NORMAL_SCF[scf_noise]
                else:
                    scf_normal += bs.read_codebook(book) - 60
                    if not 0 <= scf_normal < 256:
                        raise AacError(DECODE, "scale")
                    v = NORMAL_SCF[scf_normal]
                self.scales[g][sfb] = v

    def decode_spectrum(self, bs, lcg):
        self.coeffs[:] = 0
        bands = self.bands()
        c = self.coeffs
        for g in range(self.window_groups):
            cur_w, next_w = self.get_group_start(g), self.get_group_start(g + 1)
            for sfb in range(self.max_sfb):
                cb, scale = self.sfb_cb[g][sfb], self.scales[g][sfb]
                for w in range(cur_w, next_w):
                    lo, hi = bands[sfb] + 128 * w, bands[sfb + 1] + 128 * w
                    if cb in (0, 12, 14, 15):
                        continue
                    if cb == 13:
                        energy = f32(0)
                        for i in range(lo, hi):
                            v = (lcg.next() >> 16) & 0xFFFF
                            c[i] = f32(v - 0x10000 if v & 0x8000 else v)
                            energy = f32(energy + f32(c[i] * c[i]))
                        with np.errstate(divide="ignore", invalid="ignore"):
                            s = f32(scale / np.sqrt(energy))
                            for i in range(lo, hi):
                                c[i] = f32(c[i] * s)
                    elif cb in (1, 2):
                        iq = [f32(-scale), f32(0), scale]
                        for i in range(lo, hi - 3, 4):
                            cw = bs.read_codebook(BOOKS[str(cb)])
                            for k, d in enumerate((cw // 27, cw // 9 % 3, cw // 3 % 3, cw % 3)):
                                c[i + k] = iq[d]
                    elif cb in (3, 4):
                        iq = [f32(0), scale, f32(QUAD_U2 * scale)]
                        for i in range(lo, hi - 3, 4):
                            cw = bs.read_codebook(BOOKS[str(cb)])
                            for k, d in enumerate((cw // 27, cw // 9 % 3, cw // 3 % 3, cw % 3)):
                                if d:
                                    c[i + k] = f32(_sign(bs.read(1)) * iq[d])
                    elif cb in (5, 6):
                        for i in range(lo, hi - 1, 2):
                            cw = bs.read_codebook(BOOKS[str(cb)])
                            for k, d in enumerate((cw // 9, cw % 9)):
                                x = f32(-POW43[4 - d]) if d < 4 else POW43[d - 4]
                                c[i + k] = f32(x * scale)
                    elif cb in (7, 8, 9, 10):
                        mod = 8 if cb < 9 else 13
                        for i in range(lo, hi - 1, 2):
                            cw = bs.read_codebook(BOOKS[str(cb)])
                            x, y = POW43[cw // mod], POW43[cw % mod]
                            sx = _sign(bs.read(1)) if x != 0 else f32(1)
                            sy = _sign(bs.read(1)) if y != 0 else f32(1)
                            c[i] = f32(f32(sx * x) * scale)
                            c[i + 1] = f32(f32(sy * y) * scale)
                    else:
                        for i in range(lo, hi - 1, 2):
                            cw = bs.read_codebook(BOOKS["11"])
                            a, b = cw // 17, cw % 17
                            sx = _sign(bs.read(1)) if a else f32(1)
                            sy = _sign(bs.read(1)) if b else f32(1)
                            x = POW43[_escape(bs) if a == 16 else a]
                            y = POW43[_escape(bs) if b == 16 else b]
                            c[i] = f32(f32(sx * x) * scale)
                            c[i + 1] = f32(f32(sy * y) * scale)

    def decode(self, bs, lcg, common_window):
        """Ics::decode (ics/mod.rs:403-447)."""
        self.global_gain = bs.read(8)
        if not common_window:
            self.decode_info(bs)
        self.decode_section_data(bs)
        self.decode_scale_factor_data(bs)
        self.pulse = None
        if bs.read_bool():
            n, start = bs.read(2) + 1, bs.read(6)
            self.pulse = (start, [(bs.read(5), bs.read(4)) for _ in range(n)])
        if self.pulse is not None and not self.long_win:
            raise AacError(DECODE, "pulse in a short window")
        self.tns = self._read_tns(bs)
        if bs.read_bool():
            raise AacError(DECODE, "gain control")
        self.decode_spectrum(bs, lcg)

    def _read_tns(self, bs):
        """Tns::read + TnsCoeffs::read (tns.rs:35-147) -> per window a list of (length, order, direction, lpc)."""
        if not bs.read_bool():
            return None
        max_order = 12 if self.long_win else 7
        out = []
        for _w in range(self.num_windows):
            n_filt = bs.read(2 if self.long_win else 1)
            coef_res = bs.read_bool() if n_filt else False
            filters = []
            for _ in range(n_filt):
                length, order = bs.read(6 if self.long_win else 4), bs.read(5 if self.long_win else 3)
                if order > max_order:
                    raise AacError(DECODE, "tns order")
                direction, lpc = False, [f32(0)] * 21
                if order:
                    direction = bs.read_bool()
                    res_bits = (4 if coef_res else 3) - (1 if bs.read_bool() else 0)
                    sign_mask, full = 1 << (res_bits - 1), 1 << res_bits
                    fac = f32(8.0 if coef_res else 4.0)
                    half_pi = f32(np.pi / 2)
                    iqfac, iqfac_m = f32(f32(fac - f32(0.5)) / half_pi), f32(f32(fac + f32(0.5)) / half_pi)
                    tmp = []
                    for _k in range(order):
                        val = bs.read(res_bits)
                        cc = f32(val - full if val & sign_mask else val)
                        tmp.append(sinf(f32(cc / iqfac) if cc >= 0 else f32(cc / iqfac_m)))
                    b = [f32(0)] * 21
                    for m in range(1, order + 1):
                        for i in range(1, m):
                            b[i] = f32(lpc[i - 1] + f32(tmp[m - 1] * lpc[m - i - 1]))
                        lpc[:m - 1] = b[1:m]
                        lpc[m - 1] = tmp[m - 1]
                filters.append((length, order, direction, lpc))
            out.append(filters)
        return out

    def apply_pulse(self):
        """Pulse::synth (pulse.rs:60-105)."""
        if self.pulse is None:
            return
        bands = self.bands()
        start, pulses = self.pulse
        if start >= len(bands) - 1:
            return
        k, band, c = bands[start], start, self.coeffs
        for off, amp in pulses:
            k += off
            if k >= 1024:
                return
            while bands[band + 1] <= k:
                band += 1
            scale = self.scales[0][band]
            base = c[k]
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                if base != 0:
                    if scale == 0:
                        base = f32(0)
                    else:
                        bval = f32(c[k] / scale)
                        base = powf(c[k], 0.75) if bval >= 0 else f32(-powf(f32(-c[k]), 0.75))
                base = f32(base + f32(amp)) if base > 0 else f32(base - f32(amp))
                p43 = f32(4.0) / f32(3.0)
                iq = f32(-powf(f32(-base), p43)) if base < 0 else powf(base, p43)
                c[k] = f32(iq * scale)

    def tns_filters(self, rate_idx):
        """The (start, end, order, direction, lpc) list Tns::synth walks (tns.rs:149-199), filters of order 0 left out."""
        if self.tns is None:
            return []
        bands = self.bands()
        max_bands = min((TNS_MAX_LONG if self.long_win else TNS_MAX_SHORT)[rate_idx], self.max_sfb)
        out = []
        for w, filters in enumerate(self.tns):
            bottom = len(bands) - 1
            for length, order, direction, lpc in filters:
                top = bottom
                bottom = max(top - length, 0)
                if order == 0:
                    continue
                out.append((w * 128 + bands[min(bottom, max_bands)], w * 128 + bands[min(top, max_bands)], order, int(direction), lpc[:20]))
        return out


def _sign(bit):
    return f32(f32(1.0) - f32(f32(2.0) * f32(bit)))


def _escape(bs):
    n = bs.read_unary_ones()
    if n >= 9:
        raise AacError(DECODE, "escape")
    return (1 << (n + 4)) + bs.read(n + 4)


class Lcg:
    def __init__(self):
        self.state = 0x1F2E3D4C

    def next(self):
        self.state = (self.state * 1664525 + 1013904223) & 0xFFFFFFFF
        return self.state  # callers take bits 16..31 as an i16


class Pair:
    def __init__(self, is_pair, channel, long_bands, short_bands):
        self.is_pair, self.channel = is_pair, channel
        self.ics = [Ics(long_bands, short_bands), Ics(long_bands, short_bands)]
        self.lcg = Lcg()
        self.ms_mask_present = 0
        self.ms_used = [[False] * 64 for _ in range(8)]

    def decode_cpe(self, bs):
        i0, i1 = self.ics
        common = bs.read_bool()
        if common:
            i0.decode_info(bs)
            self.ms_mask_present = bs.read(2)
            if self.ms_mask_present == 3:
                raise AacError(DECODE, "ms mask")
            for g in range(i0.window_groups):
                for sfb in range(i0.max_sfb):
                    self.ms_used[g][sfb] = bs.read_bool() if self.ms_mask_present == 1 else self.ms_mask_present == 2
            i1.copy_from_common(i0)
        i0.decode(bs, self.lcg, common)
        i1.decode(bs, self.lcg, common)
        if not common:
            return
        bands, g = i0.bands(), 0
        a, b = i0.coeffs, i1.coeffs
        for w in range(i0.num_windows):
            if w > 0 and not i0.grouping[w - 1]:
                g += 1
            for sfb in range(i0.max_sfb):
                lo, hi = w * 128 + bands[sfb], w * 128 + bands[sfb + 1]
                cb0, cb1 = i0.sfb_cb[g][sfb], i1.sfb_cb[g][sfb]
                if cb1 in (14, 15):
                    invert = self.ms_mask_present == 1 and self.ms_used[g][sfb]
                    scale = f32(f32(f32(1.0 if cb1 == 15 else -1.0) * f32(-1.0 if invert else 1.0)) * i1.scales[g][sfb])
                    b[lo:hi] = (scale * a[lo:hi]).astype(np.float32)
                elif cb0 == 13 or cb1 == 13:
                    pass
                elif self.ms_used[g][sfb]:
                    with np.errstate(over="ignore", invalid="ignore"):
                        tmp = (a[lo:hi] - b[lo:hi]).astype(np.float32)
                        a[lo:hi] = (a[lo:hi] + b[lo:hi]).astype(np.float32)
                        b[lo:hi] = tmp


class AacFrontend:
    """AacDecoder without extra data (ADTS parameters): `channels` 1 or 2."""

    def __init__(self, sample_rate, channels):
        if channels not in (1, 2):
            raise AacError(UNSUPPORTED, "channels")
        self.channels, self.rate_idx = channels, rate_index(sample_rate)
        _, self.long_bands, self.short_bands = SUBBAND_INFO[self.rate_idx]
        self.pairs = []

    def reset(self):
        for p in self.pairs:
            p.ics[0].reset(), p.ics[1].reset()

    def _set_pair(self, pair_no, channel, is_pair):
        if len(self.pairs) <= pair_no:
            self.pairs.append(Pair(is_pair, channel, self.long_bands, self.short_bands))
        elif self.pairs[pair_no].channel != channel or self.pairs[pair_no].is_pair != is_pair:
            raise AacError(DECODE, "element layout changed")
        if not (channel + 1 if is_pair else channel) < self.channels:
            raise AacError(DECODE, "too many channels")

    def decode(self, packet):
        """decode_inner / decode_ga (mod.rs:128-229).  Returns (channels covered, [per covered channel: dict(window_sequence,
        window_shape, prev_window_shape, tns [(start, end, order, direction, lpc[20])], coeffs [1024])])."""
        bs = BitsLtr(packet)
        cur_pair = cur_ch = 0
        while bs.bits_left() > 3:
            eid = bs.read(3)
            if eid in (0, 3):
                bs.read(4)
                self._set_pair(cur_pair, cur_ch, False)
                p = self.pairs[cur_pair]
                p.ics[0].decode(bs, p.lcg, False)
                cur_pair, cur_ch = cur_pair + 1, cur_ch + 1
            elif eid == 1:
                bs.read(4)
                self._set_pair(cur_pair, cur_ch, True)
                self.pairs[cur_pair].decode_cpe(bs)
                cur_pair, cur_ch = cur_pair + 1, cur_ch + 2
            elif eid == 2:
                raise AacError(UNSUPPORTED, "coupling channel element")
            elif eid == 4:
                bs.read(4)
                align = bs.read_bool()
                count = bs.read(8)
                if count == 255:
                    count += bs.read(8)
                if align:
                    bs.realign()
                bs.ignore(count * 8)
            elif eid == 5:
                raise AacError(UNSUPPORTED, "program config")
            elif eid == 6:
                count = bs.read(4)
                if count == 15:
                    count += bs.read(8) - 1
                if count > 0:
                    bs.read(4)
                    bs.ignore(4)
                    for _ in range(count - 1):
                        bs.ignore(8)
            else:
                break
        out = []
        for p in self.pairs[:cur_pair]:
            for ics in p.ics[:2 if p.is_pair else 1]:
                ics.apply_pulse()
                out.append(dict(window_sequence=ics.window_sequence, window_shape=int(ics.window_shape), prev_window_shape=int(ics.prev_window_shape),
                                tns=ics.tns_filters(self.rate_idx), coeffs=ics.coeffs.copy()))
        return cur_ch, out


# ---- AudioSpecificConfig::read, symphonia-common/src/mpeg/audio/mod.rs:230-439 (object types by MPEG-4 index, :87-130) -----------
ASC_RATES = [96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 7350]
ASC_GA = {1, 2, 3, 6, 7, 17, 19, 20, 21, 22, 23}
ASC_UNSUPPORTED = {8, 9, 12, 13, 14, 15, 16, 24, 25, 26, 27, 28, 30, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41}
ASC_ER = {17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 39}


def read_asc(buf):
    """dict(object_type, sample_rate, channels (0 = in-band), samples, sbr_present, ps_present, has_ext, ext_sample_rate, ext_channels)."""
    bs = BitsLtr(buf)

    def aot():
        v = bs.read(5)
        return v if v < 31 else bs.read(6) + 32

    def rate():
        i = bs.read(4)
        if i <= 12:
            return ASC_RATES[i]
        if i == 15:
            return bs.read(24)
        raise AacError(DECODE, "sample rate index")

    def chans():
        i = bs.read(4)
        if i > 7:
            raise AacError(DECODE, "channel configuration")
        return [0, 1, 2, 3, 4, 5, 6, 8][i]

    a = dict(object_type=aot(), sample_rate=rate(), samples=0, sbr_present=0, ps_present=0, has_ext=0, ext_sample_rate=0, ext_channels=0)
    if a["sample_rate"] == 0:
        raise AacError(DECODE, "sample rate 0")
    a["channels"] = chans()
    if a["object_type"] in (5, 29):
        a["sbr_present"], a["ps_present"], a["has_ext"] = 1, int(a["object_type"] == 29), 1
        a["ext_sample_rate"] = rate()
        a["object_type"] = aot()
        if a["object_type"] == 22:
            a["ext_channels"] = chans()
    t = a["object_type"]
    if t in ASC_GA:
        a["samples"] = 960 if bs.read_bool() else 1024
        if bs.read_bool():
            bs.read(14)
        ext = bs.read_bool()
        if a["channels"] == 0:
            raise AacError(UNSUPPORTED, "program config element")
        if t in (6, 20):
            bs.read(3)
        if ext:
            if t == 22:
                bs.read(5), bs.read(11)
            if t in (17, 19, 20, 23):
                bs.read_bool(), bs.read_bool(), bs.read_bool()
            if bs.read_bool():
                raise AacError(UNSUPPORTED, "version3 extensions")
    elif t in ASC_UNSUPPORTED:
        raise AacError(UNSUPPORTED, "object type")
    if t in ASC_ER:
        if bs.read(2) >= 2:
            raise AacError(UNSUPPORTED, "error protection")
    if a["has_ext"] and bs.bits_left() >= 16:
        if bs.read(11) == 0x2B7:
            e = aot()
            if e == 5:
                a["sbr_present"] = int(bs.read_bool())
                if a["sbr_present"]:
                    rate()
                    if bs.bits_left() >= 12 and bs.read(11) == 0x548:
                        a["ps_present"] = int(bs.read_bool())
            if e == 29:
                a["sbr_present"] = int(bs.read_bool())
                if a["sbr_present"]:
                    rate()
                bs.read(4)
    a["object_type"] = min(a["object_type"], 255)
    return a


def decoder_accepts(asc):
    """AacDecoder::try_new's judgement of a parsed configuration (aac/mod.rs:86-108): None, or the error kind."""
    if asc["channels"] == 0:
        return UNSUPPORTED
    if asc["object_type"] != 2 or asc["sbr_present"] or asc["channels"] > 2 or asc["samples"] != 1024:
        return UNSUPPORTED
    return None
