"""AAC-LC raw_data_block WRITER for the front-end tests (ISO/IEC 14496-3 4.4.2.1 raw_data_block, 4.4.2.7 individual_channel_stream,
4.6.2-4.6.3 quantisation and scale factors, 4.6.8 joint coding, 4.6.9 TNS, 4.6.13 PNS): single-channel and channel-pair elements with
random window sequences, grouping, section layouts (all eleven spectrum books, escapes, noise and intensity bands), scale-factor
differences, pulse data, TNS filters, data-stream and fill elements in between -- and the values a decoder must produce, kept as
ground truth from the numbers that were CODED, never from reading the bits back.  Builders only; nothing here reads a bitstream.

Where the reference departs from the standard the truth follows the reference, because identical results are the contract: scale
factors carry its 2^-14 output normalisation, pulses are added after dequantisation by its requantise / dequantise round trip
(aac/ics/pulse.rs:19-32, :60-105), noise bands come from its generator (aac/common.rs:96-111, seed cpe.rs:45)."""
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

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aac_huffman.json")) as _f:
    HUFF = json.load(_f)

# swb offsets by sampling-frequency class (ISO/IEC 14496-3 Tables 4.129-4.147), as (lower rate bound, long, short)
from oracle.aac_frontend_oracle import SUBBAND_INFO, TNS_MAX_LONG, TNS_MAX_SHORT  # noqa: E402  (tables of numbers only)

BOOK_RANGE = {1: (-1, 1), 2: (-1, 1), 3: (0, 2), 4: (0, 2), 5: (-4, 4), 6: (-4, 4), 7: (0, 7), 8: (0, 7), 9: (0, 12), 10: (0, 12), 11: (0, 8191)}


class BitWriterMsb:
    def __init__(self):
        self.bits = []

    def put(self, value, width):
        assert 0 <= value < (1 << width) or width == 0, (value, width)
        self.bits += [(value >> (width - 1 - k)) & 1 for k in range(width)]

    def huff(self, book, index):
        t = HUFF[book]
        self.put(t["codes"][index], t["lens"][index])

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(int("".join(map(str, b[i:i + 8])), 2) for i in range(0, len(b), 8))


def scale_normal(sf):
    """2^((sf - 100) / 4) with the reference's 2^-14 normalisation, correctly rounded."""
    return f32(2.0 ** (0.25 * (sf - 156)))


def scale_intensity(pos):
    return f32(0.5 ** (0.25 * (pos - 155)))


def pow43(v):
    """|v|^(4/3) as the C library's powf gives it -- the call the reference's table is built from (ics/mod.rs:44-50).  10 of the 8192
    entries are one unit in the last place away from the correctly rounded value under glibc (test_aac_frontend.py counts them),
    so the table is a property of the platform's libm in the reference as well."""
    return f32(_m.powf(float(v), float(f32(4.0) / f32(3.0))))


def pow43_correctly_rounded(v):
    return f32(float(v) ** float(f32(4.0) / f32(3.0)))


class _Lcg:
    def __init__(self):
        self.s = 0x1F2E3D4C

    def i16(self):
        self.s = (self.s * 1664525 + 1013904223) % (1 << 32)
        v = self.s >> 16
        return v - 65536 if v >= 32768 else v


class Channel:
    """What one channel of a stream remembers between packets."""

    def __init__(self):
        self.window_sequence, self.window_shape = 0, 0


class Stream:
    def __init__(self, rng, rate=44100, channels=2, layout=None):
        """layout: list of "sce" / "cpe" covering `channels` (default: one cpe for stereo, one sce for mono)."""
        self.rng, self.rate, self.channels = rng, rate, channels
        self.layout = layout or (["cpe"] if channels == 2 else ["sce"])
        self.rate_idx = next(i for i, (lo, _, _) in enumerate(SUBBAND_INFO) if rate >= lo)
        _, self.long_bands, self.short_bands = SUBBAND_INFO[self.rate_idx]
        self.ch = [Channel() for _ in range(channels)]
        self.lcg = [_Lcg() for _ in self.layout]

    def forget_windows(self):
        """What a decoder reset does to the truth: no previous window."""
        for c in self.ch:
            c.window_sequence, c.window_shape = 0, 0

    # ---- pieces ---------------------------------------------------------------------------------------------------------------
    def _info(self, w, force_seq=None):
        rng = self.rng
        seq = int(rng.integers(4)) if force_seq is None else force_seq
        shape = int(rng.integers(2))
        w.put(0, 1), w.put(seq, 2), w.put(shape, 1)
        if seq == 2:
            bands = self.short_bands
            max_sfb = int(rng.integers(0, len(bands)))
            w.put(max_sfb, 4)
            grouping = [int(rng.integers(2)) for _ in range(7)]
            for g in grouping:
                w.put(g, 1)
            groups, cur = [], [0]
            for i, g in enumerate(grouping):
                if g:
                    cur.append(i + 1)
                else:
                    groups.append(cur)
                    cur = [i + 1]
            groups.append(cur)
        else:
            bands = self.long_bands
            max_sfb = int(rng.integers(0, len(bands))) if rng.random() < 0.8 else len(bands) - 1
            w.put(max_sfb, 6)
            w.put(0, 1)  # predictor_data_present
            groups = [[0]]
        return dict(seq=seq, shape=shape, max_sfb=max_sfb, groups=groups, bands=bands, long=seq != 2)

    def _ics(self, w, info, global_gain, allow_intensity, lcg):
        """Writes one individual_channel_stream after its ics_info; returns dict(cb [g][sfb], scale [g][sfb], coeffs, pulse, tns)."""
        rng = self.rng
        G, max_sfb, bands = len(info["groups"]), info["max_sfb"], info["bands"]
        # section data
        cbs = [[0] * max_sfb for _ in range(G)]
        bits = 5 if info["long"] else 3
        esc = (1 << bits) - 1
        choices = [0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 11, 13] + ([14, 15, 15] if allow_intensity else [])
        for g in range(G):
            k = 0
            while k < max_sfb:
                if rng.random() < 0.03:
                    w.put(int(rng.integers(12)), 4), w.put(0, bits)  # an empty section
                    continue
                cb = int(rng.choice(choices))
                n = min(int(rng.geometric(0.35)), max_sfb - k)
                if rng.random() < 0.1:
                    n = max_sfb - k
                w.put(cb, 4)
                left = n
                while left >= esc:
                    w.put(esc, bits)
                    left -= esc
                w.put(left, bits)
                for s in range(k, k + n):
                    cbs[g][s] = cb
                k += n
        # scale factors
        sf = global_gain
        pos = 155          # index of the intensity table: position + 155
        noise = global_gain + 10
        first_noise = True
        scale = [[f32(0)] * max_sfb for _ in range(G)]
        for g in range(G):
            for s in range(max_sfb):
                cb = cbs[g][s]
                if cb == 0:
                    continue
                if cb in (14, 15):
                    d = int(np.clip(rng.integers(-12, 13), -pos, 255 - pos))
                    pos += d
                    w.huff("scf", d + 60)
                    scale[g][s] = scale_intensity(pos)
                elif cb == 13:
                    if first_noise:
                        first_noise = False
                        d = int(np.clip(rng.integers(-40, 41), -noise, 255 - noise))
                        w.put(d + 256, 9)
                    else:
                        d = int(np.clip(rng.integers(-12, 13), -noise, 255 - noise))
                        w.huff("scf", d + 60)
                    noise += d
                    scale[g][s] = scale_normal(noise)
                else:
                    d = int(np.clip(rng.integers(-15, 16), max(-60, 60 - sf), min(60, 230 - sf)))
                    sf += d
                    w.huff("scf", d + 60)
                    scale[g][s] = scale_normal(sf)
        # pulse data
        pulse = None
        if info["long"] and rng.random() < 0.3:
            n = int(rng.integers(1, 5))
            start = int(rng.integers(0, 64)) if rng.random() < 0.2 else int(rng.integers(0, len(bands)))
            pulse = (start, [(int(rng.integers(32)), int(rng.integers(16))) for _ in range(n)])
            w.put(1, 1), w.put(n - 1, 2), w.put(start, 6)
            for off, amp in pulse[1]:
                w.put(off, 5), w.put(amp, 4)
        else:
            w.put(0, 1)
        # TNS data
        tns = None
        if rng.random() < 0.4:
            w.put(1, 1)
            tns = []
            for _win in range(1 if info["long"] else 8):
                n_filt = int(rng.integers(0, 4 if info["long"] else 2))
                w.put(n_filt, 2 if info["long"] else 1)
                coef_res = int(rng.integers(2))
                if n_filt:
                    w.put(coef_res, 1)
                filters = []
                for _ in range(n_filt):
                    length = int(rng.integers(0, 64 if info["long"] else 16))
                    order = int(rng.integers(0, 13 if info["long"] else 8))
                    w.put(length, 6 if info["long"] else 4), w.put(order, 5 if info["long"] else 3)
                    direction, lpc = 0, [f32(0)] * 20
                    if order:
                        direction, compress = int(rng.integers(2)), int(rng.integers(2))
                        w.put(direction, 1), w.put(compress, 1)
                        res_bits = 3 + coef_res - compress
                        raw = [int(rng.integers(-(1 << (res_bits - 1)), 1 << (res_bits - 1))) for _ in range(order)]
                        for c in raw:
                            w.put(c & ((1 << res_bits) - 1), res_bits)
                        lpc = _tns_lpc(raw, coef_res)
                    filters.append((length, order, direction, lpc))
                tns.append(filters)
        else:
            w.put(0, 1)
        w.put(0, 1)  # gain_control_data_present
        # spectral data, in the order (group, band, window of the group)
        coeffs = np.zeros(1024, dtype=np.float32)
        for g, wins in enumerate(info["groups"]):
            for s in range(max_sfb):
                cb, sc = cbs[g][s], scale[g][s]
                lo, hi = bands[s], bands[s + 1]
                for win in wins:
                    at = 128 * win + lo
                    n = hi - lo
                    if cb in (0, 14, 15):
                        continue
                    if cb == 13:
                        vals = np.array([lcg.i16() for _ in range(n)], dtype=np.float32)
                        energy = np.cumsum(vals * vals, dtype=np.float32)[-1]
                        with np.errstate(divide="ignore", invalid="ignore"):
                            coeffs[at:at + n] = vals * f32(sc / np.sqrt(energy))
                        continue
                    lo_v, hi_v = BOOK_RANGE[cb]
                    if cb == 11:
                        q = rng.integers(0, 17, n)
                        big = rng.random(n) < 0.15
                        q = np.where(big, rng.integers(16, 8192, n), np.minimum(q, 15))
                        if rng.random() < 0.1:
                            q[int(rng.integers(n))] = int(rng.choice([16, 31, 32, 8191, 4096, 4095]))
                        q = q * rng.choice([-1, 1], n)
                    else:
                        q = rng.integers(lo_v, hi_v + 1, n)
                        if lo_v == 0:
                            q = q * rng.choice([-1, 1], n)
                    q = [int(v) for v in q]
                    if cb <= 4:
                        for i in range(0, n, 4):
                            d = q[i:i + 4]
                            if cb <= 2:
                                w.huff(str(cb), 27 * (d[0] + 1) + 9 * (d[1] + 1) + 3 * (d[2] + 1) + (d[3] + 1))
                            else:
                                w.huff(str(cb), 27 * abs(d[0]) + 9 * abs(d[1]) + 3 * abs(d[2]) + abs(d[3]))
                                for v in d:
                                    if v:
                                        w.put(int(v < 0), 1)
                    else:
                        for i in range(0, n, 2):
                            a, b = q[i], q[i + 1]
                            if cb <= 6:
                                w.huff(str(cb), 9 * (a + 4) + (b + 4))
                            elif cb <= 10:
                                mod = 8 if cb <= 8 else 13
                                w.huff(str(cb), mod * abs(a) + abs(b))
                                for v in (a, b):
                                    if v:
                                        w.put(int(v < 0), 1)
                            else:
                                w.huff("11", 17 * min(abs(a), 16) + min(abs(b), 16))
                                for v in (a, b):
                                    if v:
                                        w.put(int(v < 0), 1)
                                for v in (abs(a), abs(b)):
                                    if v >= 16:
                                        nbits = v.bit_length() - 1      # v = 2^nbits + rest, nbits = n + 4
                                        for _ in range(nbits - 4):
                                            w.put(1, 1)
                                        w.put(0, 1)
                                        w.put(v - (1 << nbits), nbits)
                    for i, v in enumerate(q):
                        if v == 0:
                            continue
                        sgn = f32(-1.0 if v < 0 else 1.0)
                        if cb in (3, 4) and abs(v) == 2:
                            coeffs[at + i] = f32(sgn * f32(f32(2.51984209978974632953) * sc))
                        else:
                            coeffs[at + i] = f32(f32(sgn * pow43(abs(v))) * sc)
        return dict(cb=cbs, scale=scale, coeffs=coeffs, pulse=pulse, tns=tns)

    def _finish(self, chan, info, ics):
        """Pulse restoration, TNS ranges and the window history -> the truth record of one channel."""
        c = ics["coeffs"]
        bands = info["bands"]
        if ics["pulse"] is not None:
            start, pulses = ics["pulse"]
            if start < len(bands) - 1:
                k, band = bands[start], start
                for off, amp in pulses:
                    k += off
                    if k >= 1024:
                        break
                    while bands[band + 1] <= k:
                        band += 1
                    sc = ics["scale"][0][band] if band < info["max_sfb"] else self._stale_scale(chan, band)
                    base = c[k]
                    with np.errstate(all="ignore"):
                        if base != 0:
                            if sc == 0:
                                base = f32(0)
                            else:
                                base = f32(_m.powf(float(c[k]), 0.75)) if f32(c[k] / sc) >= 0 else f32(-f32(_m.powf(float(-c[k]), 0.75)))
                        base = f32(base + f32(amp)) if base > 0 else f32(base - f32(amp))
                        p = float(f32(4.0) / f32(3.0))
                        iq = f32(-f32(_m.powf(float(-base), p))) if base < 0 else f32(_m.powf(float(base), p))
                        c[k] = f32(iq * sc)
        tns = []
        if ics["tns"] is not None:
            cap = min((TNS_MAX_LONG if info["long"] else TNS_MAX_SHORT)[self.rate_idx], info["max_sfb"])
            for win, filters in enumerate(ics["tns"]):
                bottom = len(bands) - 1
                for length, order, direction, lpc in filters:
                    top = bottom
                    bottom = max(top - length, 0)
                    if order:
                        tns.append((128 * win + bands[min(bottom, cap)], 128 * win + bands[min(top, cap)], order, direction, lpc))
        st = self.ch[chan]
        rec = dict(window_sequence=info["seq"], window_shape=info["shape"], prev_window_shape=st.window_shape, tns=tns, coeffs=c)
        st.window_sequence, st.window_shape = info["seq"], info["shape"]
        return rec

    def _stale_scale(self, chan, band):
        """A pulse may land in a band above max_sfb, whose scale the reference still holds from an earlier frame; the writer avoids
        relying on that by keeping pulses inside the coded bands (see packet())."""
        raise AssertionError("pulse outside the coded bands")

    # ---- one raw_data_block ---------------------------------------------------------------------------------------------------
    def packet(self, extras=True):
        rng = self.rng
        while True:
            try:
                return self._packet(extras)
            except AssertionError as e:
                if "pulse outside" not in str(e):
                    raise
                # roll the window history back is not needed: _finish raises before touching it for the failing channel only when
                # that channel's pulse is out of range -- regenerate with fresh random numbers
                continue

    def _packet(self, extras):
        rng = self.rng
        w = BitWriterMsb()
        saved = [(c.window_sequence, c.window_shape) for c in self.ch]
        saved_lcg = [l.s for l in self.lcg]
        try:
            truth = []
            chan = 0
            for k, kind in enumerate(self.layout):
                if extras and rng.random() < 0.25:
                    self._extra(w)
                if kind == "sce":
                    w.put(int(rng.choice([0, 0, 0, 3])), 3), w.put(int(rng.integers(16)), 4)
                    gg = int(rng.integers(110, 200))
                    w.put(gg, 8)
                    info = self._info(w)
                    ics = self._ics(w, info, gg, False, self.lcg[k])
                    truth.append(self._finish(chan, info, ics))
                    chan += 1
                else:
                    w.put(1, 3), w.put(int(rng.integers(16)), 4)
                    common = int(rng.integers(2))
                    w.put(common, 1)
                    if common:
                        info0 = info1 = self._info(w)
                        G, max_sfb = len(info0["groups"]), info0["max_sfb"]
                        mask = int(rng.integers(3))
                        w.put(mask, 2)
                        ms = [[mask == 2] * max_sfb for _ in range(G)]
                        if mask == 1:
                            for g in range(G):
                                for s in range(max_sfb):
                                    ms[g][s] = bool(rng.integers(2))
                                    w.put(int(ms[g][s]), 1)
                    gg0 = int(rng.integers(110, 200))
                    w.put(gg0, 8)
                    if not common:
                        info0 = self._info(w)
                    ics0 = self._ics(w, info0, gg0, not common, self.lcg[k])
                    gg1 = int(rng.integers(110, 200))
                    w.put(gg1, 8)
                    if not common:
                        info1 = self._info(w)
                    ics1 = self._ics(w, info1, gg1, True, self.lcg[k])
                    if common:
                        a, b, bands = ics0["coeffs"], ics1["coeffs"], info0["bands"]
                        for g, wins in enumerate(info0["groups"]):
                            for s in range(info0["max_sfb"]):
                                for win in wins:
                                    lo, hi = 128 * win + bands[s], 128 * win + bands[s + 1]
                                    c0, c1 = ics0["cb"][g][s], ics1["cb"][g][s]
                                    if c1 in (14, 15):
                                        sign = (1.0 if c1 == 15 else -1.0) * (-1.0 if (mask == 1 and ms[g][s]) else 1.0)
                                        b[lo:hi] = f32(f32(sign) * ics1["scale"][g][s]) * a[lo:hi]
                                    elif c0 == 13 or c1 == 13:
                                        pass
                                    elif ms[g][s]:
                                        with np.errstate(all="ignore"):
                                            l, r = a[lo:hi].copy(), b[lo:hi].copy()
                                            a[lo:hi], b[lo:hi] = l + r, l - r
                    truth.append(self._finish(chan, info0, ics0))
                    truth.append(self._finish(chan + 1, info1, ics1))
                    chan += 2
            if extras and rng.random() < 0.25:
                self._extra(w)
            w.put(7, 3)
            return w.bytes(), truth
        except AssertionError:
            for c, (s, sh) in zip(self.ch, saved):
                c.window_sequence, c.window_shape = s, sh
            for l, s in zip(self.lcg, saved_lcg):
                l.s = s
            raise

    def _extra(self, w):
        rng = self.rng
        if rng.random() < 0.5:  # data_stream_element
            w.put(4, 3), w.put(int(rng.integers(16)), 4)
            align, count = int(rng.integers(2)), int(rng.choice([0, 3, 255, 260]))
            w.put(align, 1)
            if count >= 255:
                w.put(255, 8), w.put(count - 255, 8)
            else:
                w.put(count, 8)
            if align:
                w.align()
            for _ in range(count):
                w.put(int(rng.integers(256)), 8)
        else:  # fill_element
            count = int(rng.choice([0, 1, 5, 14, 15, 40]))
            w.put(6, 3)
            if count >= 15:
                w.put(15, 4), w.put(count - 15 + 1, 8)
            else:
                w.put(count, 4)
            if count:
                w.put(int(rng.choice([0, 1, 2, 11])), 4), w.put(int(rng.integers(16)), 4)
                for _ in range(count - 1):
                    w.put(int(rng.integers(256)), 8)


def _tns_lpc(raw, coef_res):
    """ISO/IEC 14496-3 4.6.9.3: inverse quantisation of the reflection coefficients and the step-up to LPC coefficients (f32)."""
    fac = f32(8.0 if coef_res else 4.0)
    half_pi = f32(np.pi / 2)
    iqfac, iqfac_m = f32(f32(fac - f32(0.5)) / half_pi), f32(f32(fac + f32(0.5)) / half_pi)
    tmp = [f32(_m.sinf(float(f32(f32(c) / (iqfac if c >= 0 else iqfac_m))))) for c in raw]
    a = [f32(0)] * 20
    for m in range(1, len(raw) + 1):
        b = [f32(a[i - 1] + f32(tmp[m - 1] * a[m - i - 1])) for i in range(1, m)]
        a[:m - 1] = b
        a[m - 1] = tmp[m - 1]
    return a
