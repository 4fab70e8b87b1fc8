"""FLAC frame WRITER for the front-end tests: serialises the descriptor form of `workloads.flac_batch` (sub-frame
types, predictor orders, quantised coefficients, warm-up samples and residuals produced by an exact-integer encoder)
into frames as the FLAC format document lays them out (frame header with its CRC-8, sub-frame headers with wasted
bits, Rice / Rice2 partitions incl. escaped ones, zero padding, CRC-16).  Builders only."""
import numpy as np

from tests._mp3_bitstream import BitWriterMsb

CONSTANT, VERBATIM, FIXED, LPC = 0, 1, 2, 3
_CH_CODE = {1: 8, 3: 9, 2: 10}  # left/side, right/side, mid/side (assignment numbers of include/symgpu.h)
_BPS_CODE = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def crc8(buf):
    c = 0
    for b in buf:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(buf):
    c = 0
    for b in buf:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def utf8_encode(v):
    if v < 0x80:
        return bytes([v])
    for n, lead in ((2, 0xC0), (3, 0xE0), (4, 0xF0), (5, 0xF8), (6, 0xFC), (7, 0xFE)):
        if v < 1 << (5 * n + 1 if n < 7 else 36):
            out = [(v >> (6 * k)) & 0x3F | 0x80 for k in range(n - 1)][::-1]
            return bytes([lead | (v >> (6 * (n - 1)))] + out)
    raise ValueError(v)


def _put_signed(w, v, bits):
    if bits:
        assert -(1 << (bits - 1)) <= v < (1 << (bits - 1)), (v, bits)
        w.put(v & ((1 << bits) - 1), bits)


def _put_residuals(w, rng, res, order, n, force_po=None):
    big = max((abs(int(x)) for x in res[order:]), default=0) >= 1 << 13
    method = 1 if big else int(rng.integers(2))  # 4-bit parameters stop at 14: large residuals need the 5-bit form
    width = 5 if method else 4
    choices = [po for po in range(0, 9) if n % (1 << po) == 0 and (n >> po) >= max(order, 1)]
    po = (int(rng.choice(choices)) if choices else 0) if force_po is None else force_po
    w.put(method, 2), w.put(po, 4)
    per = n >> po
    for part in range(1 << po):
        a, b = (part * per if part else min(order, per)), (part + 1) * per
        seg = [int(x) for x in res[a:b]]
        zig = [(x << 1) ^ (x >> 63) for x in seg]
        mean = (sum(zig) / len(zig)) if zig else 0
        k = min(int(np.log2(mean + 1)), (1 << width) - 2)
        bits = max([x.bit_length() + 1 for x in seg] + [0]) if any(seg) else int(rng.integers(2))  # 0 bits: all residuals are zero
        long_runs = any((z >> k) > 150 for z in zig)
        if bits <= 31 and (rng.integers(7) == 0 or long_runs):  # escaped partition: plain two's complement
            w.put((1 << width) - 1, width), w.put(bits, 5)
            for x in seg:
                _put_signed(w, x, bits)
        else:
            if long_runs:
                k = min(max(k, max(zig).bit_length() - 6), (1 << width) - 2)
            w.put(k, width)
            for z in zig:
                w.put(1, (z >> k) + 1)  # q zeros then a one
                w.put(z & ((1 << k) - 1), k)


def write_frame(rng, frame, subs, samples, number, sample_rate=44100, stream_bps=None, lpc_precision=None, force_po=None):
    """One frame.  `subs` = the frame's sub-frame records, `samples` the global residual / warm-up array they point into."""
    n = int(subs[0]["n"])
    bps = int(frame["bits_per_sample"])
    channels, assignment = int(frame["channels"]), int(frame["assignment"])
    head = BitWriterMsb()
    head.put(0xFFF8 >> 2, 14), head.put(0, 1), head.put(0, 1)  # sync, reserved, fixed block size
    tail = b""
    fixed_codes = {192: 1, **{576 << k: 2 + k for k in range(4)}, **{256 << k: 8 + k for k in range(8)}}
    if n in fixed_codes and rng.integers(3):
        bs_code = fixed_codes[n]
    elif n <= 256 and rng.integers(2):
        bs_code, tail = 6, bytes([n - 1])
    else:
        bs_code, tail = 7, (n - 1).to_bytes(2, "big")
    sr_kind = int(rng.integers(5))
    if sr_kind == 0:
        sr_code, sr_tail = 0, b""
    elif sr_kind == 1 and sample_rate == 44100:
        sr_code, sr_tail = 9, b""
    elif sr_kind == 2 and sample_rate % 1000 == 0 and sample_rate // 1000 < 256:
        sr_code, sr_tail = 12, bytes([sample_rate // 1000])
    elif sr_kind == 3 and sample_rate % 10 == 0:
        sr_code, sr_tail = 14, (sample_rate // 10).to_bytes(2, "big")
    else:
        sr_code, sr_tail = 13, sample_rate.to_bytes(2, "big")
    bps_code = 0 if (stream_bps == bps and rng.integers(3) == 0) else _BPS_CODE[bps]
    head.put(bs_code, 4), head.put(sr_code, 4)
    head.put(channels - 1 if assignment == 0 else _CH_CODE[assignment], 4), head.put(bps_code, 3), head.put(0, 1)
    hdr = head.bytes() + utf8_encode(number) + tail + sr_tail
    hdr += bytes([crc8(hdr)])
    w = BitWriterMsb()
    for c in range(channels):
        sf = subs[c]
        side = (assignment, c) in ((1, 1), (2, 1), (3, 0))
        kind, order, wasted = int(sf["type"]), int(sf["order"]), int(sf["wasted"])
        sub_bps = bps + int(side) - wasted
        off = int(sf["offset"])
        x = samples[off:off + n].astype(np.int64)
        w.put(0, 1)
        w.put({CONSTANT: 0, VERBATIM: 1, FIXED: 8 | order, LPC: 32 | (order - 1)}[kind], 6)
        if wasted:
            w.put(1, 1), w.put(1, wasted)  # wasted - 1 zeros, then a one
        else:
            w.put(0, 1)
        if kind == CONSTANT:
            _put_signed(w, int(x[0]), sub_bps)
        elif kind == VERBATIM:
            for v in x:
                _put_signed(w, int(v), sub_bps)
        else:
            for v in x[:order]:
                _put_signed(w, int(v), sub_bps)
            if kind == LPC:
                coeffs = [int(v) for v in sf["coeffs"][:order]]
                precision = max(max((v.bit_length() + 1 for v in coeffs), default=1), 1)
                precision = min(max(precision, int(rng.integers(1, 16))), 15) if lpc_precision is None else lpc_precision
                w.put(precision - 1, 4), w.put(int(sf["shift"]), 5)
                for v in coeffs:
                    _put_signed(w, v, precision)
            _put_residuals(w, rng, x, order, n, force_po)
    body = w.bytes()
    frame_bytes = hdr + body
    return frame_bytes + crc16(frame_bytes).to_bytes(2, "big")


def stream_info_block(block_min, block_max, sample_rate, channels, bps, n_samples, frame_min=0, frame_max=0, md5=bytes(16)):
    """STREAMINFO body, 34 bytes (FLAC format: METADATA_BLOCK_STREAMINFO)."""
    bits = (sample_rate << 44) | ((channels - 1) << 41) | ((bps - 1) << 36) | n_samples
    return block_min.to_bytes(2, "big") + block_max.to_bytes(2, "big") + frame_min.to_bytes(3, "big") + frame_max.to_bytes(3, "big") + \
        bits.to_bytes(8, "big") + md5


def native_file(frames, info_block, extra_blocks=()):
    """"fLaC", STREAMINFO, optional further metadata blocks (type, body), frames."""
    blocks = [(0, info_block)] + list(extra_blocks)
    out = b"fLaC"
    for k, (kind, body) in enumerate(blocks):
        out += bytes([(0x80 if k == len(blocks) - 1 else 0) | kind]) + len(body).to_bytes(3, "big") + body
    return out + b"".join(frames)
