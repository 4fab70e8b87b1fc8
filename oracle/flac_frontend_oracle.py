"""FLAC front-end oracle (SURVEY §8f N1 for the FLAC row): what FlacDecoder::decode_inner reads from a packet, in the
reference's sequence.  TEST INFRASTRUCTURE ONLY.

  symphonia-bundle-flac/src/frame.rs:66-233 (sync, header, CRC-8), :281-333 (extended UTF-8)
  decoder.rs:139-228 (frame), :340-520 (sub-frames), :522-640 (residuals), symphonia-core/src/io/bit.rs:642-671 (unary)

Pinned by the reference's own known answers: verify_utf8_decode_be_u64 (frame.rs:335-355) and verify_rice_signed_to_i32
(decoder.rs:642-658), the CRC-8 catalogue check value, and lossless round trips (tests/test_flac_frontend.py)."""
from oracle.mp3_frontend_oracle import BitsLtr, DecodeError

CONSTANT, VERBATIM, FIXED, LPC = 0, 1, 2, 3
INDEPENDENT, LEFT_SIDE, MID_SIDE, RIGHT_SIDE = 0, 1, 2, 3


class Unsupported(Exception):
    pass


def crc8(buf, state=0):
    """crc8.rs:32-65: polynomial 0x07, no reflection."""
    for b in buf:
        state ^= b
        for _ in range(8):
            state = ((state << 1) ^ 0x07) & 0xFF if state & 0x80 else (state << 1) & 0xFF
    return state


def utf8_decode(data, at):
    """frame.rs:281-333.  Returns (value or None, next position); raises DecodeError at the end of the data."""
    if at >= len(data):
        raise DecodeError("eof")
    state = data[at]
    at += 1
    if state <= 0x7F:
        return state, at
    for lo, hi, mask in ((0xC0, 0xDF, 0x1F), (0xE0, 0xEF, 0x0F), (0xF0, 0xF7, 0x07), (0xF8, 0xFB, 0x03), (0xFC, 0xFD, 0x01), (0xFE, 0xFE, 0x00)):
        if lo <= state <= hi:
            break
    else:
        return None, at
    state &= mask
    leading_zeros = 8 - mask.bit_length()
    for _ in range(2, leading_zeros):
        if at >= len(data):
            raise DecodeError("eof")
        state = (state << 6) | (data[at] & 0x3F)
        at += 1
    return state, at


def rice_signed(word):
    """decoder.rs:612-640."""
    div2 = word >> 1
    return ~div2 if word & 1 else div2


def sign_extend(v, bits):
    if bits == 0:
        return 0
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def _unary(bs):
    """bit.rs:642-671."""
    n = 0
    while True:
        if bs.read(1):
            return n
        n += 1


def read_frame_header(data, at):
    """frame.rs:81-233.  `at` = position of the sync code.  Returns (header dict, position after the CRC-8)."""
    start = at

    def take(k):
        nonlocal at
        if at + k > len(data):
            raise DecodeError("eof")
        v = int.from_bytes(data[at:at + k], "big")
        at += k
        return v
    sync = take(2)
    desc = take(2)
    if desc & 1:
        raise DecodeError("reserved bit")
    by_sample = bool(sync & 1)
    seq, at = utf8_decode(data, at)
    if seq is None or seq > (0x000FFFFFFFFF if by_sample else 0x7FFFFFFF):
        raise DecodeError("sequence number")
    bs_enc, sr_enc, ch_enc, bps_enc = desc >> 12, (desc >> 8) & 15, (desc >> 4) & 15, (desc >> 1) & 7
    if bs_enc == 1:
        block = 192
    elif 2 <= bs_enc <= 5:
        block = 576 * (1 << (bs_enc - 2))
    elif bs_enc == 6:
        block = take(1) + 1
    elif bs_enc == 7:
        v = take(2)
        if v == 0xFFFF:
            raise DecodeError("block size")
        block = v + 1
    elif bs_enc >= 8:
        block = 256 * (1 << (bs_enc - 8))
    else:
        raise DecodeError("block size reserved")
    fixed = {1: 88200, 2: 176400, 3: 192000, 4: 8000, 5: 16000, 6: 22050, 7: 24000, 8: 32000, 9: 44100, 10: 48000, 11: 96000}
    if sr_enc == 0:
        rate = None
    elif sr_enc in fixed:
        rate = fixed[sr_enc]
    elif sr_enc == 12:
        rate = take(1) * 1000
    elif sr_enc == 13:
        rate = take(2)
    elif sr_enc == 14:
        rate = take(2) * 10
    else:
        raise DecodeError("sample rate reserved")
    if rate is not None and not 1 <= rate <= 655350:
        raise DecodeError("sample rate bounds")
    widths = {0: None, 1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}
    if bps_enc not in widths:
        raise DecodeError("bits per sample reserved")
    if ch_enc <= 7:
        channels, assignment = ch_enc + 1, INDEPENDENT
    elif ch_enc in (8, 9, 10):
        channels, assignment = 2, {8: LEFT_SIDE, 9: RIGHT_SIDE, 10: MID_SIDE}[ch_enc]
    else:
        raise DecodeError("channel assignment reserved")
    computed = crc8(data[start:at])
    if take(1) != computed:
        raise DecodeError("header crc")
    return dict(sequence=seq, by_sample=by_sample, block=block, rate=rate, bps=widths[bps_enc], channels=channels, assignment=assignment), at


def _residual(bs, prelude, buf):
    """decoder.rs:522-607."""
    method = bs.read(2)
    if method > 1:
        raise DecodeError("residual method")
    width = 5 if method else 4
    order = bs.read(4)
    per = len(buf) >> order
    if prelude > per or (per << order) != len(buf):
        raise DecodeError("partitions")
    for part in range(1 << order):
        a, b = (part * per if part else prelude), (part + 1) * per
        param = bs.read(width)
        if param < (1 << width) - 1:
            for i in range(a, b):
                q = _unary(bs)
                r = bs.read(param)
                buf[i] = rice_signed(((q << param) | r) & 0xFFFFFFFF)
        else:
            bits = bs.read(5)
            for i in range(a, b):
                buf[i] = sign_extend(bs.read(bits), bits)


def read_subframe(bs, frame_bps, n):
    """decoder.rs:340-520.  Returns a dict: type, order, shift, wasted, coeffs (coefficient j multiplies sample i-1-j), samples."""
    if bs.read(1):
        raise DecodeError("padding")
    enc = bs.read(6)
    order = 0
    if enc == 0:
        kind = CONSTANT
    elif enc == 1:
        kind = VERBATIM
    elif 8 <= enc <= 15:
        order = enc & 7
        if order > 4:
            raise DecodeError("fixed order")
        kind = FIXED
    elif enc >= 32:
        order = (enc & 31) + 1
        kind = LPC
    else:
        raise DecodeError("reserved type")
    wasted = _unary(bs) + 1 if bs.read(1) else 0
    if wasted > frame_bps:
        raise DecodeError("wasted bits")
    bps = frame_bps - wasted
    if bps > 32:
        raise Unsupported("33-bit difference channel")
    buf = [0] * n
    coeffs, shift = [0] * 32, 0
    if kind == CONSTANT:
        buf[0] = sign_extend(bs.read(bps), bps)
    elif kind == VERBATIM:
        for i in range(n):
            buf[i] = sign_extend(bs.read(bps), bps)
    else:
        if order > n:
            raise DecodeError("order > block")
        for i in range(order):
            buf[i] = sign_extend(bs.read(bps), bps)
        if kind == LPC:
            precision = bs.read(4) + 1
            if precision > 15:
                raise DecodeError("precision")
            shift = sign_extend(bs.read(5), 5)
            if shift < 0:
                raise Unsupported("negative shift")
            for j in range(order):
                coeffs[j] = sign_extend(bs.read(precision), precision)
        _residual(bs, order, buf)
    return dict(type=kind, order=order, shift=shift, wasted=wasted, coeffs=coeffs, samples=buf)


def decode_packet(packet, stream_bps=0, stream_channels=0, max_block=0):
    """decoder.rs:139-228 up to the restoration.  Returns (header, [sub-frame dicts]) or raises DecodeError / Unsupported."""
    at = 0
    while True:
        if at + 2 > len(packet):
            raise DecodeError("no sync")
        if packet[at] == 0xFF and (packet[at + 1] & 0xFC) == 0xF8:
            break
        at += 1
    h, at = read_frame_header(packet, at)
    bps = h["bps"] or stream_bps
    if not bps or bps > 32:
        raise DecodeError("bits per sample")
    if max_block and h["block"] > max_block:
        raise DecodeError("block size over the stream's maximum")
    if stream_channels and h["channels"] > stream_channels:
        raise DecodeError("channel count")
    h["bps"] = bps
    bs = BitsLtr(packet[at:])
    subs = []
    for c in range(h["channels"]):
        side = (h["assignment"], c) in ((LEFT_SIDE, 1), (MID_SIDE, 1), (RIGHT_SIDE, 0))
        subs.append(read_subframe(bs, bps + int(side), h["block"]))
    return h, subs
