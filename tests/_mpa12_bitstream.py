"""Layer I / II bitstream WRITER for the sample-decoder tests (ISO/IEC 11172-3 2.4.1.5 / 2.4.1.6 and Annex B Tables
3-B.2a-d, 13818-3 Table B.1): random allocations, scale factor selections, scale factors and raw sample codes packed
into frames of the right size.  Builders only."""
import numpy as np

from tests import _streams as st
from tests._mp3_bitstream import BitWriterMsb

# per table: list of (nbal, [levels by allocation index]) per sub-band up to sblimit, written from the standard's tables
_L = lambda *xs: list(xs)
_FULL = [3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535]
_ROW_A0 = (4, [0] + _FULL)                                                            # sub-bands 0-2 of tables a / b
_ROW_A1 = (4, [0, 3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 65535])  # 3-10
_ROW_A2 = (3, [0, 3, 5, 7, 9, 15, 31, 65535])                                         # 11-22
_ROW_A3 = (2, [0, 3, 5, 65535])                                                       # 23-
_ROW_C0 = (4, [0, 3, 5, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767])  # tables c / d, 0-1
_ROW_C1 = (3, [0, 3, 5, 9, 15, 31, 63, 127])
_ROW_L0 = (4, [0, 3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383])  # 13818-3 B.1, 0-3
_ROW_L2 = (2, [0, 3, 5, 9])
TABLES = {
    "a": [_ROW_A0] * 3 + [_ROW_A1] * 8 + [_ROW_A2] * 12 + [_ROW_A3] * 4,
    "b": [_ROW_A0] * 3 + [_ROW_A1] * 8 + [_ROW_A2] * 12 + [_ROW_A3] * 7,
    "c": [_ROW_C0] * 2 + [_ROW_C1] * 6,
    "d": [_ROW_C0] * 2 + [_ROW_C1] * 10,
    "lsf": [_ROW_L0] * 4 + [_ROW_C1] * 7 + [_ROW_L2] * 19,
}
KBPS_L2 = [32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384]


def layer2_table(version, bitrate_idx, rate_idx, n_ch):
    """Which allocation table a frame uses (ISO 11172-3 2.4.2.3 / Table 3-B.2 headings)."""
    if version != "1":
        return "lsf"
    per = KBPS_L2[bitrate_idx - 1] // n_ch
    rate = [44100, 48000, 32000][rate_idx]
    if per <= 48:
        return "d" if rate == 32000 else "c"
    if per <= 80:
        return "a"
    return "a" if rate == 48000 else "b"


def gen_layer1_frame(rng, version="1", bitrate_idx=9, rate_idx=0, mode=0, mode_ext=0, protected=False, density=0.7):
    n_ch = 1 if mode == 3 else 2
    bound = (mode_ext + 1) * 4 if mode == 1 else 32
    pad = int(rng.integers(2))
    total = st.mpa_frame_len(version, 1, bitrate_idx, rate_idx, pad)
    budget = (total - 4 - (2 if protected else 0)) * 8
    while True:
        alloc = [[0] * 32, [0] * 32]
        for sb in range(32):
            for ch in range(n_ch if sb < bound else 1):
                alloc[ch][sb] = int(rng.integers(0, 15)) if rng.random() < density else 0  # field value; 15 is forbidden
            if sb >= bound:
                alloc[1][sb] = alloc[0][sb]
        need = 4 * sum(n_ch if sb < bound else 1 for sb in range(32))
        need += 6 * sum(1 for sb in range(32) for ch in range(n_ch) if alloc[ch][sb])
        need += 12 * sum((alloc[ch][sb] + 1) for sb in range(32) for ch in range(n_ch if sb < bound else 1) if alloc[ch][sb])
        if need <= budget:
            break
        density *= 0.8
    w = BitWriterMsb()
    for sb in range(32):
        for ch in range(n_ch if sb < bound else 1):
            w.put(alloc[ch][sb], 4)
    sf = [[0] * 32, [0] * 32]
    for sb in range(32):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                sf[ch][sb] = int(rng.integers(0, 64))
                w.put(sf[ch][sb], 6)
    raw = np.zeros((2, 32, 12), dtype=np.int64)
    for s in range(12):
        for sb in range(32):
            for ch in range(n_ch if sb < bound else 1):
                if alloc[ch][sb]:
                    bits = alloc[ch][sb] + 1
                    v = int(rng.integers(0, 1 << bits)) if rng.integers(8) else int(rng.choice([0, (1 << bits) - 1, 1 << (bits - 1)]))
                    raw[ch, sb, s] = v
                    w.put(v, bits)
    assert w.n == need
    body = w.bytes()
    word = st.mpa_word(version=version, layer=1, bitrate_idx=bitrate_idx, rate_idx=rate_idx, mode=mode, mode_ext=mode_ext, padding=pad, protected=protected)
    head = word.to_bytes(4, "big") + (bytes(rng.integers(0, 256, 2, dtype=np.uint8)) if protected else b"")
    frame = head + body + bytes(rng.integers(0, 256, total - len(head) - len(body), dtype=np.uint8))
    return frame, dict(alloc=alloc, sf=sf, raw=raw, bound=bound, n_ch=n_ch)


def gen_layer2_frame(rng, version="1", bitrate_idx=8, rate_idx=0, mode=0, mode_ext=0, protected=False, density=0.7):
    n_ch = 1 if mode == 3 else 2
    table = TABLES[layer2_table(version, bitrate_idx, rate_idx, n_ch)]
    sblimit = len(table)
    bound = min((mode_ext + 1) * 4 if mode == 1 else 32, sblimit)
    pad = int(rng.integers(2))
    total = st.mpa_frame_len(version, 2, bitrate_idx, rate_idx, pad)
    budget = (total - 4 - (2 if protected else 0)) * 8

    def code_bits(levels):
        return {3: 5, 5: 7, 9: 10}.get(levels) or 3 * (levels + 1).bit_length() - 3

    while True:
        alloc = [[0] * 32, [0] * 32]
        scfsi = [[0] * 32, [0] * 32]
        for sb in range(sblimit):
            nbal = table[sb][0]
            for ch in range(n_ch if sb < bound else 1):
                alloc[ch][sb] = int(rng.integers(1, 1 << nbal)) if rng.random() < density else 0
            if sb >= bound:
                alloc[1][sb] = alloc[0][sb]
            for ch in range(n_ch):
                scfsi[ch][sb] = int(rng.integers(4))
        need = sum(table[sb][0] * (n_ch if sb < bound else 1) for sb in range(sblimit))
        for sb in range(sblimit):
            for ch in range(n_ch):
                if alloc[ch][sb]:
                    need += 2 + 6 * {0: 3, 1: 2, 2: 1, 3: 2}[scfsi[ch][sb]]
            for ch in range(n_ch if sb < bound else 1):
                if alloc[ch][sb]:
                    need += 12 * code_bits(table[sb][1][alloc[ch][sb]])
        if need <= budget:
            break
        density *= 0.8
    w = BitWriterMsb()
    for sb in range(sblimit):
        for ch in range(n_ch if sb < bound else 1):
            w.put(alloc[ch][sb], table[sb][0])
    for sb in range(sblimit):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                w.put(scfsi[ch][sb], 2)
    sf = np.zeros((2, 3, 32), dtype=np.int64)
    for sb in range(sblimit):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                vals = [int(x) for x in rng.integers(0, 64, 3)]
                sel = scfsi[ch][sb]
                if sel == 0:
                    sent = vals
                elif sel == 1:
                    vals[1] = vals[0]
                    sent = [vals[0], vals[2]]
                elif sel == 2:
                    vals[1] = vals[2] = vals[0]
                    sent = [vals[0]]
                else:
                    vals[2] = vals[1]
                    sent = [vals[0], vals[1]]
                for v in sent:
                    w.put(v, 6)
                sf[ch, :, sb] = vals
    raw = np.zeros((2, 32, 36), dtype=np.int64)
    for gr in range(12):
        for sb in range(sblimit):
            for ch in range(n_ch if sb < bound else 1):
                if alloc[ch][sb]:
                    levels = table[sb][1][alloc[ch][sb]]
                    if levels in (3, 5, 9):
                        trip = [int(x) for x in rng.integers(0, levels, 3)]
                        w.put(trip[0] + levels * trip[1] + levels * levels * trip[2], code_bits(levels))
                    else:
                        bits = (levels + 1).bit_length() - 1
                        trip = [int(x) for x in rng.integers(0, levels, 3)]  # the all-ones code is forbidden (it would be a sync pattern)
                        for v in trip:
                            w.put(v, bits)
                    raw[ch, sb, 3 * gr:3 * gr + 3] = trip
    assert w.n == need, (w.n, need)
    body = w.bytes()
    word = st.mpa_word(version=version, layer=2, bitrate_idx=bitrate_idx, rate_idx=rate_idx, mode=mode, mode_ext=mode_ext, padding=pad, protected=protected)
    head = word.to_bytes(4, "big") + (bytes(rng.integers(0, 256, 2, dtype=np.uint8)) if protected else b"")
    frame = head + body + bytes(rng.integers(0, 256, total - len(head) - len(body), dtype=np.uint8))
    return frame, dict(alloc=alloc, scfsi=scfsi, sf=sf, raw=raw, bound=bound, n_ch=n_ch, table=table, sblimit=sblimit)
