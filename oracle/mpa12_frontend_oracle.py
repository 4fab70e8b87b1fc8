"""Layer I / II sample decoder oracle (Layer I / II path): Layer1::decode and Layer2::decode up to their
call of the polyphase synthesis, restated in the reference's sequence with numpy float32 arithmetic (every product and
sum is one IEEE f32 operation, as in the Rust).  TEST INFRASTRUCTURE ONLY.

  symphonia-bundle-mp3/src/layer1/mod.rs:19-176, layer2/mod.rs:19-369, layer12.rs:9-75, decoder.rs:84-131

Constants come from tests/golden/mpa12_constants.json -- the reference's own decimal literals as f32 bit patterns and
its allocation tables, written by tools/make_mpa12_golden.py -- so nothing numeric is shared with the C++ (which uses
closed forms).  Pinned by hand-computed samples in tests/test_mpa12_frontend.py."""
import json
import os

import numpy as np

from oracle import packetizer_oracle as po
from oracle.mp3_frontend_oracle import BitsLtr, DecodeError

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(os.path.dirname(_HERE), "tests", "golden", "mpa12_constants.json")) as _f:
    G = json.load(_f)
f32 = np.float32
SCALE = np.array(G["scalefactors"], dtype=np.uint32).view(np.float32)
CLASSES = [dict(q, c=np.array([q["c"]], dtype=np.uint32).view(np.float32)[0], d=np.array([q["d"]], dtype=np.uint32).view(np.float32)[0])
           for q in G["quant_classes"]]


def _sign_extend(v, bits):
    v &= (1 << bits) - 1
    return v - (1 << bits) if v & (1 << (bits - 1)) else v


def _factor(nb):
    """layer1/mod.rs:19-48."""
    a, b = 1 << nb, 1 << (nb - 1)
    return (f32(a) / f32(a - 1)) * (f32(1.0) / f32(b))


def decode_layer1(bs, h):
    """layer1/mod.rs:73-176 -> samples [2][384]."""
    n_ch = h["n_channels"]
    bound = h["bound"] if h["mode"] == "joint" else 32
    alloc = [[0] * 32, [0] * 32]
    sf = [[f32(0)] * 32, [f32(0)] * 32]
    for sb in range(bound):
        for ch in range(n_ch):
            bits = bs.read(4)
            if bits > 14:
                raise DecodeError("bit allocation")
            alloc[ch][sb] = bits + 1 if bits else 0
    for sb in range(bound, 32):
        bits = bs.read(4)
        if bits > 14:
            raise DecodeError("bit allocation")
        alloc[0][sb] = alloc[1][sb] = bits + 1 if bits else 0
    for sb in range(32):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                sf[ch][sb] = SCALE[bs.read(6)]
    out = np.zeros((2, 384), dtype=np.float32)
    for s in range(12):
        for sb in range(bound):
            for ch in range(n_ch):
                bits = alloc[ch][sb]
                if bits:
                    raw = bs.read(bits)
                    a = _sign_extend(raw ^ (1 << (bits - 1)), bits)
                    out[ch, 12 * sb + s] = sf[ch][sb] * (_factor(bits) * f32(a + 1))
        for sb in range(bound, 32):
            bits = alloc[0][sb]
            if bits:
                raw = bs.read(bits)
                a = _sign_extend(raw ^ (1 << (bits - 1)), bits)
                sample = _factor(bits) * f32(a + 1)
                for ch in range(n_ch):
                    out[ch, 12 * sb + s] = sf[ch][sb] * sample
    return out


def _sb_info(h):
    """layer2/mod.rs:136-166."""
    if h["version"] == "1":
        per = h["bitrate"] // h["n_channels"]
        if per <= 48000:
            idx = 3 if h["sample_rate"] == 32000 else 2
        elif per <= 80000:
            idx = 0
        else:
            idx = int(h["sample_rate"] != 48000)
    else:
        idx = 4
    return G["sb_info"][idx]


def _dequantize(bs, q):
    """layer2/mod.rs:169-219."""
    raw = [0, 0, 0]
    if q["grouping"]:
        c = bs.read(q["bits"])
        for k in range(3):
            raw[k] = c % q["nlevels"]
            c //= q["nlevels"]
        bits = (q["nlevels"] - 1).bit_length()  # next_power_of_two().trailing_zeros()
        if 1 << bits < q["nlevels"]:
            bits += 1
    else:
        bits = q["bits"]
        for k in range(3):
            raw[k] = bs.read(bits)
    divisor = f32(1 << (bits - 1))
    out = []
    for k in range(3):
        a = _sign_extend(raw[k] ^ (1 << (bits - 1)), bits)
        out.append(q["c"] * (f32(a) / divisor + q["d"]))
    return out


def decode_layer2(bs, h):
    """layer2/mod.rs:230-369 -> samples [2][1152]."""
    n_ch = h["n_channels"]
    info = _sb_info(h)
    sblimit = info["sblimit"]
    bound = min(h["bound"] if h["mode"] == "joint" else 32, sblimit)
    quant = [G["sb_quant_info"][b] for b in info["bands"]]
    alloc = [[0] * 32, [0] * 32]
    scfsi = [[0] * 32, [0] * 32]
    sf = [[[0] * 32 for _ in range(3)] for _ in range(2)]
    for sb in range(bound):
        for ch in range(n_ch):
            alloc[ch][sb] = bs.read(quant[sb]["nbal"])
    for sb in range(bound, sblimit):
        alloc[0][sb] = alloc[1][sb] = bs.read(quant[sb]["nbal"])
    for sb in range(sblimit):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                scfsi[ch][sb] = bs.read(2)
    for sb in range(sblimit):
        for ch in range(n_ch):
            if alloc[ch][sb]:
                idx = [bs.read(6)] * 3
                sel = scfsi[ch][sb]
                if sel == 0:
                    idx[1] = bs.read(6)
                    idx[2] = bs.read(6)
                elif sel == 1:
                    idx[2] = bs.read(6)
                elif sel == 3:
                    idx[1] = bs.read(6)
                    idx[2] = idx[1]
                for p in range(3):
                    sf[ch][p][sb] = idx[p]
    out = np.zeros((2, 1152), dtype=np.float32)
    for gr in range(12):
        for sb in range(bound):
            for ch in range(n_ch):
                if alloc[ch][sb]:
                    t = _dequantize(bs, CLASSES[quant[sb]["classes"][alloc[ch][sb]]])
                    s = SCALE[sf[ch][gr // 4][sb]]
                    for k in range(3):
                        out[ch, 36 * sb + 3 * gr + k] = s * t[k]
        for sb in range(bound, sblimit):
            if alloc[0][sb]:
                t = _dequantize(bs, CLASSES[quant[sb]["classes"][alloc[0][sb]]])
                for ch in range(n_ch):
                    s = SCALE[sf[ch][gr // 4][sb]]
                    for k in range(3):
                        out[ch, 36 * sb + 3 * gr + k] = s * t[k]
    return out


class Mpa12Frontend:
    """MpaDecoder::decode_inner for a Layer I or Layer II stream (decoder.rs:84-131)."""

    def __init__(self, layer):
        self.layer, self.spec = layer, None

    def decode(self, packet):
        r = po.Reader(packet)
        try:
            h = po.mpa_parse_header(po.mpa_sync_frame(r))
        except po.ReaderError as e:
            raise DecodeError(str(e))
        if h["frame_size"] != r.available():
            raise DecodeError("packet length")
        spec = (h["sample_rate"], h["n_channels"])
        if self.spec is None:
            self.spec = spec
        elif self.spec != spec:
            raise DecodeError("signal spec")
        if h["layer"] != self.layer:
            raise DecodeError("layer")
        body = packet[r.pos:]
        if h["crc"]:
            if len(body) < 2:
                raise DecodeError("crc")
            body = body[2:]
        bs = BitsLtr(body)
        return h, (decode_layer1(bs, h) if self.layer == 1 else decode_layer2(bs, h))
