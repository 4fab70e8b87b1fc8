#!/usr/bin/env python3
"""Dev-time generator (needs /root/reference; never run by tests, smoke or bench): the constants of the Layer I / II
sample decoders as the reference spells them -- scale factors (ISO 11172-3 Table 3-B.1, layer12.rs:9-75), quantisation
classes (Table 3-B.4, layer2/mod.rs:45-63), allocation tables (Tables 3-B.2a-d and 13818-3 Table B.1,
layer2/mod.rs:66-118) -- as f32 bit patterns / integers in tests/golden/mpa12_constants.json, for the tests to hold the
product's closed-form tables against."""
import json
import os
import re
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/symphonia-bundle-mp3/src"


def bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def main():
    l12 = open(os.path.join(REF, "layer12.rs")).read()
    l2 = open(os.path.join(REF, "layer2", "mod.rs")).read()
    scale = [bits(v) for v in re.findall(r"^\s+(\d\.\d{14}),", l12, re.M)]
    assert len(scale) == 64
    qc = re.findall(r"QuantClass \{ c: ([\d.]+), d: ([\d.]+), grouping: (\w+), bits: (\d+), nlevels: (\d+) \}", l2)
    assert len(qc) == 17
    classes = [dict(c=bits(c), d=bits(d), grouping=g == "true", bits=int(b), nlevels=int(n)) for c, d, g, b, n in qc]
    sbq = [dict(nbal=int(n), classes=[int(x) for x in cl.split(",")]) for n, cl in re.findall(r"SbQuantInfo \{ nbal: (\d+), classes: \[([\d, ]+)\] \}", l2)]
    assert len(sbq) == 8
    sbi = [dict(sblimit=int(n), bands=[int(x) for x in re.findall(r"\d+", b)]) for n, b in re.findall(r"sblimit: (\d+),\s+bands: \[([\d,\s]+)\]", l2)]
    assert len(sbi) == 5 and all(len(t["bands"]) == 32 for t in sbi)
    with open(os.path.join(ROOT, "tests", "golden", "mpa12_constants.json"), "w") as f:
        json.dump(dict(scalefactors=scale, quant_classes=classes, sb_quant_info=sbq, sb_info=sbi), f, separators=(",", ":"))
    print("wrote", len(scale), "scale factors,", len(classes), "classes,", len(sbq), "+", len(sbi), "tables")


if __name__ == "__main__":
    main()
