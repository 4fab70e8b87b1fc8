#!/usr/bin/env python
"""Dev-time (needs /root/reference): copies the literal known-answer INPUT vectors of the reference's
own unit tests into tests/golden/ as small JSON fixtures, so the oracle can be pinned against them
on machines where /root/reference does not exist (the GPU box).  Only test-vector literals are
extracted; expected values are recomputed from the analytical definitions, as the reference does."""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

src = open(os.path.join(REF, "symphonia-core/src/dsp/fft/mod.rs")).read()
m = re.search(r"const TEST_VECTOR: \[Complex<f32>; 64\] = \[(.*?)\];", src, re.S)
pairs = re.findall(r"re:\s*(-?[0-9.]+),\s*im:\s*(-?[0-9.]+)", m.group(1))
assert len(pairs) == 64
json.dump({"source": "symphonia-core/src/dsp/fft/mod.rs:88-153 (TEST_VECTOR)",
           "re": [float(a) for a, _ in pairs], "im": [float(b) for _, b in pairs]},
          open(os.path.join(OUT, "fft64_test_vector.json"), "w"), indent=0)

src = open(os.path.join(REF, "symphonia-core/src/dsp/mdct.rs")).read()
m = re.search(r"const TEST_VECTOR: \[f32; 32\] = \[(.*?)\];", src, re.S)
vals = [float(v) for v in re.findall(r"-?[0-9]+\.[0-9]+", m.group(1))]
assert len(vals) == 32
json.dump({"source": "symphonia-core/src/dsp/mdct.rs:180-185 (TEST_VECTOR), scale sqrt(2/64)", "x": vals},
          open(os.path.join(OUT, "imdct32_test_vector.json"), "w"), indent=0)
print("wrote fixtures to", OUT)
