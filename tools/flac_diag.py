#!/usr/bin/env python
"""Where the FLAC kernel differs from the oracle (GPU box)."""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symphonia_b200 as sb  # noqa: E402
from symphonia_b200 import workloads  # noqa: E402
from tests import _oracle, test_oracle_kat_flac as kat  # noqa: E402

orc = _oracle.load()
eng = sb.Engine(0)
for bps, ch, block in ((16, 2, 512), (16, 1, 300)):
    frames, subs, samples, expect = workloads.flac_batch(40, block, seed=900 + bps + block, bps=bps, channels=ch, return_pcm=True)
    rc, want = kat._restore(orc, frames, subs, samples)
    got = eng.flac_restore_host(frames, subs, samples.copy())
    agg = collections.Counter()
    first = None
    for k, sf in enumerate(subs):
        a, n = int(sf["offset"]), int(sf["n"])
        bad = np.nonzero(got[a:a + n] != want[a:a + n])[0]
        asg = int(frames[k // ch]["assignment"])
        key = (int(sf["type"]), int(sf["order"]), int(sf["wasted"]) > 0, asg, k % ch, k % 8)
        agg[(key, len(bad) > 0)] += 1
        if len(bad) and first is None:
            first = (k, key, int(bad[0]), len(bad), n, got[a + bad[0]:a + bad[0] + 4].tolist(), want[a + bad[0]:a + bad[0] + 4].tolist(),
                     samples[a + bad[0]:a + bad[0] + 4].tolist())
    print("config", bps, ch, block, "bad subframes", sum(v for (k, b), v in agg.items() if b), "of", len(subs))
    print(" first:", first)
    print(" bad keys (type, order, wasted, assignment, ch, lane):", sorted(k for (k, b) in agg if b)[:24])
    print(" good keys:", sorted(k for (k, b) in agg if not b)[:24])
