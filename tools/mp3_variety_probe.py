#!/usr/bin/env python
"""How much of the MP3 kernel time is the variety of granule jobs (block switching, joint stereo)?
Times the device-resident kernel on the bench shape for workloads with the variety switched off."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symphonia_b200 as sb  # noqa: E402
from symphonia_b200 import workloads  # noqa: E402

dev = torch.device("cuda", 0)
eng = sb.Engine(0)
S, F = 64, 128
eng.mp3_streams_alloc(S)
ext = torch.cuda.ExternalStream(eng.cuda_stream)
res = {}
for name, kw in (("default", {}), ("long_blocks_only", {"block_switching": False}), ("no_joint_stereo", {"joint": False}),
                 ("uniform", {"block_switching": False, "joint": False})):
    units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1, **kw)
    sets = [(torch.from_numpy(units.view(np.uint8).reshape(-1).copy()).to(dev), torch.from_numpy(spectra).to(dev),
             torch.empty((S * F, 2, 1152), dtype=torch.float32, device=dev)) for _ in range(4)]
    for i in range(5):
        eng.mp3_synth_dev(sets[i % 4][0], sets[i % 4][1], runs, sets[i % 4][2])
    eng.sync()
    with torch.cuda.stream(ext):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(100):
            eng.mp3_synth_dev(sets[i % 4][0], sets[i % 4][1], runs, sets[i % 4][2])
        b.record()
    eng.sync()
    torch.cuda.synchronize()
    res[name] = round(a.elapsed_time(b) * 10.0, 1)  # us per launch
    del sets
print(json.dumps(res))
