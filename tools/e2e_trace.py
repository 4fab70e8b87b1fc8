#!/usr/bin/env python
"""Timeline of one end-to-end MP3 step (GPU box): runs symgpu_mp3_synth_host on pinned buffers with SYMGPU_E2E_TRACE=1, which makes
the library print, per slice, when its H2D copy, kernel and D2H copy started and ended (CUDA events on their streams)."""
import os
import sys
import time

import numpy as np
import torch

os.environ["SYMGPU_E2E_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symphonia_b200 as sb  # noqa: E402
from symphonia_b200 import workloads  # noqa: E402

S, F = 64, 128
N = S * F
units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1)
u_pin = torch.from_numpy(units.view(np.uint8).reshape(-1)).pin_memory()
s_pin = torch.from_numpy(spectra).pin_memory()
p_pin = torch.empty((N, 2, 1152), dtype=torch.float32).pin_memory()
u_np = u_pin.numpy().view(sb._native.MP3_GC_DTYPE).reshape(N, 2, 2)
eng = sb.Engine(0)
eng.mp3_streams_alloc(S)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t = time.perf_counter()
    eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy())
    print(f"call {rep}: {1e3 * (time.perf_counter() - t):.3f} ms wall", file=sys.stderr)
eng.close()
