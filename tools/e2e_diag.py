#!/usr/bin/env python
"""Where the end-to-end time of the host entry points goes (GPU box): per-variant wall time for several
pipeline depths, next to raw pinned-copy times of the very buffers the calls use."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symphonia_b200 as sb  # noqa: E402
from symphonia_b200 import workloads  # noqa: E402

S, F = 64, 128
N = S * F
units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1)
dev = torch.device("cuda", 0)
u_pin = torch.from_numpy(units.view(np.uint8).reshape(-1)).pin_memory()
s_pin = torch.from_numpy(spectra).pin_memory()
g_pin = torch.from_numpy(workloads.mp3_quantize(spectra)).pin_memory()
p_pin = torch.empty((N, 2, 1152), dtype=torch.float32).pin_memory()
q_pin = torch.empty((N * 1152, 2), dtype=torch.int16).pin_memory()
u_np = u_pin.numpy().view(sb._native.MP3_GC_DTYPE).reshape(N, 2, 2)
res = {"pinned": [t.is_pinned() for t in (u_pin, s_pin, g_pin, p_pin, q_pin)]}


def timeit(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return round(1e3 * (time.perf_counter() - t) / reps, 3)


d_s = torch.empty_like(s_pin, device=dev)
d_g = torch.empty_like(g_pin, device=dev)
d_p = torch.empty_like(p_pin, device=dev)
d_q = torch.empty_like(q_pin, device=dev)
res["raw_h2d_f32_ms"] = timeit(lambda: d_s.copy_(s_pin, non_blocking=True))
res["raw_h2d_i16_ms"] = timeit(lambda: d_g.copy_(g_pin, non_blocking=True))
res["raw_d2h_f32_ms"] = timeit(lambda: p_pin.copy_(d_p, non_blocking=True))
res["raw_d2h_i16_ms"] = timeit(lambda: q_pin.copy_(d_q, non_blocking=True))
for slices in (6, 8, 10):
    os.environ["SYMGPU_SLICES"] = str(slices)
    eng = sb.Engine(0)
    eng.mp3_streams_alloc(S)
    FMT = sb._native.FMT_S16
    res[f"s{slices}_f32_f32_ms"] = timeit(lambda: eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy()))
    res[f"s{slices}_f32_i16_ms"] = timeit(lambda: eng.mp3_synth_host_packed(u_np, s_pin.numpy(), runs, FMT, out=q_pin.numpy()))
    res[f"s{slices}_i16_f32_ms"] = timeit(lambda: eng.mp3_synth_host_quantized(u_np, g_pin.numpy(), runs, None, out=p_pin.numpy()))
    res[f"s{slices}_i16_i16_ms"] = timeit(lambda: eng.mp3_synth_host_quantized(u_np, g_pin.numpy(), runs, FMT, out=q_pin.numpy()))
    eng.close()
print(json.dumps(res))
