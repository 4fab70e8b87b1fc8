#!/bin/bash
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02w_build.log 2>&1 || { tail -20 gpurun_out/r02w_build.log; exit 1; }
timeout 600 python - <<'PY'
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import symphonia_b200 as sb
from symphonia_b200 import workloads
S, F = 64, 128; N = S * F
units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1)
u_pin = torch.from_numpy(units.view(np.uint8).reshape(-1)).pin_memory(); s_pin = torch.from_numpy(spectra).pin_memory()
p_pin = torch.empty((N, 2, 1152), dtype=torch.float32).pin_memory()
u_np = u_pin.numpy().view(sb._native.MP3_GC_DTYPE).reshape(N, 2, 2)
res = {}
for rnd in range(3):
    for slices in (6, 8, 12):
        for ahead in (2, 4, 6):
            os.environ["SYMGPU_H2D_AHEAD"] = str(ahead); os.environ["SYMGPU_SLICES"] = str(slices)
            eng = sb.Engine(0); eng.mp3_streams_alloc(S)
            for _ in range(4): eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy())
            ts = []
            for _ in range(25):
                t = time.perf_counter(); eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy()); ts.append(time.perf_counter() - t)
            res.setdefault((slices, ahead), []).append(1e3 * np.median(ts))
            eng.close()
for k, v in sorted(res.items()): print("slices %2d ahead %d  median ms per round: %s" % (k[0], k[1], " ".join("%.3f" % x for x in v)))
PY
