#!/bin/bash
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02v_build.log 2>&1 || { tail -20 gpurun_out/r02v_build.log; exit 1; }
for cs in 2 1; do
  echo "== SYMGPU_COPY_STREAMS=$cs"
  SYMGPU_COPY_STREAMS=$cs timeout 300 python tools/e2e_trace.py 3 2>&1 | tail -2
  SYMGPU_COPY_STREAMS=$cs SYMGPU_E2E_TRACE=0 timeout 300 python - <<'PY'
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import symphonia_b200 as sb
from symphonia_b200 import workloads
S, F = 64, 128; N = S * F
units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1)
u_pin = torch.from_numpy(units.view(np.uint8).reshape(-1)).pin_memory(); s_pin = torch.from_numpy(spectra).pin_memory()
p_pin = torch.empty((N, 2, 1152), dtype=torch.float32).pin_memory()
u_np = u_pin.numpy().view(sb._native.MP3_GC_DTYPE).reshape(N, 2, 2)
for ahead in (2, 3, 4, 8):
    os.environ["SYMGPU_H2D_AHEAD"] = str(ahead)
    eng = sb.Engine(0); eng.mp3_streams_alloc(S)
    for _ in range(5): eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy())
    ts = []
    for _ in range(30):
        t = time.perf_counter(); eng.mp3_synth_host(u_np, s_pin.numpy(), runs, out=p_pin.numpy()); ts.append(time.perf_counter() - t)
    print("ahead", ahead, "e2e ms mean %.3f median %.3f min %.3f" % (1e3 * np.mean(ts), 1e3 * np.median(ts), 1e3 * np.min(ts)))
    eng.close()
PY
done
timeout 600 python -m pytest tests/test_mp3_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
