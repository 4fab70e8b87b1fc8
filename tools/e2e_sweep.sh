#!/bin/bash
# PCIe ceiling + end-to-end time of the host entry point for several pipeline depths (GPU box).
cd "$(dirname "$0")/.."
python tools/pcie_probe.py
for s in "$@"; do
  SYMGPU_SLICES=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('slices', $s, 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'e2e_s16_ms', round(d['e2e_s16']['ms_per_step'],3))"
done
