#!/bin/bash
# End-to-end time of the host entry points for several pipeline depths, next to the box's PCIe ceiling (GPU box).
cd "$(dirname "$0")/.."
python tools/pcie_probe.py
for s in "$@"; do
  SYMGPU_SLICES=$s python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('slices', $s, 'e2e_ms', round(d['e2e']['ms_per_step'],3), 'e2e_s16_ms', round(d['e2e_s16']['ms_per_step'],3), 'e2e_compact_ms', round(d['e2e_compact']['ms_per_step'],3))"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/e2e_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/e2e_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    agg[r[4].split('(')[0][-36:]].append(float(r[-1])/1000)
for k,v in agg.items(): print(k, len(v), 'avg us', round(sum(v)/len(v),1), 'max', round(max(v),1))
PY
