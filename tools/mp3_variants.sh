#!/bin/bash
# Runs bench.py for several (T, NW) builds of the MP3 kernel on the GPU box and restores the default build.
# usage: tools/mp3_variants.sh "T NW" "T NW" ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  set -- $v
  make -B -C symphonia_b200/csrc EXTRA="-DSYMGPU_MP3_T=$1 -DSYMGPU_MP3_NW=$2" > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  regs=$(grep -A6 "mp3_synth_kernel" symphonia_b200/csrc/build.log | grep -E "Used [0-9]+ registers" | head -1 | sed 's/.*Used \([0-9]*\) registers.*/\1/')
  timeout 120 python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('T=$1 NW=$2 regs=$regs kernel_us=%.1f e2e_ms=%.2f' % (1e3*d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))"
done
make -B -C symphonia_b200/csrc > /dev/null 2>&1
