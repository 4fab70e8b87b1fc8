#!/bin/bash
# small host batches on the caller's pinned memory (one launch) against staged copies: parity + per-packet time
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02s_build.log 2>&1 || { tail -20 gpurun_out/r02s_build.log; exit 1; }
timeout 600 python -m pytest tests/test_mp3_parity_gpu.py tests/test_cpp_host.py tests/test_abi_errors_gpu.py -m gpu -x -q 2>&1 | tail -3
for z in default s; do
  if [ $z = default ]; then unset SYMGPU_ZERO_COPY; else export SYMGPU_ZERO_COPY=$z; fi
  echo "== SYMGPU_ZERO_COPY=$z"
  timeout 300 python -m pytest tests/test_cpp_host.py -m gpu -q -s -k packet_by_packet 2>&1 | grep -E "us_per_packet|passed|failed"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['configs']['plumbing']; print('plumbing us_per_packet', round(p['us_per_packet'],2), 'audio-s/s', round(p['value']), 'mixed e2e', round(d['configs']['mixed']['e2e']['value']))"
done
