#!/bin/bash
# same-box A/B of three small changes: TNS history as a ring of registers, Vorbis floor posts in level order (vs. without each)
cd "$(dirname "$0")/.."
for v in "" "-DSYMGPU_TNS_NO_RING" ""; do
  make -C symphonia_b200/csrc -B EXTRA="$v" > gpurun_out/r02ab_build.log 2>&1 || { tail -5 gpurun_out/r02ab_build.log; exit 1; }
  echo "== EXTRA='$v'"
  for r in 1 2; do
    timeout 300 python bench_codecs.py --codec aac --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('aac tns20 us', round(1e3*d['kernel_ms'],2))"
    timeout 300 python bench_codecs.py --codec vorbis --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('vorbis us', round(1e3*d['kernel_ms'],2))"
  done
done
timeout 600 python -m pytest tests/test_aac_vorbis_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
