#!/bin/bash
# same-box A/B: AAC pre-twiddle iterations in flight (2 / 4 / 8), then the Vorbis + many-files GPU tests
cd "$(dirname "$0")/../.."
for v in 2 8 4 2; do
  make -C symphonia_b200/csrc -B EXTRA="-DSYMGPU_AAC_PRE_UNROLL=$v" > gpurun_out/r02ab_build.log 2>&1 || { tail -5 gpurun_out/r02ab_build.log; exit 1; }
  echo "== SYMGPU_AAC_PRE_UNROLL=$v"
  for r in 1 2; do
    timeout 300 python bench_codecs.py --codec aac --steps 40 --warmup 5 --tns 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('aac no tns us', round(1e3*d['kernel_ms'],2))"
  done
done
timeout 600 python -m pytest tests/test_aac_vorbis_parity_gpu.py tests/test_zz_ogg_vorbis_to_pcm.py tests/test_zz_many_files.py -m gpu -x -q 2>&1 | tail -3
