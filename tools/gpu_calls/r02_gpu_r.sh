#!/bin/bash
# AAC kernel: one warp per frame (SYMGPU_AAC_KERNEL=warp) against two warps per frame.
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02r_build.log 2>&1 || { tail -20 gpurun_out/r02r_build.log; exit 1; }
for k in ${AAC_VARIANTS:-pair warp z}; do
  echo "== SYMGPU_AAC_KERNEL=$k"
  SYMGPU_AAC_KERNEL=$k timeout 600 python -m pytest tests/test_aac_vorbis_parity_gpu.py -m gpu -x -q -k "aac or Aac or AAC" 2>&1 | tail -3
  SYMGPU_AAC_KERNEL=$k timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 2>&1 | tail -2
  SYMGPU_AAC_KERNEL=$k timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 --tns 0 2>&1 | tail -1
  SYMGPU_AAC_KERNEL=$k timeout 300 python bench_codecs.py --codec mixed --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-600
done
