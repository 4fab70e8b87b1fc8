#!/bin/bash
# AAC Z kernel: 16 warps x 2 CTAs per SM against 10 x 3 and 8 x 4 (same box)
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02zk_build.log 2>&1 || { tail -20 gpurun_out/r02zk_build.log; exit 1; }
for k in 15 9 7 15; do
  echo "== SYMGPU_AAC_Z_FRAMES=$k"
  SYMGPU_AAC_Z_FRAMES=$k timeout 300 python bench_codecs.py --codec aac --steps 40 --warmup 5 --tns 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('aac no tns us', round(1e3*d['kernel_ms'],2))"
  SYMGPU_AAC_Z_FRAMES=$k timeout 300 python bench_codecs.py --codec aac --steps 40 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('aac tns20 us', round(1e3*d['kernel_ms'],2))"
  SYMGPU_AAC_Z_FRAMES=$k timeout 300 python bench_codecs.py --codec mixed --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mixed step ms', round(d['step_ms'],4))"
done
for k in 9 7; do SYMGPU_AAC_Z_FRAMES=$k timeout 600 python -m pytest tests/test_aac_vorbis_parity_gpu.py -m gpu -x -q -k "aac_mixed or aac_chunk or aac_state or aac_heavy or aac_mono or aac_full" 2>&1 | tail -1; done
