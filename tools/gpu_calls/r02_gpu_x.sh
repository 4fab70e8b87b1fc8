#!/bin/bash
# AAC Z kernel: TNS on the frame's own warp against the three-kernel pre-pass
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02x_build.log 2>&1 || { tail -20 gpurun_out/r02x_build.log; exit 1; }
for t in ${TNS_MODES:-frames sorted}; do
  echo "== SYMGPU_AAC_TNS=$t"
  SYMGPU_AAC_TNS=$t timeout 900 python -m pytest tests/test_aac_vorbis_parity_gpu.py tests/test_zz_adts_aac_to_pcm.py -m gpu -x -q -k "aac or Aac or adts" 2>&1 | tail -2
  SYMGPU_AAC_TNS=$t timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-330
  SYMGPU_AAC_TNS=$t timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 --tns 0.05 2>&1 | tail -1 | cut -c1-330
  SYMGPU_AAC_TNS=$t timeout 300 python bench_codecs.py --codec mixed --steps 20 --warmup 5 2>&1 | tail -1 | cut -c200-420
done
