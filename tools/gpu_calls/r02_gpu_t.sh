#!/bin/bash
# Vorbis kernel: Z layout (one warp per packet-channel) against the 64-thread array layout
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02t_build.log 2>&1 || { tail -20 gpurun_out/r02t_build.log; exit 1; }
for k in ${VORBIS_VARIANTS:-z pair}; do
  echo "== SYMGPU_VORBIS_KERNEL=$k"
  SYMGPU_VORBIS_KERNEL=$k timeout 900 python -m pytest tests/test_aac_vorbis_parity_gpu.py tests/test_zz_ogg_vorbis_to_pcm.py -m gpu -x -q -k "vorbis or Vorbis" 2>&1 | tail -3
  SYMGPU_VORBIS_KERNEL=$k timeout 300 python bench_codecs.py --codec vorbis --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-500
  SYMGPU_VORBIS_KERNEL=$k timeout 300 python bench_codecs.py --codec mixed --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-500
done
# full captures of the two Z kernels
SYMGPU_AAC_KERNEL=z timeout 300 ncu --set full --clock-control none --import-source on -k regex:aac_synth -c 1 -s 4 -o gpurun_out/r02t_prof_aac_z -f python bench_codecs.py --codec aac --tns 0 --steps 3 --warmup 3 > gpurun_out/r02t_prof_aac_z.log 2>&1
SYMGPU_VORBIS_KERNEL=z timeout 300 ncu --set full --clock-control none --import-source on -k regex:vorbis_synth -c 1 -s 3 -o gpurun_out/r02t_prof_vorbis_z -f python bench_codecs.py --codec vorbis --steps 3 --warmup 3 > gpurun_out/r02t_prof_vorbis_z.log 2>&1
tail -1 gpurun_out/r02t_prof_aac_z.log; tail -1 gpurun_out/r02t_prof_vorbis_z.log
