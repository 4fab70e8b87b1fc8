#!/bin/bash
# AAC / Vorbis kernels: parity of both codecs (incl. the selectable older shapes) + kernel times (+ mixed)
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02u_build.log 2>&1 || { tail -20 gpurun_out/r02u_build.log; exit 1; }
timeout 900 python -m pytest tests/test_aac_vorbis_parity_gpu.py tests/test_zz_ogg_vorbis_to_pcm.py tests/test_zz_adts_aac_to_pcm.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-330
timeout 300 python bench_codecs.py --codec aac --steps 30 --warmup 5 --tns 0 2>&1 | tail -1 | cut -c1-330
timeout 300 python bench_codecs.py --codec vorbis --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-330
timeout 300 python bench_codecs.py --codec mixed --steps 20 --warmup 5 2>&1 | tail -1 | cut -c200-420
