#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 900 python -m pytest tests/test_mp3_parity_gpu.py tests/test_aac_vorbis_parity_gpu.py tests/test_zz_file_to_pcm.py tests/test_output_stage_gpu.py -m gpu -q 2>&1 | tail -12 | tee $out/${tag}_pytest.txt
timeout 600 python tools/mp3_variant_bench.py v1p auto 12:81 12:17 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
timeout 300 python bench_codecs.py --codec vorbis --steps 30 2>&1 | cut -c1-330
timeout 300 python bench_codecs.py --codec mp3-short --steps 30 2>&1 | cut -c1-330
