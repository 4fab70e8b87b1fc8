#!/bin/bash
# Round 2, GPU call B: A/B of the v2 kernel variants; re-run of the two tests that failed in call A.
cd "$(dirname "$0")/../.."
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 600 python tools/mp3_variant_bench.py 2>&1 | tee $out/${tag}_variants.txt
timeout 600 python -m pytest tests/test_flac_parity_gpu.py tests/test_zz_many_files.py tests/test_flac_frontend.py -m gpu -q 2>&1 | tail -15 | tee $out/${tag}_pytest_fixed.txt
