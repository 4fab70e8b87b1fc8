#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
for ah in 1 2 3 4; do for sl in 8 12; do SYMGPU_H2D_AHEAD=$ah SYMGPU_SLICES=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ahead $ah slices $sl e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3), 's16', round(d['e2e_s16']['ms_per_step'],3), 'compact', round(d['e2e_compact']['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],4))"; done; done
timeout 600 python tools/mp3_variant_bench.py v1 v1p auto 12:33 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $out/${tag}_pytest_gpu.txt
