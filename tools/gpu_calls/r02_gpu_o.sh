#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02o}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 600 python -m pytest tests/test_mp3_parity_gpu.py tests/test_abi_errors_gpu.py tests/test_output_stage_gpu.py -m gpu -q 2>&1 | tail -4
for ah in 1 2 3; do for sl in 8 10; do SYMGPU_H2D_AHEAD=$ah SYMGPU_SLICES=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ahead $ah slices $sl e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3), 'value', round(d['e2e']['value']), 's16', round(d['e2e_s16']['ms_per_step'],3), 'compact', round(d['e2e_compact']['ms_per_step'],3))"; done; done
