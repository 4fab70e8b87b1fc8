#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02k}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 900 python -m pytest tests/test_mp3_parity_gpu.py tests/test_aac_vorbis_parity_gpu.py -m gpu -q -x 2>&1 | tail -12 | tee $out/${tag}_pytest.txt
for zc in 1 0; do for k in auto v2; do SYMGPU_ZERO_COPY=$zc SYMGPU_MP3_KERNEL=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zero_copy $zc kernel $k e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3), 'value', round(d['e2e']['value']), 's16', round(d['e2e_s16']['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'parity', d['parity']['ranks_bit_exact_vs_oracle'])"; done; done
