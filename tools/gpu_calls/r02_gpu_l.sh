#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 900 python -m pytest tests/test_mp3_parity_gpu.py tests/test_aac_vorbis_parity_gpu.py tests/test_zz_ogg_vorbis_to_pcm.py -m gpu -q -x 2>&1 | tail -12 | tee $out/${tag}_pytest.txt
for zc in 0 1 2; do for sl in 8 16; do SYMGPU_ZERO_COPY=$zc SYMGPU_SLICES=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zero_copy $zc slices $sl e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3), 'value', round(d['e2e']['value']), 's16', round(d['e2e_s16']['ms_per_step'],3), 'parity', d['parity']['ranks_bit_exact_vs_oracle'])"; done; done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_n1.json 2>$out/${tag}_bench_n1.err; tail -3 $out/${tag}_bench_n1.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_n1.json"))
print("mp3 value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],3))
for k,c in d.get("configs",{}).items():
    print(k, "value", round(c["value"]), "kernel_ms", c.get("kernel_ms"), "frac", round(c.get("roofline",{}).get("frac",0),4), "e2e", round(c["e2e"]["value"]), "e2e_ms", round(c["e2e"]["ms_per_step"],3), c.get("us_per_packet"))
PY
