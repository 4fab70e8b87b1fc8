#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 600 python tools/mp3_variant_bench.py v1 v1p 12:33 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
timeout 600 python -m pytest tests/test_cpp_host.py tests/test_mp3_parity_gpu.py tests/test_abi_errors_gpu.py -m gpu -q 2>&1 | tail -8 | tee $out/${tag}_pytest.txt
./tests/cpp/decoder_host threads 64 /dev/null /dev/null 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_n1.json 2>$out/${tag}_bench_n1.err; tail -3 $out/${tag}_bench_n1.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_n1.json"))
print("mp3 value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],3), "med", round(d["e2e"]["ms_per_step_median"],3), "s16", round(d["e2e_s16"]["ms_per_step"],3), "compact", round(d["e2e_compact"]["ms_per_step"],3), "numa", d["numa"])
for k,c in d.get("configs",{}).items():
    print(k, "value", round(c["value"]), "e2e_ms", round(c["e2e"]["ms_per_step"],3), c.get("us_per_packet"))
PY
for sl in 6 12 16; do SYMGPU_SLICES=$sl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('slices $sl e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3))"; done
