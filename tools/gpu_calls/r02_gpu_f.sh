#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 600 python tools/mp3_variant_bench.py v1 12:33 14:33 14:97 10:33 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2>$out/${tag}_bench_n1.err; tail -3 $out/${tag}_bench_n1.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_n1.json"))
print("mp3 value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ranks_bit_exact_vs_oracle"], "cpu", round(d.get("cpu_baseline",{}).get("value",0)))
for k,c in d.get("configs",{}).items():
    print(k, "value", round(c["value"]), "kernel_ms", c.get("kernel_ms"), "frac", round(c.get("roofline",{}).get("frac",0),4), "e2e", round(c["e2e"]["value"]), "e2e_ms", round(c["e2e"]["ms_per_step"],3), "cpu", round(c.get("cpu_baseline",{}).get("value",0)), c.get("us_per_packet"))
PY
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $out/${tag}_pytest_gpu.txt
