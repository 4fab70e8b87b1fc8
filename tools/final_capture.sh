#!/bin/bash
# One-GPU measurement + profile capture for profiles/ (run on the GPU box through gpurun).
# usage: tools/final_capture.sh <tag>      e.g. r01c
cd "$(dirname "$0")/.."
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest_gpu.txt
python tools/pcie_probe.py > $out/${tag}_pcie_probe.json 2>/dev/null; cat $out/${tag}_pcie_probe.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference.json 2>$out/${tag}_bench_reference.err
timeout 400 python bench.py > $out/${tag}_bench_n1.json 2>$out/${tag}_bench_n1.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_n1.json"))
print("mp3 value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4),
      "e2e_ms", round(d["e2e"]["ms_per_step"],3), "e2e_s16_ms", round(d["e2e_s16"]["ms_per_step"],3), "e2e_compact_ms", round(d["e2e_compact"]["ms_per_step"],3),
      "cpu", round(d["cpu_baseline"]["value"]), d["clocks"])
r=json.load(open("$out/${tag}_bench_reference.json")); print("reference arm", round(r["value"]), r["cpu_baseline"]["cores"])
PY
timeout 400 python bench_codecs.py --codec all --steps 30 > $out/${tag}_bench_codecs.json 2>$out/${tag}_bench_codecs.err
cut -c1-40,100-260 $out/${tag}_bench_codecs.json
# launch lists (a number printed under ncu is never a bench value)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_mp3_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_codecs_launches.csv python bench_codecs.py --codec both --steps 4 --warmup 2 > /dev/null 2>&1
# full captures of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mp3_synth -c 1 -s 4 -o $out/${tag}_prof_mp3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_prof_mp3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"aac_synth|aac_tns_apply" -c 2 -s 6 -o $out/${tag}_prof_aac -f python bench_codecs.py --codec aac --steps 3 --warmup 3 > $out/${tag}_prof_aac.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vorbis_synth -c 1 -s 3 -o $out/${tag}_prof_vorbis -f python bench_codecs.py --codec vorbis --steps 3 --warmup 3 > $out/${tag}_prof_vorbis.log 2>&1
tail -1 $out/${tag}_prof_mp3.log; tail -1 $out/${tag}_prof_aac.log; tail -1 $out/${tag}_prof_vorbis.log
