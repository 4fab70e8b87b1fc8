#!/bin/bash
# Round-2 measurement + profile capture for profiles/ (one GPU, through gpurun).  usage: tools/r02_capture.sh <tag> [quick]
cd "$(dirname "$0")/.."
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
if [ "$2" != "quick" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $out/${tag}_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $out/${tag}_bench_reference.json 2>$out/${tag}_bench_reference.err
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2>$out/${tag}_bench_n1.err; tail -2 $out/${tag}_bench_n1.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench_n1.json"))
print("mp3 value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],3), "cpu", round(d.get("cpu_baseline",{}).get("value",0)), d["clocks"])
for k,c in d.get("configs",{}).items():
    print(k, "value", round(c["value"]), "kernel_ms", c.get("kernel_ms"), "frac", round(c.get("roofline",{}).get("frac",0),4), "e2e", round(c["e2e"]["value"]), "cpu", round(c.get("cpu_baseline",{}).get("value",0)), c.get("us_per_packet"))
r=json.load(open("$out/${tag}_bench_reference.json")); print("reference arm", round(r["value"]), r["cpu_baseline"]["cores"])
PY
fi
# launch lists (a number printed under ncu is never a bench value)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full captures of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mp3_synth -c 1 -s 4 -o $out/${tag}_prof_mp3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs > $out/${tag}_prof_mp3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mp3v2_synth -c 1 -s 4 -o $out/${tag}_prof_mp3_serving -f python bench_codecs.py --codec mp3-short --steps 3 --warmup 3 > $out/${tag}_prof_mp3_serving.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"aac_synth|aac_tns_frames" -c 2 -s 6 -o $out/${tag}_prof_aac -f python bench_codecs.py --codec aac --steps 3 --warmup 3 > $out/${tag}_prof_aac.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vorbis_synth -c 1 -s 3 -o $out/${tag}_prof_vorbis -f python bench_codecs.py --codec vorbis --steps 3 --warmup 3 > $out/${tag}_prof_vorbis.log 2>&1
tail -1 $out/${tag}_prof_mp3.log; tail -1 $out/${tag}_prof_mp3_serving.log; tail -1 $out/${tag}_prof_aac.log; tail -1 $out/${tag}_prof_vorbis.log
ls -la $out | grep ${tag}
