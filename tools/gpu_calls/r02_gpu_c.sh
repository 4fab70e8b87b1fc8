#!/bin/bash
# Round 2, GPU call C: window-chain interleave variants + ncu of one of them
cd "$(dirname "$0")/../.."
tag=${1:-r02c}
prof=${2:-12:9}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
timeout 600 python tools/mp3_variant_bench.py v1 12:0 12:1 12:4 12:5 12:7 12:8 12:9 12:11 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
SYMGPU_MP3_V2_VARIANT=$prof timeout 400 ncu --set full --clock-control none --import-source on -k regex:mp3v2_synth -c 1 -s 4 -o $out/${tag}_prof_mp3 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_prof_mp3.log 2>&1
tail -2 $out/${tag}_prof_mp3.log
