#!/bin/bash
cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02p_build.log 2>&1 || { tail -20 gpurun_out/r02p_build.log; exit 1; }
for t in 16 15 14 13 12; do echo "T=$t"; SYMGPU_MP3_T=$t timeout 300 python tools/mp3_variant_bench.py v1p 2>&1 | grep -v "^{"; done
