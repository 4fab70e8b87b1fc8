#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
shift
timeout 600 python tools/mp3_variant_bench.py "$@" 2>&1 | grep -v "^{" | tee $out/${tag}_variants.txt
