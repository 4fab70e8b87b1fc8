#!/bin/bash
# first-generation MP3 kernel with 18 warps (16 granule jobs + 2 helpers in the DCT / window phases) against 16
cd "$(dirname "$0")/../.."
for nw in 18 16; do
  make -C symphonia_b200/csrc -B EXTRA="-DSYMGPU_MP3_T=16 -DSYMGPU_MP3_NW=$nw" > gpurun_out/r02y_build_$nw.log 2>&1 || { tail -5 gpurun_out/r02y_build_$nw.log; exit 1; }
  echo "== NW=$nw"
  timeout 600 python -m pytest tests/test_mp3_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
  timeout 300 python tools/mp3_variant_bench.py v1p 2>&1 | grep -v "^{" | tail -3
done
