#!/bin/bash
# usage: tools/r02_gpu_multi.sh <tag> <N>      (gpurun --gpus N)
cd "$(dirname "$0")/.."
tag=${1:-r02q}
N=${2:-2}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
nvidia-smi topo -m 2>/dev/null | head -12 > $out/${tag}_topo.txt
timeout 600 python -m pytest tests/test_cpp_host.py -m gpu -q -k "nccl" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_n$N.json 2>$out/${tag}_bench_n$N.err; tail -3 $out/${tag}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "mp3 value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["ranks_bit_exact_vs_oracle"], "numa", d["numa"])
for k,c in d.get("configs",{}).items():
    print(k, "value", round(c["value"]), "e2e", round(c["e2e"]["value"]), "e2e_ms", round(c["e2e"]["ms_per_step"],3), c.get("streams"))
PY
