#!/bin/bash
cd "$(dirname "$0")/../.."
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1 || { tail -20 $out/${tag}_build.log; exit 1; }
python tools/pcie_probe.py > $out/${tag}_pcie_probe.json 2>/dev/null; cat $out/${tag}_pcie_probe.json | cut -c1-600
nvidia-smi topo -m 2>/dev/null | head -12
for nb in 0 1 0 1; do SYMGPU_NUMA_BIND=$nb timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('numa_bind $nb e2e ms', round(d['e2e']['ms_per_step'],3), round(d['e2e']['ms_per_step_median'],3), 's16', round(d['e2e_s16']['ms_per_step'],3), 'compact', round(d['e2e_compact']['ms_per_step'],3), d['numa']['node_rank0'], d['numa']['cpus_rank0'])"; done
timeout 600 python -m pytest tests/test_aac_vorbis_parity_gpu.py -m gpu -q 2>&1 | tail -3
