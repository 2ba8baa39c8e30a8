#!/bin/bash
# Weak scaling of the default workload on one 8-GPU box (the materials half of tools/scale_round.sh).
R=${1:-r01}; O=gpurun_out; mkdir -p $O
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 64 --warmup 3 2>/dev/null | tail -1 > $O/${R}_bench_materials_${n}gpu.json
done
python bench.py --steps 64 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_bench_materials_1gpu.json
for f in materials_1gpu materials_2gpu materials_4gpu materials_8gpu; do python -c "
import json; d=json.loads(open('$O/${R}_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; done
