#!/bin/bash
# Scaling on one multi-GPU box: bench.py --gpus N under torchrun for every N given (default 1 2 4 8), one JSON line per N in
# gpurun_out/<tag>_bench_<workload>_<N>gpu.json. stderr is kept next to it and a failing rank fails the script.
# usage: [EXTRA='--spp-total 1024'] tools/scale_round.sh <tag> <workload> <steps> [N ...]
R=${1:-r02}; W=${2:-materials}; STEPS=${3:-64}; shift 3 2>/dev/null
NS=${@:-1 2 4 8}
O=gpurun_out; mkdir -p $O
status=0
for n in $NS; do
  out=$O/${R}_bench_${W}_${n}gpu.json; err=$O/${R}_bench_${W}_${n}gpu.err
  if [ $n -eq 1 ]; then python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --workload $W $EXTRA > $out.all 2> $err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps $STEPS --warmup 3 --workload $W $EXTRA > $out.all 2> $err
  fi
  rc=$?
  if [ $rc -ne 0 ]; then echo "N=$n FAILED rc=$rc"; tail -15 $err; status=$rc; continue; fi
  grep '^{' $out.all | tail -1 > $out; rm -f $out.all
  python -c "
import json; d=json.loads(open('$out').read()); print('$W N=$n', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms/step  e2e', round(d['e2e']['value'],1), 'nonfinite', d.get('nonfinite_samples'), 'frames_nonfinite', d.get('frames_with_nonfinite_pixels'))"
done
exit $status
