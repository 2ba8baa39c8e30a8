#!/bin/bash
# On the GPU box: the GPU test-suite on the default build, then for every library in bifrost3d_b200/variants/ its parity tests and
# one bench line per workload. Everything (stderr and exit codes included) lands in gpurun_out/variants.log.
O=gpurun_out; mkdir -p $O
L=$O/variants.log
: > $L
echo "== default build: pytest -m gpu" | tee -a $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a $L
echo "rc=${PIPESTATUS[0]}" | tee -a $L
for lib in bifrost3d_b200/variants/*.so; do
  echo "== $lib" | tee -a $L
  BPT_LIB=$PWD/$lib python -m pytest tests/test_traversal_parity.py tests/test_render_parity.py tests/test_shading_parity.py -m gpu -q 2>&1 | tail -3 | tee -a $L
done
for w in ${WORKLOADS:-materials cornell}; do
  for lib in "" bifrost3d_b200/variants/*.so; do
    [ -n "$lib" ] && export BPT_LIB=$PWD/$lib || unset BPT_LIB
    python bench.py --steps 32 --warmup 3 --no-cpu-baseline --workload $w > $O/line.json 2> $O/line.err; rc=$?
    if [ $rc -ne 0 ]; then echo "$w ${lib:-default} FAILED rc=$rc: $(tail -3 $O/line.err)" | tee -a $L; continue; fi
    tail -1 $O/line.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w ${lib:-default}', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms  extend', round(r['share_of_step']['extend']*d['ms_per_step'],3), 'shade', round(r['share_of_step']['shade']*d['ms_per_step'],3), 'shadow', round(r['share_of_step']['shadow']*d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'nonfinite', d.get('nonfinite_samples'))" | tee -a $L
  done
done
