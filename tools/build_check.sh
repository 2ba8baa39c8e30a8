#!/bin/bash
# Build-path check on one B200: the traversal / incremental-update / sort tests, then the build time of every workload with the
# PLOC passes looped on the device (default) and with the host reading the pass state back (BPT_GRAPH=0 also turns the sample
# graphs off, so only the build times of those lines are compared).
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_traversal_parity.py tests/test_incremental_updates.py tests/test_sort.py tests/test_render_parity.py -m gpu -q -x 2>&1 | tail -3
for wl in cornell materials terrain; do
  for g in 1 0; do
    BPT_GRAPH=$g python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $wl 2> $O/build_${wl}_$g.err | grep '^{' | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl BPT_GRAPH=$g', d['bvh'], 'value', round(d['value'],1))"
  done
done
