#!/bin/bash
# A/B of the four-wide nodes against the binary nodes (BPT_WIDE=0), same library.
python -m pytest tests/test_traversal_parity.py tests/test_render_parity.py -m gpu -q -x 2>&1 | tail -3
for w in materials cornell terrain; do
  steps=32; [ $w = terrain ] && steps=8
  for wide in 0 1; do
    BPT_WIDE=$wide python bench.py --steps $steps --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w wide=$wide', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms  extend', round(r['share_of_step']['extend']*d['ms_per_step'],3), 'shade', round(r['share_of_step']['shade']*d['ms_per_step'],3), 'shadow', round(r['share_of_step']['shadow']*d['ms_per_step'],3), 'build_ms', round(d['bvh']['build_ms'],2))"
  done
done
