#!/bin/bash
# usage: bench_variants_line.sh <workload> <steps>: one line per library in bifrost3d_b200/variants (and the default build)
w=$1; steps=$2
for lib in "" bifrost3d_b200/variants/*.so; do
  [ -n "$lib" ] && export BPT_LIB=$PWD/$lib || unset BPT_LIB
  python bench.py --steps $steps --warmup 3 --no-cpu-baseline --workload $w 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$w ${lib:-default}', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms  extend', round(r['share_of_step']['extend']*d['ms_per_step'],3), 'shade', round(r['share_of_step']['shade']*d['ms_per_step'],3), 'shadow', round(r['share_of_step']['shadow']*d['ms_per_step'],3), 'build_ms', round(d['bvh']['build_ms'],2))"
done
