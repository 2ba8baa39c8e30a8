#!/bin/bash
# ncu captures of the round (run through gpurun, ONE GPU): launch list + `ncu --set full` of the traversal and shading kernels.
# The captures use plain stream launches (BPT_GRAPH=0): same kernels, same order, one launch per graph node. In that mode a
# sample is max_bounce_count + 2 iterations (the last one on empty queues), each = extend, shadow, shade_escaped, shade_surface.
R=${1:-r02}; W=${2:-materials}; IT=${3:-6}
O=gpurun_out; mkdir -p $O
export BPT_GRAPH=0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"generate_kernel|extend_kernel|shade_kernel|shadow_kernel|advance_kernel|accumulate_kernel|sort_|set_frame" -c 800 --csv --log-file $O/${R}_launches_$W.csv $B > $O/ncu_launches_$W.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
# skip the 3 warm-up samples, capture every launch of one whole sample (so that per-launch averages use the same launches as
# bench.py's algorithmic figure)
$NCU -k regex:extend_kernel -s $((3 * IT)) -c $IT -o $O/prof_${R}_extend_$W $B > $O/ncu_extend_$W.log 2>&1
if [ "$W" = "materials" ]; then
  $NCU -k regex:shadow_kernel -s $((3 * (IT + 1) + 1)) -c 2 -o $O/prof_${R}_shadow_$W $B > $O/ncu_shadow_$W.log 2>&1
  $NCU -k regex:shade_kernel -s $((3 * 2 * IT)) -c 4 -o $O/prof_${R}_shade_$W $B > $O/ncu_shade_$W.log 2>&1
fi
ls -la $O/*.ncu-rep; tail -2 $O/ncu_extend_$W.log
