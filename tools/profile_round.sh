#!/bin/bash
# ncu captures of the round (run through gpurun, ONE GPU): launch list + `ncu --set full` of the traversal and shading kernels.
# The captures use plain stream launches (BPT_GRAPH=0): same kernels, same order, one launch per graph node.
R=${1:-r02}; W=${2:-materials}
O=gpurun_out; mkdir -p $O
export BPT_GRAPH=0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"generate_kernel|extend_kernel|shade_kernel|shadow_kernel|advance_kernel|accumulate_kernel|sort_|set_frame" -c 800 --csv --log-file $O/${R}_launches_$W.csv $B > $O/ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
# after 3 warm-up samples x 6 iterations: launches 18.. are sample 3; capture camera rays, bounce 1, bounce 2
$NCU -k regex:extend_kernel -s 18 -c 3 -o $O/prof_${R}_extend_$W $B > $O/ncu_extend.log 2>&1
$NCU -k regex:shadow_kernel -s 22 -c 2 -o $O/prof_${R}_shadow_$W $B > $O/ncu_shadow.log 2>&1
$NCU -k regex:shade_kernel -s 36 -c 4 -o $O/prof_${R}_shade_$W $B > $O/ncu_shade.log 2>&1
ls -la $O/*.ncu-rep; tail -2 $O/ncu_extend.log
