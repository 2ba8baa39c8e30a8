#!/bin/bash
# Traversal statistics (stats build) for both node formats, then ncu captures of the traversal kernels of one sample.
R=${1:-r02cw}; O=gpurun_out; mkdir -p $O
for wl in materials cornell; do
  for cw in 1 0; do
    BPT_CW=$cw BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_stats.so python tools/traversal_stats.py $wl 2>> $O/${R}_stats.err | tee -a $O/${R}_traversal_stats.jsonl
  done
done
export BPT_GRAPH=0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload materials"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:extend_kernel -s 18 -c 6 -o $O/prof_${R}_extend_materials $B > $O/ncu_extend_materials.log 2>&1
$NCU -k regex:shadow_kernel -s 22 -c 2 -o $O/prof_${R}_shadow_materials $B > $O/ncu_shadow_materials.log 2>&1
ls -la $O/*.ncu-rep; tail -2 $O/ncu_extend_materials.log
