#!/bin/bash
# Round measurements on one B200 (run through gpurun): bench lines, the ncu launch list and `ncu --set full` captures.
# Everything lands in gpurun_out/; tools/summarize_ncu.py turns the captures into the summaries under profiles/.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py --steps 64 --warmup 3 2>$O/bench_materials.err | tail -1 > $O/${R}_bench_materials.json
python bench.py --steps 64 --warmup 3 --workload cornell --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_bench_cornell.json
python bench.py --steps 16 --warmup 3 --workload terrain --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_bench_terrain.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/${R}_bench_reference.json
python tools/bench_bsdf.py 2>/dev/null | tail -1 > $O/${R}_bench_bsdf_tuples.json
BPT_BVH=lbvh python bench.py --steps 32 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${R}_bench_materials_lbvh.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"generate_kernel|extend_kernel|shade_kernel|shadow_kernel|advance_kernel|accumulate_kernel" -c 600 --csv --log-file $O/${R}_launches_materials.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_a.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:extend_kernel -s 6 -c 3 -o $O/prof_${R}_extend_materials python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_c.log 2>&1
$NCU -k regex:shade_kernel -s 12 -c 4 -o $O/prof_${R}_shade_materials python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_d.log 2>&1
$NCU -k regex:shadow_kernel -s 6 -c 2 -o $O/prof_${R}_shadow_materials python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_f.log 2>&1
$NCU -k regex:extend_kernel -s 10 -c 12 -o $O/prof_${R}_extend_terrain python bench.py --workload terrain --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_g.log 2>&1
$NCU -k regex:ploc_nearest_kernel -c 2 -o $O/prof_${R}_ploc_nearest python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_h.log 2>&1
ls -la $O/*.ncu-rep
for f in materials cornell terrain reference bsdf_tuples materials_lbvh; do echo $f; cut -c1-300 $O/${R}_bench_$f.json; done
