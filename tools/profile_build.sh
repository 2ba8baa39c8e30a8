#!/bin/bash
# ncu launch list of the acceleration-structure build of one workload (the first launches of the process).
W=${1:-materials}; O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_build_$W.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $W > $O/ncu_build_$W.log 2>&1
python tools/summarize_ncu.py launches $O/r02_launches_build_$W.csv $O/r02_launches_build_$W.md | head -40
