#!/bin/bash
# Round measurements on one B200 (run through gpurun): the GPU test-suite and one bench line per workload, each with its exit
# code and stderr kept (gpurun_out/<tag>_*.json, .err). tools/profile_round.sh takes the ncu captures, tools/scale_round.sh
# the multi-GPU lines.
R=${1:-r02}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -rA > $O/${R}_pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/${R}_pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/${R}_pytest_gpu_full.log >> $O/${R}_pytest_gpu.log
run() { # <name> <bench.py arguments...>
  local name=$1; shift
  python bench.py "$@" > $O/${R}_bench_$name.all 2> $O/${R}_bench_$name.err; local rc=$?
  grep '^{' $O/${R}_bench_$name.all | tail -1 > $O/${R}_bench_$name.json; rm -f $O/${R}_bench_$name.all
  echo "$name rc=$rc $(cut -c1-160 $O/${R}_bench_$name.json)"
}
run materials --steps 64 --warmup 3
run materials_cdf_nee --steps 64 --warmup 3 --no-cpu-baseline --environment-nee cdf
run cornell --steps 64 --warmup 3 --workload cornell
run terrain --steps 16 --warmup 3 --workload terrain --no-cpu-baseline
run terrain_rr3 --steps 16 --warmup 3 --workload terrain --no-cpu-baseline --russian-roulette 3
run bsdf --steps 20 --warmup 3 --workload bsdf
run reference --impl reference --steps 3 --warmup 1
run reference_bsdf --impl reference --steps 3 --warmup 1 --workload bsdf
BPT_GRAPH=0 run materials_stream_launches --steps 32 --warmup 3 --no-cpu-baseline
cat $O/${R}_pytest_gpu.log
