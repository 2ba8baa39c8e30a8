#!/bin/bash
# Sweep of the samples in flight (BPT_LANES) and of the traversal kernels' persistent grid (BPT_TRACE_CTAS, CTAs per SM) on one B200.
O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
for w in materials terrain; do
  steps=64; [ $w = terrain ] && steps=16
  for cfg in "4 3" "4 4" "4 5" "4 6" "4 8" "6 3" "6 4" "3 5" "3 8" "2 8"; do set -- $cfg
    BPT_LANES=$1 BPT_TRACE_CTAS=$2 tools/bench_line.sh ${w}_lanes$1_trace$2 --steps $steps --warmup 8 --no-cpu-baseline --workload $w
  done
done
