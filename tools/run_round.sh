O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
for w in materials cornell; do
  for cfg in "4 3" "4 5" "6 4" "8 4" "6 3" "3 4" "4 2"; do set -- $cfg
    BPT_LANES=$1 BPT_TRACE_CTAS=$2 tools/bench_line.sh ${w}_lanes$1_trace$2 --steps 64 --warmup 8 --no-cpu-baseline --workload $w
  done
done
for cfg in "4 6" "6 8" "8 8" "6 6"; do set -- $cfg
  BPT_LANES=$1 BPT_TRACE_CTAS=$2 tools/bench_line.sh terrain_lanes$1_trace$2 --steps 16 --warmup 8 --no-cpu-baseline --workload terrain
done
cat $O/bench_lines.log
