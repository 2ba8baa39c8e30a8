O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
for w in materials cornell; do
  tools/bench_line.sh ${w}_lanes2 --steps 64 --warmup 3 --no-cpu-baseline --workload $w
  BPT_LANES=1 tools/bench_line.sh ${w}_lanes1 --steps 64 --warmup 3 --no-cpu-baseline --workload $w
done
tools/bench_line.sh terrain_lanes2 --steps 16 --warmup 3 --no-cpu-baseline --workload terrain
BPT_LANES=1 tools/bench_line.sh terrain_lanes1 --steps 16 --warmup 3 --no-cpu-baseline --workload terrain
cat $O/pytest_gpu.log; cat $O/bench_lines.log
