O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
for w in materials cornell; do
  tools/bench_line.sh ${w}_nosort --steps 32 --warmup 3 --no-cpu-baseline --workload $w
  tools/bench_line.sh ${w}_sort0 --steps 32 --warmup 3 --no-cpu-baseline --workload $w --sort-hits 0
  tools/bench_line.sh ${w}_sort1 --steps 32 --warmup 3 --no-cpu-baseline --workload $w --sort-hits 1
  tools/bench_line.sh ${w}_sort2 --steps 32 --warmup 3 --no-cpu-baseline --workload $w --sort-hits 2
done
tools/bench_line.sh terrain_nosort --steps 8 --warmup 3 --no-cpu-baseline --workload terrain
tools/bench_line.sh terrain_sort1 --steps 8 --warmup 3 --no-cpu-baseline --workload terrain --sort-hits 1
cat $O/pytest_gpu.log; cat $O/bench_lines.log
