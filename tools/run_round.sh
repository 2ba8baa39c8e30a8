O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
tools/bench_line.sh materials_graph --steps 32 --warmup 3 --no-cpu-baseline
BPT_GRAPH=0 tools/bench_line.sh materials_serial --steps 32 --warmup 3 --no-cpu-baseline
tools/bench_line.sh cornell_graph --steps 32 --warmup 3 --no-cpu-baseline --workload cornell
BPT_GRAPH=0 tools/bench_line.sh cornell_serial --steps 32 --warmup 3 --no-cpu-baseline --workload cornell
tools/bench_line.sh terrain_graph --steps 8 --warmup 3 --no-cpu-baseline --workload terrain
cat $O/pytest_gpu.log; cat $O/bench_lines.log
