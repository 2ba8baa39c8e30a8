O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests/test_traversal_parity.py tests/test_render_parity.py -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
for w in materials cornell; do
  tools/bench_line.sh ${w} --steps 32 --warmup 3 --no-cpu-baseline --workload $w
done
cat $O/pytest_gpu.log; cat $O/bench_lines.log
