O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
for lib in "" bifrost3d_b200/variants/*.so; do
  [ -n "$lib" ] && export BPT_LIB=$PWD/$lib || unset BPT_LIB
  name=$(basename "${lib:-default}" .so)
  for w in materials cornell; do
    tools/bench_line.sh ${w}_$name --steps 32 --warmup 3 --no-cpu-baseline --workload $w
  done
done
cat $O/pytest_gpu.log; cat $O/bench_lines.log
