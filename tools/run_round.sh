O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
python -m pytest tests -m gpu -q -rA > $O/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu_full.log >> $O/pytest_gpu.log
for w in materials cornell; do
  tools/bench_line.sh ${w}_new --steps 32 --warmup 3 --no-cpu-baseline --workload $w
  BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_oldslab.so tools/bench_line.sh ${w}_oldslab --steps 32 --warmup 3 --no-cpu-baseline --workload $w
done
tools/bench_line.sh terrain_new --steps 8 --warmup 3 --no-cpu-baseline --workload terrain
BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_oldslab.so tools/bench_line.sh terrain_oldslab --steps 8 --warmup 3 --no-cpu-baseline --workload terrain
python bench.py --workload bsdf --steps 10 --warmup 3 > $O/line_bsdf.json 2> $O/line_bsdf.err; echo "bsdf rc=$?" | tee -a $O/bench_lines.log; tail -1 $O/line_bsdf.json | cut -c1-900 | tee -a $O/bench_lines.log
cat $O/pytest_gpu.log; cat $O/bench_lines.log
