O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
for lib in "" bifrost3d_b200/variants/*.so; do
  [ -n "$lib" ] && export BPT_LIB=$PWD/$lib || unset BPT_LIB
  name=$(basename "${lib:-default}" .so)
  python bench.py --workload bsdf --steps 10 --warmup 3 --no-cpu-baseline > $O/line_bsdf_$name.json 2> $O/line_bsdf_$name.err; echo "bsdf $name rc=$?" | tee -a $O/bench_lines.log
  tail -1 $O/line_bsdf_$name.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ', round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms  frac', round(d['roofline']['frac'],4), 'others', {k: round(v) for k, v in d['other_bsdfs_mtuples_per_s'].items()})" | tee -a $O/bench_lines.log
done
unset BPT_LIB
cat $O/bench_lines.log
