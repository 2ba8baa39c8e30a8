O=gpurun_out; mkdir -p $O; : > $O/bench_lines.log
for w in materials cornell; do for c in 1 2 3 4; do BPT_SHADE_CTAS=$c tools/bench_line.sh ${w}_shadectas$c --steps 64 --warmup 8 --no-cpu-baseline --workload $w > /dev/null 2>&1; done; done
cut -c1-150 $O/bench_lines.log
