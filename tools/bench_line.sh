#!/bin/bash
# usage: [ENV=...] tools/bench_line.sh <label> <bench.py arguments...>
# Runs bench.py once and appends a one-line summary to gpurun_out/bench_lines.log (the full JSON line goes to
# gpurun_out/line_<label>.json). A failing run is reported with its exit code and the tail of its stderr, never silently dropped.
label=$1; shift
O=gpurun_out; mkdir -p $O
python bench.py "$@" > $O/line_$label.json 2> $O/line_$label.err; rc=$?
if [ $rc -ne 0 ]; then echo "$label FAILED rc=$rc: $(tail -5 $O/line_$label.err | tr '\n' ' ')" | tee -a $O/bench_lines.log; exit $rc; fi
tail -1 $O/line_$label.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; m=r['ms_per_step']
print('$label', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],3), 'ms/step | instrumented', round(m['instrumented_total'],3), 'extend', round(m['extend'],3), 'shade', round(m['shade'],3), 'shadow', round(m['shadow'],3), '| e2e', round(d['e2e']['value'],1), '| Grays/s', round(r['grays_per_s_extend'],2), 'frac', round(r['frac'],3), '| iter/step', round(d['iterations_per_step'],2), 'launches/step', round(d['gpu_launches_per_step'],1), 'nonfinite', d.get('nonfinite_samples'), 'build_ms', round(d['bvh']['build_ms'],2))" | tee -a $O/bench_lines.log
