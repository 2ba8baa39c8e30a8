#!/bin/bash
# Probe: compute-sanitizer memcheck on one small render test in graph mode, by number of samples in flight (BPT_LANES) and
# with the tool's launches forced to block. Context for tools/sanitizer_round.sh, which runs the suite with stream launches.
O=gpurun_out; mkdir -p $O
probe() { # <label> <extra sanitizer flags...>
  local label=$1; shift
  timeout 600 compute-sanitizer --tool memcheck --print-limit 5 "$@" --log-file $O/san_probe_$label.log python -m pytest tests/test_render_parity.py -m gpu -q -x -k "cornell_box_matches" > $O/san_probe_$label.pytest.log 2>&1
  echo "$label: pytest rc=$? | $(grep -c 'Invalid __' $O/san_probe_$label.log) memcheck access records | $(grep 'ERROR SUMMARY' $O/san_probe_$label.log | head -1)"
}
for lanes in 1 2 3 4; do BPT_GRAPH=1 BPT_LANES=$lanes probe lanes$lanes; done
BPT_GRAPH=1 BPT_LANES=4 probe lanes4_blocking --force-blocking-launches yes
