#!/bin/bash
# Build-path check on one B200: the traversal / incremental-update / sort / render tests, then the build time of every workload
# by phase (BPT_BUILD_DEBUG) with the PLOC passes inside one cooperative launch (default) and as stream launches with a host
# read per pass (BPT_PLOC=host). Each process builds twice (warm-up render + bench): the second line of a pair is the warm one.
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_traversal_parity.py tests/test_incremental_updates.py tests/test_sort.py tests/test_render_parity.py -m gpu -q -x 2>&1 | tail -3
for wl in cornell materials terrain; do
  for mode in device host; do
    echo "== $wl BPT_PLOC=$mode"
    BPT_PLOC=$mode BPT_BUILD_DEBUG=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $wl 2>&1 | grep -E "bpt_build_accel" | tail -2
  done
done
