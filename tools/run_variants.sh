#!/bin/bash
# On the GPU box: parity tests for every library in bifrost3d_b200/variants/, then one bench line per workload and library.
for lib in bifrost3d_b200/variants/*.so; do
  echo "== $lib"
  BPT_LIB=$PWD/$lib python -m pytest tests/test_traversal_parity.py tests/test_render_parity.py tests/test_shading_parity.py -m gpu -q 2>&1 | tail -2
done
tools/bench_variants_line.sh materials 32
tools/bench_variants_line.sh cornell 32
tools/bench_variants_line.sh terrain 8
