#!/bin/bash
# Host-side sanitizers (no GPU): the oracle library (the reference's headers behind a C API + the restated integrator and
# intersector) and the eight-wide node harness built with AddressSanitizer + UndefinedBehaviorSanitizer, then the CPU tests
# that use the oracle, one oracle render and one brute-force intersection batch run against that build.
set -e
cd "$(dirname "$0")/.."
OUT=${1:-/tmp/oracle_asan}; mkdir -p $OUT
make -C oracle OUT=$OUT OPT="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer" $OUT/libbifrost_ref.so -j8 > $OUT/build.log 2>&1
g++ -O1 -g -std=c++17 -ffp-contract=off -fsanitize=address,undefined -I/usr/local/cuda/include tests/host/cw_host_test.cpp -o $OUT/cw_host_test
export BPT_ORACLE_LIB=$OUT/libbifrost_ref.so LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0
$OUT/cw_host_test > $OUT/cw_host_test.log 2>&1; echo "cw_host_test rc=$?"
python -m pytest tests/test_oracle_reference.py tests/test_environment.py tests/test_traversal_parity.py tests/test_image_io_compare.py tests/test_tonemap.py -q -m "not gpu" > $OUT/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $OUT/pytest.log)"
python - > $OUT/render.log 2>&1 <<'PY'
import numpy as np
from bifrost3d_b200 import scenes
from tests import oracle_lib
scene = scenes.cornell_box(sphere_quads=(16, 8))
sc = oracle_lib.OracleScene(scene)
accum, rays = sc.render(scene["camera"], 48, 40, 0, 3)
rng = np.random.default_rng(1)
o = rng.uniform(-0.45, 0.45, (20000, 3)).astype(np.float32); d = rng.normal(size=(20000, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
p_brute = sc.intersect(o, d.astype(np.float32), brute=True)[0]; p_bvh = sc.intersect(o, d.astype(np.float32), brute=False)[0]
assert np.array_equal(p_brute, p_bvh) and np.isfinite(accum).all()
sc.close()
print("oracle render + intersect ok", rays)
PY
echo "render rc=$? $(tail -1 $OUT/render.log)"
echo "sanitizer reports: $(cat $OUT/*.log | grep -c -E 'runtime error|AddressSanitizer|ERROR: ' )"
