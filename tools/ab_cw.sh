#!/bin/bash
# A/B of the compressed eight-wide nodes against the four-wide ones on one B200 (run through gpurun): the GPU tests first,
# then one bench line per workload and node format, then the prepared build variants (bifrost3d_b200/variants/*.so, built
# here with `make -C bifrost3d_b200/csrc OUT=../variants/libbpt_<name>.so BUILD=build_<name> EXTRA="-D..."`: nospec =
# -DBPT_CW_SPECULATE=0 and the mbN = -DBPT_TRACE_MIN_BLOCKS=N variants of this script were measured in round 2 (profiles/
# r02_ab_node_formats.txt) and their switches have since been removed from the source; tools/ab_variants.sh is the general form).
# Every run keeps its exit code and stderr.
R=${1:-r02cw}
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -rA -x > $O/${R}_pytest_gpu_full.log 2>&1; echo "pytest rc=$?" > $O/${R}_pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|nested clusters" $O/${R}_pytest_gpu_full.log >> $O/${R}_pytest_gpu.log; cat $O/${R}_pytest_gpu.log
run() { # <name> <bench.py arguments...>
  local name=$1; shift
  python bench.py "$@" --no-cpu-baseline > $O/${R}_$name.all 2> $O/${R}_$name.err; local rc=$?
  grep '^{' $O/${R}_$name.all | tail -1 > $O/${R}_$name.json; rm -f $O/${R}_$name.all
  python - "$O/${R}_$name.json" "$name" "$rc" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); r = j["roofline"]
    print(sys.argv[2], "rc=" + sys.argv[3], "value %.1f e2e %.1f" % (j["value"], j["e2e"]["value"]), "width", j["bvh"].get("node_width"), "nodes", j["bvh"].get("traversed_nodes"), "levels", j["bvh"].get("levels"),
          "build %.2f ms" % j["bvh"]["build_ms"], "frac %.3f" % r["frac"], "extend Grays/s %.3f" % r["grays_per_s_extend"], {k: round(v, 3) for k, v in r["ms_per_step"].items()})
except Exception as e:
    print(sys.argv[2], "rc=" + sys.argv[3], "no line:", e)
PY
}
for wl in materials cornell; do
  run ${wl}_cw8 --steps 48 --warmup 3 --workload $wl
  BPT_CW=0 run ${wl}_w4 --steps 48 --warmup 3 --workload $wl
done
run terrain_cw8 --steps 12 --warmup 3 --workload terrain
BPT_CW=0 run terrain_w4 --steps 12 --warmup 3 --workload terrain
for v in nospec mb6 mb5; do
  for wl in materials cornell; do
    BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_$v.so run ${wl}_$v --steps 48 --warmup 3 --workload $wl
  done
done
