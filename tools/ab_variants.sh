#!/bin/bash
# A/B of prepared build variants (bifrost3d_b200/variants/libbpt_<name>.so) on one B200: tools/ab_variants.sh <tag> "<workloads>" <variant>...
# "default" = the in-tree libbpt.so.
R=$1; WLS=$2; shift 2
O=gpurun_out; mkdir -p $O
run() { # <name> <bench.py arguments...>
  local name=$1; shift
  python bench.py "$@" --no-cpu-baseline > $O/${R}_$name.all 2> $O/${R}_$name.err; local rc=$?
  grep '^{' $O/${R}_$name.all | tail -1 > $O/${R}_$name.json; rm -f $O/${R}_$name.all
  python - "$O/${R}_$name.json" "$name" "$rc" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); r = j["roofline"]
    print(sys.argv[2], "rc=" + sys.argv[3], "value %.1f e2e %.1f" % (j["value"], j["e2e"]["value"]), "width", j["bvh"].get("node_width"),
          "extend Grays/s %.3f" % r["grays_per_s_extend"], {k: round(v, 3) for k, v in r["ms_per_step"].items()})
except Exception as e:
    print(sys.argv[2], "rc=" + sys.argv[3], "no line:", e)
PY
}
for v in "$@"; do
  for wl in $WLS; do
    steps=48; [ $wl = terrain ] && steps=12
    if [ $v = default ]; then run ${wl}_$v --steps $steps --warmup 3 --workload $wl
    else BPT_LIB=$PWD/bifrost3d_b200/variants/libbpt_$v.so run ${wl}_$v --steps $steps --warmup 3 --workload $wl; fi
  done
done
