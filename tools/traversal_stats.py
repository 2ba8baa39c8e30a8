"""Node visits and triangle tests per closest-hit ray for one workload (needs a libbpt.so built with -DBPT_TRAVERSAL_STATS:
`make -C bifrost3d_b200/csrc OUT=../variants/libbpt_stats.so BUILD=build_stats EXTRA=-DBPT_TRAVERSAL_STATS`, then
BPT_LIB=bifrost3d_b200/variants/libbpt_stats.so python tools/traversal_stats.py materials)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bifrost3d_b200 as b  # noqa: E402
from bifrost3d_b200 import scenes  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "materials"
scene = {"materials": scenes.material_grid, "cornell": scenes.cornell_box, "terrain": scenes.instanced_terrain}[workload]()
ctx = b.Bpt(0)
scenes.upload(ctx, scene)
info = ctx.accel_info()
bounces = 8 if workload == "terrain" else 4
ctx.counters(reset=True)
ctx.render(scene["camera"], scene["width"], scene["height"], 0, 4, max_bounces=bounces, reset=True)
ctx.resolve_float4()
c = ctx.counters()
print(json.dumps({"workload": workload, "BPT_CW": os.environ.get("BPT_CW", "1"), "node_width": info["node_width"], "traversed_nodes": info["traversed_nodes"], "levels": info["levels"],
                  "extend_rays": c["extend_rays"], "node_visits_per_ray": c["extend_node_visits"] / max(c["extend_rays"], 1),
                  "triangle_tests_per_ray": c["extend_triangle_tests"] / max(c["extend_rays"], 1)}))
ctx.close()
