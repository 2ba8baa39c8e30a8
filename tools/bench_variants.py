#!/usr/bin/env python3
"""Runs bench.py once per library build in bifrost3d_b200/variants (kernel tuning experiments) and prints one line each."""
import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
libs = sorted((REPO / "bifrost3d_b200" / "variants").glob("*.so"))
args = sys.argv[1:] or ["--steps", "32", "--warmup", "3", "--no-cpu-baseline"]
for lib in [None] + libs:
    env = dict(os.environ)
    if lib:
        env["BPT_LIB"] = str(lib)
    out = subprocess.run([sys.executable, str(REPO / "bench.py")] + args, capture_output=True, text=True, env=env)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        r = d["roofline"]
        print(f"{(lib.name if lib else 'libbpt.so'):28s} {d['value']:8.1f} Msamples/s {d['ms_per_step']:7.3f} ms/step  extend {r['share_of_step']['extend'] * d['ms_per_step']:.3f} "
              f"shade {r['share_of_step']['shade'] * d['ms_per_step']:.3f} shadow {r['share_of_step']['shadow'] * d['ms_per_step']:.3f} ms  {r['grays_per_s_extend']:.2f} Grays/s"
              f"  nodes/ray {d.get('extend_node_visits_per_ray', 0):.1f} tris/ray {d.get('extend_triangle_tests_per_ray', 0):.1f}", flush=True)
    except Exception as e:
        print(lib, "failed", e, out.stderr[-500:])
