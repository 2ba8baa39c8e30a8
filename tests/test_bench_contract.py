"""bench.py contract checks that run without a GPU: the reference arm prints one well-formed JSON line."""
import json
import subprocess
import sys

import pytest

from tests import oracle_lib
from tests.oracle_lib import REPO


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--workload", "cornell_small", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_is_silent_on_nonzero_ranks():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "cornell_small"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_per_ray_matches_the_survey():
    sys.path.insert(0, str(REPO))
    import bench
    # 1 072 B at 20 k triangles, 1 392 B at 1 M, 1 776 B at 50 M
    assert bench.bytes_per_ray(20_000) == 1072 and bench.bytes_per_ray(1_000_000) == 1392 and bench.bytes_per_ray(50_000_000) == 1776
