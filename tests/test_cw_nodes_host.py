"""CPU test of the compressed eight-wide node format: bifrost3d_b200/csrc/bpt_cw.cuh is __host__ __device__, so the encoder and
the ray / node test that the traversal kernels run are compiled here with g++ and checked against exact boxes and against a
brute-force closest hit (tests/host/cw_host_test.cpp). The warp-level loop around them is covered by the -m gpu traversal tests."""
import shutil
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
CUDA_INCLUDE = Path("/usr/local/cuda/include")


@pytest.mark.skipif(shutil.which("g++") is None or not (CUDA_INCLUDE / "cuda_runtime.h").exists(), reason="needs g++ and the CUDA headers")
def test_encoder_and_node_test_on_the_host(tmp_path):
    exe = tmp_path / "cw_host_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", f"-I{CUDA_INCLUDE}", str(REPO / "tests" / "host" / "cw_host_test.cpp"), "-o", str(exe)],
                   check=True, timeout=300)
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(run.stdout)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert run.stdout.strip().endswith("OK")
