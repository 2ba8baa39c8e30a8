"""World-size-2 CPU test (gloo) of the sample-sharding host logic used by bench.py: disjoint sample ranges per rank and a
SUM reduce of the double4 accumulation buffers reproduce the single-process result. Runs the CPU oracle as the renderer
stand-in (the sharding logic is renderer independent: the sample value is a pure function of pixel and sample index)."""
import os
import sys

import numpy as np
import pytest

from tests import oracle_lib

pytestmark = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")

W, H, SPP_PER_RANK = 24, 16, 2


from bifrost3d_b200.sharding import sample_range


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(oracle_lib.REPO))
    from bifrost3d_b200 import scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    sc = oracle_lib.OracleScene(scene)
    first, count = sample_range(rank, SPP_PER_RANK)
    accum, _ = sc.render(scene["camera"], W, H, first, count, threads=1)
    t = torch.from_numpy(accum)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_ranges_are_disjoint_and_cover():
    seen = []
    for r in range(8):
        first, count = sample_range(r, 128)
        seen += list(range(first, first + count))
    assert seen == list(range(8 * 128))


def test_two_ranks_reduce_to_the_single_process_image(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "reduced.npy"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    reduced = np.load(out)
    from bifrost3d_b200 import scenes
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    sc = oracle_lib.OracleScene(scene)
    single, _ = sc.render(scene["camera"], W, H, 0, 2 * SPP_PER_RANK, threads=1)
    assert np.array_equal(reduced[..., 3], single[..., 3])
    # fp64 sums of the same float samples in a different association: equal to ~1e-15 relative
    assert np.allclose(reduced, single, rtol=1e-13, atol=1e-15)
