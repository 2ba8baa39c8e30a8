"""Checkpoint / resume (bifrost3d_b200/checkpoint.py over bpt_read_accumulation / bpt_write_accumulation): a render resumed from
a saved state equals the uninterrupted one bit for bit."""
import numpy as np
import pytest

import bifrost3d_b200 as b
from bifrost3d_b200 import checkpoint, scenes


class FakeContext:
    def __init__(self, sums=None):
        self.sums = sums

    def read_accumulation(self):
        return self.sums

    def write_accumulation(self, sums):
        self.sums = np.array(sums)


def test_checkpoint_file_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    sums = rng.random((7, 5, 4))
    sums[..., 3] = 12.0
    path = tmp_path / "state.npz"
    assert checkpoint.save(FakeContext(sums), path, 12, scene="cornell") == (5, 7)
    restored = FakeContext()
    next_sample, meta = checkpoint.load(restored, path)
    assert next_sample == 12 and str(meta["scene"]) == "cornell"
    assert np.array_equal(restored.sums, sums) and restored.sums.dtype == np.float64
    np.savez(tmp_path / "bad.npz", format=np.int32(99), sums=sums, next_sample=np.uint32(0))
    with pytest.raises(ValueError):
        checkpoint.load(restored, tmp_path / "bad.npz")


@pytest.mark.gpu
def test_resumed_render_equals_the_uninterrupted_one(tmp_path):
    scene = scenes.cornell_box(sphere_quads=(16, 8))
    w, h = 64, 48
    first = b.Bpt(0)
    scenes.upload(first, scene)
    first.render(scene["camera"], w, h, 0, 5, reset=True)
    path = tmp_path / "render.npz"
    checkpoint.save(first, path, 5)
    state = first.read_accumulation()
    assert state.shape == (h, w, 4) and np.all(state[..., 3] == 5.0)
    first.close()

    resumed = b.Bpt(0)                       # a new context, as after a restart
    scenes.upload(resumed, scene)
    next_sample, _ = checkpoint.load(resumed, path)
    assert next_sample == 5
    resumed.render(scene["camera"], w, h, next_sample, 6)   # no reset: continues the restored sums
    got = resumed.resolve_float4()
    assert np.all(resumed.read_accumulation()[..., 3] == 11.0)
    resumed.close()

    whole = b.Bpt(0)
    scenes.upload(whole, scene)
    whole.render(scene["camera"], w, h, 0, 11, reset=True)
    want = whole.resolve_float4()
    whole.close()
    assert np.array_equal(got, want)
