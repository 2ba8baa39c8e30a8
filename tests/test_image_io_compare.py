"""image output (PNG / EXR) and ImageOperations::Compare (rms / ssim / mssim) on the device against the
reference's own header compiled for the host."""
import numpy as np
import pytest

from tests import oracle_lib
from bifrost3d_b200 import image_io

needs_oracle = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")


def test_png_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    for channels in (3, 4):
        pixels = rng.integers(0, 256, (37, 53, channels), dtype=np.uint8)
        path = tmp_path / f"image{channels}.png"
        image_io.write_png(path, pixels)
        assert np.array_equal(image_io.read_png(path), pixels)


def test_png_reader_handles_every_filter_type(tmp_path):
    """A PNG written by another encoder uses the sub / up / average / Paeth filters: build one by hand."""
    import struct, zlib
    rng = np.random.default_rng(2)
    h, w, c = 10, 7, 3
    pixels = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    rows, previous = [], np.zeros(w * c, np.int32)
    for y in range(h):
        line = pixels[y].reshape(-1).astype(np.int32)
        kind = y % 5
        left = np.concatenate([np.zeros(c, np.int32), line[:-c]])
        up_left = np.concatenate([np.zeros(c, np.int32), previous[:-c]])
        if kind == 0: predictor = 0
        elif kind == 1: predictor = left
        elif kind == 2: predictor = previous
        elif kind == 3: predictor = (left + previous) >> 1
        else:
            estimate = left + previous - up_left
            pa, pb, pc = np.abs(estimate - left), np.abs(estimate - previous), np.abs(estimate - up_left)
            predictor = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, previous, up_left))
        rows.append(bytes([kind]) + ((line - predictor) & 255).astype(np.uint8).tobytes())
        previous = line
    chunk = image_io._chunk
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"".join(rows))) + chunk(b"IEND", b"")
    path = tmp_path / "filters.png"
    path.write_bytes(data)
    assert np.array_equal(image_io.read_png(path, flip_vertically=False), pixels)


@pytest.mark.parametrize("half", [True, False])
def test_exr_round_trip(tmp_path, half):
    rng = np.random.default_rng(3)
    pixels = (rng.random((19, 31, 4)) * 100).astype(np.float32)
    path = tmp_path / "image.exr"
    image_io.write_exr(path, pixels, half=half)
    back = image_io.read_exr(path)
    expected = pixels.astype(np.float16).astype(np.float32) if half else pixels
    assert np.array_equal(back, expected)
    assert path.read_bytes()[:4] == bytes([0x76, 0x2F, 0x31, 0x01])  # the OpenEXR magic number


@needs_oracle
def test_reference_compare_sanity():
    """The oracle binding itself: identical images compare as rms 0 / ssim 1."""
    lib = oracle_lib.load()
    image = np.random.default_rng(4).random((16, 24, 4)).astype(np.float32)
    r = oracle_lib.reference_compare_images(image, image, 3)
    assert r["rms"] == 0.0 and abs(r["ssim"] - 1.0) < 1e-6 and abs(r["mssim"] - 1.0) < 1e-6


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("shape,support", [((48, 64), 4), ((135, 240), 6), ((7, 5), 2)])
def test_device_compare_matches_the_reference_header(bpt, shape, support):
    rng = np.random.default_rng(5)
    reference = rng.random(shape + (4,)).astype(np.float32) * 2.0
    target = np.clip(reference + rng.normal(scale=0.05, size=reference.shape).astype(np.float32), 0, None).astype(np.float32)
    want = oracle_lib.reference_compare_images(reference, target, support, diff_images=True)
    got = bpt.compare_images(reference, target, support, diff_images=True)
    for key in ("rms", "ssim", "mssim"):
        print(key, got[key], want[key])
        assert abs(got[key] - want[key]) <= 1e-6 * max(1.0, abs(want[key])), key
    assert np.array_equal(got["rms_diff"][..., :3], want["rms_diff"][..., :3])
    assert np.abs(got["mssim_diff"][..., :3] - want["mssim_diff"][..., :3]).max() <= 2e-6


@pytest.mark.gpu
def test_render_to_png_and_exr(bpt, tmp_path):
    """Tonemapped sRGB frame -> PNG, linear frame -> EXR, both read back."""
    from bifrost3d_b200 import scenes
    scene = scenes.cornell_box(sphere_quads=(16, 8))
    scenes.upload(bpt, scene)
    bpt.render(scene["camera"], 64, 48, 0, 8, reset=True)
    rgba8 = bpt.resolve_tonemapped("filmic", rgba8=True)
    linear = bpt.resolve_float4()
    image_io.write_png(tmp_path / "frame.png", rgba8); image_io.write_exr(tmp_path / "frame.exr", linear, half=False)
    assert np.array_equal(image_io.read_png(tmp_path / "frame.png"), rgba8)
    assert np.array_equal(image_io.read_exr(tmp_path / "frame.exr"), linear)
    assert rgba8[..., :3].mean() > 5
