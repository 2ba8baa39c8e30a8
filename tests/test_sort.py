"""The device-wide primitives of the BVH build (bpt_sort.cuh), bit-exact against numpy: integer work, no tolerance."""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 255, 2048, 2049, 100_003, 1 << 20])
def test_radix_sort_pairs_matches_stable_argsort(bpt, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
    if n > 1000:  # many duplicates and long runs: stability and the digit-run write-out
        keys[: n // 3] = keys[0]
        keys[n // 3: n // 2] &= np.uint64(0xFF)
    values = np.arange(n, dtype=np.uint32)
    k, v = bpt.sort_pairs(keys, values)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, values[order])


@pytest.mark.gpu
def test_radix_sort_bit_range(bpt):
    """Only bits [8, 24) take part; everything else keeps its input order (least-significant-digit passes are stable)."""
    rng = np.random.default_rng(3)
    n = 50_000
    keys = rng.integers(0, 1 << 40, n, dtype=np.uint64)
    values = np.arange(n, dtype=np.uint32)
    k, v = bpt.sort_pairs(keys, values, 8, 24)
    order = np.argsort((keys >> np.uint64(8)) & np.uint64(0xFFFF), kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, values[order])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 7, 2048, 2049, 1_000_001])
def test_exclusive_scan_matches_cumsum(bpt, n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 5, n, dtype=np.uint32)
    out, total = bpt.exclusive_scan(v)
    want = np.concatenate([[0], np.cumsum(v, dtype=np.uint64)[:-1]]).astype(np.uint32)
    assert np.array_equal(out, want)
    assert total == int(v.sum())
