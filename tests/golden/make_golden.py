#!/usr/bin/env python3
"""Generates the committed golden fixtures from the REFERENCE (oracle/_ref/libbifrost_ref.so).

Run here (where /root/reference exists and `make -C oracle` has been run):
    python -m tests.golden.make_golden
The fixtures travel to the GPU box, where they pin the CUDA path even if oracle/_ref is absent.
"""
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
N = 4096


def inputs():
    from bifrost3d_b200.workloads import bsdf_tuples
    return bsdf_tuples(N, seed=20261017, with_coat=True)


def rng_inputs():
    rng = np.random.default_rng(7)
    n = 2048
    acc = rng.integers(0, 1 << 20, n, dtype=np.uint32)
    acc[:64] = np.arange(64, dtype=np.uint32)
    return acc, rng.integers(0, 1 << 32, n, dtype=np.uint32), rng.integers(0, 80, n, dtype=np.uint32)


def generate(ref):
    t = inputs()
    out = {}
    for kind, name in enumerate(["default", "ggx_r", "oren_nayar", "burley"]):
        r = ref.bsdf_eval_sample_pdf(kind, t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"] if kind == 0 else None)
        for k, v in r.items():
            out[f"{name}_{k}"] = v
    acc, ph, dim = rng_inputs()
    ui, f = ref.sobol_sample4(acc, ph, dim)
    out["sobol_ui"] = ui
    out["sobol_f"] = f
    out["halton_offsets"] = ref.reverse_halton4(256)
    return out


def check(ref):
    stored = np.load(HERE / "bsdf_c1_small.npz")
    fresh = generate(ref)
    for k, v in fresh.items():
        assert np.array_equal(stored[k], v, equal_nan=True), k


if __name__ == "__main__":
    import sys
    sys.path.insert(0, str(HERE.parent.parent))
    from tests import oracle_lib
    np.savez_compressed(HERE / "bsdf_c1_small.npz", **generate(oracle_lib.load()))
    print("wrote", HERE / "bsdf_c1_small.npz")
