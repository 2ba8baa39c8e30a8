"""Smoke test body: a small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np


def run():
    import bifrost3d_b200 as b
    from bifrost3d_b200.workloads import bsdf_tuples
    from tests import oracle_lib
    from tests.parity import rel_err

    ctx = b.Bpt(0)
    t = bsdf_tuples(4096, seed=1)
    got = ctx.bsdf_eval_sample_pdf(0, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
    ref = oracle_lib.load()
    want = ref.bsdf_eval_sample_pdf(0, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
    e = rel_err(got["eval_f"], want["eval_f"], 1e-6)
    frac = float(np.mean(e > 1e-5))
    print(f"smoke: DefaultShading eval over 4096 tuples, fraction above 1e-5 rel: {frac:.2e}")
    assert frac < 1e-3
    ctx.close()
