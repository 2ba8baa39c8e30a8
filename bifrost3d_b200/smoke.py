"""Smoke test body: a small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np


def run():
    import bifrost3d_b200 as b
    from bifrost3d_b200 import scenes
    from bifrost3d_b200.workloads import bsdf_tuples
    from tests import oracle_lib
    from tests.parity import rel_err

    ctx = b.Bpt(0)
    # 1. one small progressive render of the Cornell box (BVH build, wavefront integrator), vs the CPU oracle
    scene = scenes.cornell_box(sphere_quads=(24, 12))
    scenes.upload(ctx, scene)
    ctx.render(scene["camera"], 64, 64, 0, 4, reset=True)
    gpu = ctx.resolve_float4()[..., :3]
    counters = ctx.counters()
    assert np.isfinite(gpu).all() and gpu.mean() > 0.01
    if oracle_lib.available():
        sc = oracle_lib.OracleScene(scene)
        accum, oc = sc.render(scene["camera"], 64, 64, 0, 4)
        sc.close()
        cpu = accum[..., :3] / accum[..., 3:4]
        rel_mse = float(np.mean((gpu - cpu) ** 2 / (cpu ** 2 + 1e-2)))
        print(f"smoke: Cornell 64x64x4spp relMSE vs oracle {rel_mse:.3e}; rays {counters['extend_rays']}+{counters['shadow_rays']} (oracle {oc[0]}+{oc[1]})")
        assert rel_mse < 1e-3
        # 2. a BSDF batch vs the reference's host-compiled headers
        t = bsdf_tuples(4096, seed=1)
        got = ctx.bsdf_eval_sample_pdf(0, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
        want = oracle_lib.load().bsdf_eval_sample_pdf(0, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
        frac = float(np.mean(rel_err(got["eval_f"], want["eval_f"], 1e-6) > 1e-5))
        print(f"smoke: DefaultShading eval over 4096 tuples, fraction above 1e-5 rel: {frac:.2e}")
        assert frac < 1e-3
    else:
        print("smoke: oracle/_ref not available, rendered without comparison")
    # 3. the same render through the compressed eight-wide nodes (what scenes from 131 072 triangles on get): closest hits are a
    #    minimum over (t, primitive), so the image must not depend on the node format
    import os
    before = os.environ.get("BPT_CW")
    os.environ["BPT_CW"] = "1"
    try:
        wide = b.Bpt(0)
    finally:
        if before is None:
            del os.environ["BPT_CW"]
        else:
            os.environ["BPT_CW"] = before
    scenes.upload(wide, scene)
    assert wide.accel_info()["node_width"] == 8 and ctx.accel_info()["node_width"] == 4
    wide.render(scene["camera"], 64, 64, 0, 4, reset=True)
    same = np.array_equal(wide.resolve_float4()[..., :3], gpu)
    print(f"smoke: eight-wide nodes render the same image as four-wide ones: {same}")
    assert same
    wide.close()
    ctx.close()
