"""Seeded synthetic inputs for the batched unit workloads (BASELINE.json configs[0])."""
import numpy as np


def uniform_hemisphere(rng, n):
    z = rng.random(n, dtype=np.float32)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z)).astype(np.float32)
    phi = (2.0 * np.pi * rng.random(n)).astype(np.float32)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)


def transmissive_tuples(n, seed=4321, combined_ggx=False):
    """Tuples for TransmissiveShading (rms = {roughness, signed cos_theta_o, specularity}) or, with combined_ggx, for the
    reflection + transmission GGX BSDF (rms = {roughness, ior_i_over_o, specularity}, wo on the whole sphere).
    wi covers both hemispheres; 1/16 of the roughnesses are 0 (perfectly specular interface)."""
    rng = np.random.default_rng(seed)
    wo = uniform_hemisphere(rng, n)
    wi = uniform_hemisphere(rng, n)
    wi[rng.random(n) < 0.5, 2] *= -1.0
    tint = rng.random((n, 3), dtype=np.float32)
    roughness = rng.random(n, dtype=np.float32)
    roughness[rng.random(n) < 1.0 / 16.0] = 0.0
    specularity = (0.01 + 0.19 * rng.random(n)).astype(np.float32)
    leaving = rng.random(n) < 0.5
    if combined_ggx:
        medium_ior = (2.0 / (1.0 - np.sqrt(specularity, dtype=np.float32)) - 1.0).astype(np.float32)
        middle = np.where(leaving, np.float32(1.0) / medium_ior, medium_ior).astype(np.float32)
        wo[leaving, 2] *= -1.0
    else:
        middle = np.where(leaving, -wo[:, 2], wo[:, 2]).astype(np.float32)
    rms = np.stack([roughness, middle, specularity], axis=1).astype(np.float32)
    u = rng.random((n, 3), dtype=np.float32)
    return {"wo": wo, "wi": wi, "tint": tint, "rms": rms, "u": u, "coat": None}


def bsdf_tuples(n, seed=1234, with_coat=False):
    """(wo, wi, tint, rms, u, coat) tuples: wo/wi uniform over the hemisphere (1/8 of wi below the horizon to
    exercise the early-outs), tint in [0,1]^3, roughness in [0,1] with 1/16 forced to 0 (delta path),
    metallic in {0, 1, U[0,1]} thirds, specularity in [0, 0.2], coat in {0 (half), U[0,1]}."""
    rng = np.random.default_rng(seed)
    wo = uniform_hemisphere(rng, n)
    wi = uniform_hemisphere(rng, n)
    below = rng.random(n) < 0.125
    wi[below, 2] *= -1.0
    tint = rng.random((n, 3), dtype=np.float32)
    roughness = rng.random(n, dtype=np.float32)
    roughness[rng.random(n) < 1.0 / 16.0] = 0.0
    third = rng.integers(0, 3, n)
    metallic = np.where(third == 0, 0.0, np.where(third == 1, 1.0, rng.random(n))).astype(np.float32)
    specularity = (0.2 * rng.random(n)).astype(np.float32)
    rms = np.stack([roughness, metallic, specularity], axis=1).astype(np.float32)
    u = rng.random((n, 3), dtype=np.float32)
    coat = None
    if with_coat:
        c = np.where(rng.random(n) < 0.5, 0.0, rng.random(n)).astype(np.float32)
        coat = np.stack([c, rng.random(n, dtype=np.float32)], axis=1).astype(np.float32)
    return {"wo": wo, "wi": wi, "tint": tint, "rms": rms, "u": u, "coat": coat}
