"""Comparison helpers shared by the parity tests."""
import numpy as np

REL_TOL = 1e-5  # BASELINE.json north_star: "BSDF/light eval, sample and PDF must agree within 1e-5 relative"


def pdf_class(p):
    """PDF wrapper classes (Types.h:155-204): 0 invalid (NaN), 1 delta dirac / MIS disabled (negative or -0),
    2 too small to be valid (|p| <= 1e-6), 3 valid."""
    p = np.asarray(p, np.float32)
    c = np.full(p.shape, 3, np.int8)
    c[np.abs(p) <= 1e-6] = 2
    c[np.signbit(p) & (np.abs(p) > 1e-6)] = 1
    c[np.isnan(p)] = 0
    return c


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    with np.errstate(invalid="ignore"):
        e = np.abs(a - b) / denom
    e[both_nan | both_inf] = 0.0
    e[np.isnan(e)] = np.inf
    return e


def report(name, err, tol):
    bad = err > tol
    return f"{name}: max rel err {np.nanmax(err):.3e}, {int(bad.sum())}/{err.size} above {tol:g}"
