"""Shared helpers of the test-suite (fixtures live in conftest.py)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix)
                  and not f.startswith("stages_"))


def model_errors(a, b, comp=0, floor=1e-2):
    """Per-model max relative error of component `comp` over bins above `floor` x the band peak."""
    out = np.zeros(a.shape[0])
    for i in range(a.shape[0]):
        bb, aa = b[i, comp], a[i, comp]
        if not np.any(bb > 0):
            out[i] = float(np.max(np.abs(aa)))
            continue
        peak = bb.max(axis=-1, keepdims=True) if bb.ndim == 2 else bb.max()
        m = bb > floor * peak
        out[i] = float(np.max(np.abs(aa[m] - bb[m]) / bb[m]))
    return out


# Parity bar (BASELINE.json north_star): max relative flux error <= 1e-6 in FP64.  The reference's
# adaptive theta-grid (a dopri5 CDF integration whose step-size controller is driven by rounding
# noise, src/core/grid-refinement.h:137-189) makes the reference ITSELF reproducible only to
# ~1e-8..1e-5 between two builds of the same source; fixtures carry both builds, and a model
# passes when it is within max(1e-6, SPREAD_FACTOR x that model's reference-vs-reference spread);
# the factor allows for the spread being a single sample of the reference's noise (structured-jet
# reverse shocks are chaotic in the reference itself: tests/python/test_golden.py:95).
FLUX_RTOL = 1e-6
SPREAD_FACTOR = 4.0


def assert_parity(flux, g, what):
    for comp in (0, 1, 3):
        err = model_errors(flux, g["flux"], comp)
        floor = model_errors(g["flux_alt"], g["flux"], comp)
        tol = np.maximum(FLUX_RTOL, SPREAD_FACTOR * floor)
        bad = np.nonzero(err > tol)[0]
        assert bad.size == 0, (f"{what} comp {comp}: models {bad.tolist()} exceed tolerance: err={err[bad]}, "
                               f"tol={tol[bad]}")
    return model_errors(flux, g["flux"], 0)


