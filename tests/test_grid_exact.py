"""The (theta, phi) grid must equal the reference build's BIT FOR BIT.

The reference's grids are inverse CDFs of adaptive quadratures that are chaotic in the last bit
(src/core/grid-refinement.h:137-189; csrc/vag_libm.cuh header), so "close" does not exist here: either every
operation upstream of a step position is the reference build's, or the nodes move by ~1e-8 and a structured-jet
reverse shock by up to 5e-3.  csrc/vag_grid.cuh restates the reference build's instruction sequence; this file pins it.

CPU tier: the host build of the kernel source (oracle/hostemu) against the unmodified reference (oracle/_ref, built
          in this container from /root/reference) on seeded draws over every jet family and switch, and against the
          committed fixtures' stage dumps where the reference is absent.
GPU tier: the device build (through the C ABI) against the host build and against the reference.
"""
import numpy as np
import pytest

from oracle.hostemu import emu
from tests.helpers import load_golden
from vegasafterglow_b200 import configs


def _ref():
    try:
        from oracle import ref

        return ref if ref.available() else None
    except Exception:
        return None


def draws(n, seed):
    """Seeded parameter sets over every jet family / grid switch of the reference's pybind factories."""
    r = np.random.default_rng(seed)
    out = []
    for jet, med, tv, kw in (("tophat", "ism", 0.0, {}), ("tophat", "wind", 0.35, {}), ("gaussian", "ism", 0.4, {}),
                             ("gaussian", "wind", 0.0, {}), ("powerlaw", "wind", 0.3, {}), ("powerlaw", "ism", 0.0, {})):
        for rvs in (False, True):
            out.append(configs.random_draw(n, seed=int(r.integers(1 << 30)), rvs=rvs, jet=jet, medium=med, theta_obs_max=tv))
    P = np.concatenate(out)
    # Ejecta forms of the typed jets (a magnetar or sigma0 > 0 routes them through math::*_plus_one, pymodel.cpp:47-95)
    E = P[r.permutation(P.size)[: 3 * n]].copy()
    E["sigma0"][: n] = 10 ** r.uniform(-2, 1, n)
    E["has_magnetar"][n:] = 1
    E["magnetar_L0"][n:], E["magnetar_t0"][n:], E["magnetar_q"][n:] = 1e47, 1e3, 2.0
    # named Ejecta profiles (pymodel.cpp:97-146)
    named = []
    for jet in ("two_component", "step_powerlaw", "powerlaw_wing"):
        for _ in range(n):
            named.append(configs.make(jet=jet, theta_c=r.uniform(0.03, 0.12), theta_w=r.uniform(0.2, 0.5),
                                      E_iso=10 ** r.uniform(51, 54), Gamma0=10 ** r.uniform(1.8, 2.9),
                                      E_iso_w=10 ** r.uniform(49, 51), Gamma0_w=10 ** r.uniform(0.5, 1.7),
                                      k_e=r.uniform(1.5, 4), k_g=r.uniform(1.5, 4), theta_obs=r.uniform(0, 0.5),
                                      rvs=(0.1, 0.01, 2.3) if r.random() < 0.5 else None)[0])
    # grid switches: resolutions, axisymmetric=False, spreading
    S = P[r.permutation(P.size)[: 3 * n]].copy()
    S["phi_resol"][: n], S["theta_resol"][: n], S["t_resol"][: n] = 0.2, 0.6, 8
    S["axisymmetric"][n: 2 * n] = 0
    S["spreading"][n: 2 * n] = 0
    S["spreading"][2 * n:] = 1
    return np.concatenate([P, E, np.array(named, dtype=P.dtype), S])


def assert_same_grid(a, b, what):
    assert a["theta"].shape == b["theta"].shape and a["phi"].shape == b["phi"].shape, what
    np.testing.assert_array_equal(a["theta"], b["theta"], err_msg=f"{what}: theta nodes")
    np.testing.assert_array_equal(a["phi"], b["phi"], err_msg=f"{what}: phi nodes")
    assert a["t_rows"].shape == b["t_rows"].shape, what
    # the time lattice is smooth in its inputs (no quadrature): a few ulps of libm / contraction freedom
    np.testing.assert_allclose(a["t_rows"], b["t_rows"], rtol=2e-9, err_msg=f"{what}: time lattice")


def test_host_build_equals_reference_bitwise():
    ref = _ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    P = draws(3, seed=11)
    for i in range(P.size):
        p = P[i:i + 1]
        if p["axisymmetric"][0] == 0 and p["spreading"][0]:
            continue
        assert_same_grid(emu.details(p, 1e2, 1e7), ref.details(p, 1e2, 1e7), f"draw {i} (jet {int(p['jet_type'][0])})")


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_host_build_equals_committed_stage_dumps(name):
    g = load_golden("stages_" + name)
    d = emu.details(g["params"], float(g["t_min"]), float(g["t_max"]))
    np.testing.assert_array_equal(d["theta"], g["theta"])
    np.testing.assert_array_equal(d["phi"], g["phi"])


@pytest.mark.gpu
def test_device_grid_equals_host_build_and_reference_bitwise(engine):
    ref = _ref()
    P = draws(4, seed=23)
    n = 0
    for i in range(P.size):
        p = P[i:i + 1]
        if p["axisymmetric"][0] == 0 and p["spreading"][0]:
            continue
        d = engine.details(p, 1e2, 1e7)
        e = emu.details(p, 1e2, 1e7)
        np.testing.assert_array_equal(d["theta"], e["theta"], err_msg=f"draw {i}: device vs host theta")
        np.testing.assert_array_equal(d["phi"], e["phi"], err_msg=f"draw {i}: device vs host phi")
        np.testing.assert_allclose(d["t_rows"], e["t_rows"], rtol=2e-9)
        if ref is not None:
            assert_same_grid(d, ref.details(p, 1e2, 1e7), f"draw {i}: device vs reference")
        n += 1
    assert n > 60
