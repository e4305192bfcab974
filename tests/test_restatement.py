"""The independent numpy restatement (oracle/restatement.py) pinned against the unmodified reference
(oracle/_ref): physics stages a2-a6 on the reference's own grid."""
import numpy as np
import pytest

from vegasafterglow_b200 import configs


def _ref():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not present")
    return ref


@pytest.mark.parametrize("name", ["C1", "gauss_offaxis", "powerlaw_wind"])
def test_restatement_matches_reference(name):
    from oracle import restatement

    ref = _ref()
    if name == "C1":
        p, t, nu = configs.C1()
        t = t[::4]
    elif name == "gauss_offaxis":
        p, t, nu = configs.make(jet="gaussian", theta_obs=0.3), np.logspace(2.5, 7.5, 16), np.array([1e9, 1e14, 1e17])
    else:
        p = configs.make(jet="powerlaw", medium="wind", theta_obs=0.15, k_e=2.5, k_g=1.5, fwd=(0.05, 1e-2, 2.6))
        t, nu = np.logspace(2.5, 7.5, 16), np.array([1e9, 1e14, 1e17])
    d = ref.details(p, float(t[0]), float(t[-1]))
    info = d["info"]
    F = restatement.flux_density_grid(p, d["theta"], d["phi"], d["t_rows"], d["reps"], bool(info["phi_mirrored"]),
                                      int(info["n_phi_eff"]), t, nu)
    R = ref.flux_density_grid(p, t, nu)[0, 1]
    m = R > 1e-3 * R.max(axis=-1, keepdims=True)
    err = np.max(np.abs(F[m] - R[m]) / R[m])
    assert err < 1e-8, err


@pytest.mark.parametrize("name", ["C3_wind_thick", "config5_truth", "thin_ism_offaxis"])
def test_restatement_pair_matches_reference(name):
    """Forward + reverse shock (the config-5 / C3 path): pair ODE with crossing detection, shock-table
    completion, early-time extrapolation, relic-cell cooling, both spectra through one EAT geometry."""
    from oracle import restatement

    ref = _ref()
    if name == "C3_wind_thick":
        p, t, nu = configs.C3()
        t = t[::5]
    elif name == "config5_truth":
        p, t, nu = configs.make(rvs=(0.1, 1e-2, 2.5), duration=100.0), np.logspace(2.5, 6.5, 14), np.array([1e9, 4.84e14, 1e18])
    else:
        p = configs.make(E_iso=3e52, Gamma0=120.0, n_ism=0.1, theta_obs=0.05, rvs=(0.05, 3e-3, 2.2), duration=1.0)
        t, nu = np.logspace(2, 7, 14), np.array([1e9, 1e14, 1e17])
    d = ref.details(p, float(t[0]), float(t[-1]))
    info = d["info"]
    Ff, Fr = restatement.flux_density_grid(p, d["theta"], d["phi"], d["t_rows"], d["reps"], bool(info["phi_mirrored"]),
                                           int(info["n_phi_eff"]), t, nu)
    R = ref.flux_density_grid(p, t, nu)[0]
    for F, comp in ((Ff, 1), (Fr, 3)):
        b = R[comp]
        m = b > 1e-3 * b.max(axis=-1, keepdims=True)
        err = np.max(np.abs(F[m] - b[m]) / b[m])
        assert err < 1e-7, (name, comp, err)


def _full_cases():
    return {
        "C1": (configs.C1()[0], configs.C1()[1][::4], configs.C1()[2]),
        "C2": (configs.C2()[0], configs.C2()[1][::10], configs.C2()[2][::3]),
        "C3": (configs.C3()[0], configs.C3()[1][::5], configs.C3()[2]),
        "powerlaw_wind": (configs.make(jet="powerlaw", medium="wind", theta_obs=0.15, k_e=2.5, k_g=1.5),
                          np.logspace(2.5, 7.5, 16), np.array([1e9, 1e14, 1e17])),
    }


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "powerlaw_wind"])
def test_restated_grid_and_full_path_match_reference(name):
    """Row a1 restated as well (auto_grid): identical node counts / symmetry groups / mirroring, nodes to the
    quadrature-noise level, and the whole restated path (grid + physics) within the 1e-6 parity bar."""
    from oracle import restatement

    ref = _ref()
    p, t, nu = _full_cases()[name]
    g = restatement.auto_grid(p, float(t[0]), float(t[-1]))
    d = ref.details(p, float(t[0]), float(t[-1]))
    assert g["theta"].shape == d["theta"].shape and g["phi"].shape == d["phi"].shape and g["t_rows"].shape == d["t_rows"].shape
    assert np.array_equal(g["reps"], d["reps"])
    assert bool(g["phi_mirrored"]) == bool(d["info"]["phi_mirrored"]) and g["n_phi_eff"] == int(d["info"]["n_phi_eff"])
    np.testing.assert_allclose(g["theta"], d["theta"], rtol=1e-5)  # tophat grids carry ~1e-6 of quadrature noise
    np.testing.assert_allclose(g["phi"], d["phi"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["t_rows"], d["t_rows"], rtol=1e-8)
    F = restatement.model_flux(p, t, nu)
    R = ref.flux_density_grid(p, t, nu)[0]
    for f, comp in (zip(F, (1, 3)) if isinstance(F, tuple) else ((F, 1),)):
        b = R[comp]
        m = b > 1e-3 * b.max(axis=-1, keepdims=True)
        assert np.max(np.abs(f[m] - b[m]) / b[m]) < 1e-6, (name, comp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1", "C2", "C3", "powerlaw_wind"])
def test_gpu_matches_full_restatement(name):
    """GPU against the fully restated path (its own grid builder included)."""
    from oracle import restatement
    from vegasafterglow_b200.engine import Engine

    p, t, nu = _full_cases()[name]
    F = restatement.model_flux(p, t, nu)
    G = Engine(0).flux_density_grid(p, t, nu)[0]
    for f, comp in (zip(F, (1, 3)) if isinstance(F, tuple) else ((F, 1),)):
        m = f > 1e-3 * f.max(axis=-1, keepdims=True)
        assert np.max(np.abs(G[comp][m] - f[m]) / f[m]) < 2e-6, (name, comp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1", "gauss_offaxis"])
def test_gpu_matches_restatement_on_its_own_grid(name):
    """GPU flux against the numpy restatement evaluated on the GPU's OWN (phi, theta, t) grid: isolates the
    physics stages (ODE, radiation, EATS) from the quadrature noise of the grid builder -- 1e-8 for every bin."""
    from oracle import restatement
    from vegasafterglow_b200.engine import Engine

    eng = Engine(0)
    if name == "C1":
        p, t, nu = configs.C1()
        t = t[::4]
    else:
        p, t, nu = configs.make(jet="gaussian", theta_obs=0.3), np.logspace(2.5, 7.5, 16), np.array([1e9, 1e14, 1e17])
    d = eng.details(p, float(t[0]), float(t[-1]))
    info = d["info"]
    F = restatement.flux_density_grid(p, d["theta"], d["phi"], d["t_rows"], d["reps"], bool(info["phi_mirrored"]),
                                      int(info["n_phi_eff"]), t, nu)
    Gf = eng.flux_density_grid(p, t, nu)[0, 1]
    m = F > 1e-3 * F.max(axis=-1, keepdims=True)
    assert np.max(np.abs(Gf[m] - F[m]) / F[m]) < 1e-8


@pytest.mark.gpu
def test_gpu_pair_matches_restatement_on_its_own_grid():
    """Config-5 shape (FS+RS tophat): GPU against the numpy restatement on the GPU's own grid."""
    from oracle import restatement
    from vegasafterglow_b200.engine import Engine

    eng = Engine(0)
    p, t, nu = configs.make(rvs=(0.1, 1e-2, 2.5), duration=100.0), np.logspace(2.5, 6.5, 14), np.array([1e9, 4.84e14, 1e18])
    d = eng.details(p, float(t[0]), float(t[-1]))
    info = d["info"]
    Ff, Fr = restatement.flux_density_grid(p, d["theta"], d["phi"], d["t_rows"], d["reps"], bool(info["phi_mirrored"]),
                                           int(info["n_phi_eff"]), t, nu)
    G = eng.flux_density_grid(p, t, nu)[0]
    for F, comp in ((Ff, 1), (Fr, 3)):
        m = F > 1e-3 * F.max(axis=-1, keepdims=True)
        assert np.max(np.abs(G[comp][m] - F[m]) / F[m]) < 1e-7, comp
