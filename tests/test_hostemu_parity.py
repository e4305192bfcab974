"""CPU tier: the kernel bodies of vegasafterglow_b200/csrc executed sequentially on the host
(oracle/hostemu) against the committed reference fixtures -- stage by stage (grid, dynamics) and end
to end (flux).  This is the same code the GPU runs; the -m gpu tier repeats the flux checks through
the C ABI on the device."""
import numpy as np
import pytest

from tests.helpers import CHAOTIC, assert_parity, assert_reference_contract, golden_names, load_golden
from oracle.hostemu import emu
from vegasafterglow_b200 import configs


def _rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_stage_tables(name):
    g = load_golden("stages_" + name)
    d = emu.details(g["params"], float(g["t_min"]), float(g["t_max"]))
    n_phi, n_theta, n_t, n_reps, sym, mirrored, n_phi_eff, _ = g["info"]
    i = d["info"]
    # grid sizes and symmetry class are integers: exact
    assert (i["n_phi"], i["n_theta"], i["n_t"], i["n_reps"], i["symmetry"], i["phi_mirrored"], i["n_phi_eff"]) == \
           (n_phi, n_theta, n_t, n_reps, sym, mirrored, n_phi_eff)
    np.testing.assert_array_equal(d["reps"], g["reps"])
    # theta / phi nodes: the reference build's arithmetic, bit for bit (csrc/vag_grid.cuh, tests/test_grid_exact.py)
    np.testing.assert_array_equal(d["theta"], g["theta"])
    np.testing.assert_array_equal(d["phi"], g["phi"])
    assert _rel(d["t_rows"], g["t_rows"]) < 1e-12
    # shock tables of the representative rows [t_comv, r, theta, Gamma, Gamma_th, B, N_p]
    for a, nm in enumerate(("t_comv", "r", "theta", "Gamma", "Gamma_th", "B", "N_p")):
        assert _rel(d["fwd_shock"][a], g["fwd_shock"][a]) < 5e-6, nm
    if g["params"]["has_rvs"][0]:
        np.testing.assert_array_equal(d["inj_idx"], g["inj_idx"])
        for a, nm in enumerate(("t_comv", "r", "theta", "Gamma", "Gamma_th", "B", "N_p")):
            assert _rel(d["rvs_shock"][a], g["rvs_shock"][a]) < 5e-6, nm


@pytest.mark.parametrize("name", [n for n in golden_names() if n not in ("config_C2_dense", "config_C4")])
def test_flux_parity(name):
    g = load_golden(name)
    fn = emu.flux_density_series if bool(g["series"]) else emu.flux_density_grid
    f, st = fn(g["params"], g["t"], g["nu"])
    assert (st == 0).all()
    if name in CHAOTIC:
        assert_reference_contract(f, g, name)
    else:
        assert_parity(f, g, name)


def test_series_equals_grid():
    # tests/python/test_physics_invariants.py:80-85 (series == grid to 1e-12)
    p, t, nu = configs.C3()
    fg, _ = emu.flux_density_grid(p, t, nu)
    ts, nus = np.repeat(t, nu.size), np.tile(nu, t.size)
    fs, _ = emu.flux_density_series(p, ts, nus)
    np.testing.assert_allclose(fs[0, 0].reshape(t.size, nu.size).T, fg[0, 0], rtol=1e-12)


def test_banded_series_equals_per_point_series():
    # multi-band light curve (5 bands x 12 epochs, each band on its own epochs): the banded evaluation (boundary
    # luminosities staged per band) and the per-point one evaluate the same spectra at the same nodes
    p, _, _ = configs.C3()
    nus = np.tile([1e9, 5e9, 4.84e14, 1e17, 1e18], 12)
    ts = np.sort(np.logspace(2.3, 6.8, 60) * (1 + 0.01 * np.arange(60) % 0.05))
    f1, st1 = emu.flux_density_series(p, ts, nus, series_mode=1)
    f2, st2 = emu.flux_density_series(p, ts, nus, series_mode=2)
    assert (st1 == 0).all() and (st2 == 0).all()
    assert (f1[0, 0] > 0).all()
    np.testing.assert_allclose(f2, f1, rtol=1e-12)
    # nine distinct frequencies: not a banded request, mode 2 falls back to per-point spectra (identical bits)
    nus9 = np.tile(np.logspace(9, 18, 9), 7)[:60]
    g1, _ = emu.flux_density_series(p, ts, nus9, series_mode=1)
    g2, _ = emu.flux_density_series(p, ts, nus9, series_mode=2)
    np.testing.assert_array_equal(g1, g2)


def test_distance_scaling_and_total():
    # F ~ 1/d_L^2 to 1e-9 and total == sum of parts to 1e-12 (test_physics_invariants.py:37-60)
    p, t, nu = configs.C3()
    f1, _ = emu.flux_density_grid(p, t, nu)
    q = p.copy()
    q["lumi_dist"] *= 3.0
    f2, _ = emu.flux_density_grid(q, t, nu)
    np.testing.assert_allclose(f2[0, 0] * 9.0, f1[0, 0], rtol=1e-9)
    np.testing.assert_allclose(f1[0, 0], f1[0, 1] + f1[0, 3], rtol=1e-12)
