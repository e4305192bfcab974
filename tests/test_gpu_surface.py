"""GPU tier: the reference-facing surfaces -- the pybind11 `Model` mirror side by side with the
reference's own pybind11 module (oracle/_ref), and the batched likelihood against the manual
ln-flux chi-squared of tests/python/test_fit_smoke.py:119-127."""
import numpy as np
import pytest

from vegasafterglow_b200 import abi, configs, fitting

pytestmark = pytest.mark.gpu


def _both():
    from oracle import ref
    from vegasafterglow_b200 import VegasAfterglowC_b200 as ours

    if not ref.available():
        pytest.skip("oracle/_ref not present")
    return ours, ref.pymodule()


def _build(va, **kw):
    return va.Model(jet=va.TophatJet(0.1, 1e52, 300), medium=va.ISM(1), observer=va.Observer(1e26, 0.1, 0),
                    fwd_rad=va.Radiation(0.1, 1e-3, 2.3), **kw)


def test_quick_start_side_by_side():
    # README.md:303-304 of the reference, run through both modules with identical calls
    ours, theirs = _both()
    t, nu = np.logspace(2, 8, 100), np.array([1e9, 1e14, 1e17])
    fo, fr = _build(ours).flux_density_grid(t, nu), _build(theirs).flux_density_grid(t, nu)
    assert fo.total.shape == np.asarray(fr.total).shape == (3, 100)
    np.testing.assert_allclose(fo.total, np.asarray(fr.total), rtol=1e-6)
    np.testing.assert_allclose(fo.fwd.sync, np.asarray(fr.fwd.sync), rtol=1e-6)
    # absent components are 0-d empty arrays in both (pymodel.h:361-383)
    for a, b in ((fo.fwd.ssc, fr.fwd.ssc), (fo.rvs.sync, fr.rvs.sync), (fo.rvs.ssc, fr.rvs.ssc)):
        assert np.asarray(a).size == np.asarray(b).size and np.asarray(a).ndim == np.asarray(b).ndim


def test_reverse_shock_wind_series_side_by_side():
    ours, theirs = _both()
    def build(va):
        return va.Model(jet=va.GaussianJet(0.1, 1e52, 300, duration=100), medium=va.Wind(0.1), observer=va.Observer(1e28, 1.0, 0.15),
                        fwd_rad=va.Radiation(0.1, 0.01, 2.3), rvs_rad=va.Radiation(0.05, 0.02, 2.6, xi_e=0.5),
                        resolutions=(0.06, 0.2, 8))
    t = np.sort(np.tile(np.logspace(2, 7, 12), 3))
    nu = np.tile([1e9, 4.84e14, 1e18], 12)
    fo, fr = build(ours).flux_density(t, nu), build(theirs).flux_density(t, nu)
    for a, b in ((fo.total, fr.total), (fo.fwd.sync, fr.fwd.sync), (fo.rvs.sync, fr.rvs.sync)):
        b = np.asarray(b)
        m = b > 1e-3 * b.max()
        # structured-jet reverse shock: reference's own reproducibility floor (tests/helpers.py)
        np.testing.assert_allclose(np.asarray(a)[m], b[m], rtol=2e-3)
    np.testing.assert_allclose(fo.total, fo.fwd.sync + fo.rvs.sync, rtol=1e-12)


def test_mirror_raises_like_the_reference():
    ours, theirs = _both()
    for va in (ours, theirs):
        m = _build(va)
        with pytest.raises(ValueError):
            m.flux_density_grid(np.array([3.0, 2.0, 1.0]), np.array([1e9]))
        with pytest.raises(ValueError):
            m.flux_density(np.array([1.0, 2.0]), np.array([1e9]))
        with pytest.raises(ValueError):
            _build(va, rtol=1.0)


def test_batched_likelihood_matches_manual_chi2(engine):
    rng = np.random.default_rng(42)
    truth = configs.make(rvs=(0.1, 0.01, 2.3), lumi_dist=1e27)
    t = np.tile(np.logspace(2.5, 6.5, 20), 5)
    nu = np.repeat([1e9, 5e9, 4.84e14, 1e17, 1e18], 20)
    order = np.argsort(t, kind="stable")
    f_true = engine.flux_density_series(truth, t[order], nu[order])[0, 0]
    flux = np.empty_like(f_true)
    flux[order] = f_true * (1 + 0.05 * rng.standard_normal(t.size))
    err = 0.1 * flux
    lk = fitting.BatchedLikelihood(engine, truth, ["E_iso", "Gamma0", "theta_c", "n_ism", "eps_e", "eps_B", "p"],
                                   [True, True, False, True, True, True, False], t, nu, flux, err,
                                   lower=[50, 1.5, 0.02, -4, -3, -5, 2.05], upper=[55, 3.2, 0.5, 2, -0.3, -0.5, 2.95])
    samples = np.column_stack([rng.uniform(51, 54, 64), rng.uniform(1.7, 3, 64), rng.uniform(0.03, 0.4, 64),
                               rng.uniform(-3, 1, 64), rng.uniform(-2, -0.5, 64), rng.uniform(-4, -1, 64),
                               rng.uniform(2.1, 2.8, 64)])
    samples[5, 0] = 60.0  # out of bounds -> -inf without evaluation (samplers.py:73-74)
    logp = lk(samples)
    assert logp[5] == -np.inf and np.isfinite(np.delete(logp, 5)).all()
    P = lk.to_params(samples)
    f = engine.flux_density_series(P, lk.t, lk.nu)[:, 0]
    manual = -0.5 * np.sum(lk.w * ((lk.lnF - np.log(np.maximum(f, 1e-300))) / lk.sig) ** 2, axis=1)
    np.testing.assert_allclose(np.delete(logp, 5), np.delete(manual, 5), rtol=1e-12)
    # the truth parameters are (close to) the best of the sample
    best = lk(np.array([[52.0, np.log10(300.0), 0.1, 0.0, -1.0, -3.0, 2.3]]))[0]
    assert best > np.max(np.delete(logp, 5))


def test_band_flux_matches_reference_fixture(engine):
    # Model.flux: Boole-rule band integration (observer.h:555-567, quadrature.h:138-191), every
    # leftover branch of the rule (num_nu = 2, 4, 5, 7, 12, 21)
    from tests.helpers import load_golden

    g = load_golden("method_flux_band")
    for key in [k for k in g if k.startswith("num_")]:
        num = int(key.split("_")[1])
        f = engine.flux_band(g["params"], g["t"], float(g["nu_min"]), float(g["nu_max"]), num)
        for ours, theirs in zip((f[0, 0], f[0, 1], f[0, 3]), g[key]):
            np.testing.assert_allclose(ours, theirs, rtol=1e-6)
    from vegasafterglow_b200 import VegasAfterglowC_b200 as va

    m = va.Model(jet=va.TophatJet(0.1, 1e52, 300, duration=1e3), medium=va.ISM(1), observer=va.Observer(1e27, 0.5, 0),
                 fwd_rad=va.Radiation(0.1, 1e-3, 2.3), rvs_rad=va.Radiation(0.1, 1e-2, 2.5))
    fb = m.flux(g["t"], float(g["nu_min"]), float(g["nu_max"]), 12)
    np.testing.assert_allclose(fb.total, g["num_12"][0], rtol=1e-6)
    with pytest.raises(ValueError):
        m.flux(g["t"], 1e17, 1e16, 5)
    with pytest.raises(ValueError):
        m.flux(g["t"], 1e16, 1e17, 1)


def test_exposure_averaging_matches_reference_fixture(engine):
    from tests.helpers import load_golden

    g = load_golden("method_exposures")
    f = engine.flux_density_exposures(g["params"], g["t"], g["nu"], g["expo"], int(g["num_points"]))
    for ours, theirs in zip((f[0, 0], f[0, 1], f[0, 3]), g["flux"]):
        np.testing.assert_allclose(ours, theirs, rtol=1e-6)
    from vegasafterglow_b200 import VegasAfterglowC_b200 as va

    m = va.Model(jet=va.TophatJet(0.1, 1e52, 300, duration=1e3), medium=va.ISM(1), observer=va.Observer(1e27, 0.5, 0),
                 fwd_rad=va.Radiation(0.1, 1e-3, 2.3), rvs_rad=va.Radiation(0.1, 1e-2, 2.5))
    fe = m.flux_density_exposures(g["t"], g["nu"], g["expo"], 10)
    np.testing.assert_allclose(fe.total, g["flux"][0], rtol=1e-6)
    with pytest.raises(ValueError):
        m.flux_density_exposures(g["t"], g["nu"], -g["expo"], 10)


def test_present_only_output_mode_matches_dense():
    """VAG_OUT_PRESENT (include/vag.h): planes of components no model of the batch has are not written,
    every other plane equals the dense transfer (to 1e-13: a batch this small splits a model's rows over
    several CTAs that combine with atomicAdd, so two runs differ in the last bits)."""
    from vegasafterglow_b200.engine import Engine

    eng = Engine(0)
    t, nu = np.logspace(2, 8, 40), np.array([1e9, 1e14, 1e17])
    for rvs in (False, True):
        P = configs.random_draw(8, seed=77, rvs=rvs)
        dense = eng.flux_density_grid(P, t, nu)
        eng.set_output_mode(True)
        try:
            sparse = eng.flux_density_grid(P, t, nu)
            sd = eng.flux_density_series(P, np.sort(np.tile(t, 3)), np.tile(nu, 40))
        finally:
            eng.set_output_mode(False)
        dd = eng.flux_density_series(P, np.sort(np.tile(t, 3)), np.tile(nu, 40))
        present = [abi.COMPONENTS.index("total"), abi.COMPONENTS.index("fwd_sync")] + ([abi.COMPONENTS.index("rvs_sync")] if rvs else [])
        for c in range(abi.NCOMP):
            if c in present:
                np.testing.assert_allclose(sparse[:, c], dense[:, c], rtol=1e-13)
                np.testing.assert_allclose(sd[:, c], dd[:, c], rtol=1e-13)
            else:
                assert not sparse[:, c].any() and not dense[:, c].any()


def test_magnetar_side_by_side():
    """TophatJet(..., magnetar=Magnetar(L0, t0, q)) through both pybind11 modules (pybind/pybind.cpp:198-209)."""
    ours, theirs = _both()
    t, nu = np.logspace(2, 7, 40), np.array([1e9, 1e14, 1e17])
    out = []
    for va in (ours, theirs):
        m = va.Model(jet=va.TophatJet(0.1, 1e52, 300, magnetar=va.Magnetar(1e48, 1e3, 2.0)), medium=va.ISM(1),
                     observer=va.Observer(1e26, 0.1, 0), fwd_rad=va.Radiation(0.1, 1e-3, 2.3))
        out.append(np.asarray(m.flux_density_grid(t, nu).total))
    np.testing.assert_allclose(out[0], out[1], rtol=1e-6)
    plain = np.asarray(_build(ours).flux_density_grid(t, nu).total)
    assert np.max(np.abs(out[0] / plain - 1)) > 0.05  # the injection visibly re-brightens the afterglow
    for va in (ours, theirs):
        with pytest.raises(ValueError):
            va.Magnetar(-1.0, 1e3, 2.0)


def test_spreading_side_by_side():
    """GaussianJet(..., spreading=True) off-axis through both pybind11 modules: theta(k) dynamics
    (forward-shock.tpp:36-40) and the per-node EATS geometry (observer.cpp:51-141)."""
    ours, theirs = _both()
    t, nu = np.logspace(2, 8, 50), np.array([1e9, 1e14, 1e17])
    out = []
    for va in (ours, theirs):
        m = va.Model(jet=va.GaussianJet(0.1, 1e52, 300, spreading=True), medium=va.ISM(1),
                     observer=va.Observer(1e26, 0.1, 0.3), fwd_rad=va.Radiation(0.1, 1e-3, 2.3))
        out.append(np.asarray(m.flux_density_grid(t, nu).total))
    np.testing.assert_allclose(out[0], out[1], rtol=1e-6)
    m0 = ours.Model(jet=ours.GaussianJet(0.1, 1e52, 300), medium=ours.ISM(1), observer=ours.Observer(1e26, 0.1, 0.3),
                    fwd_rad=ours.Radiation(0.1, 1e-3, 2.3))
    assert np.max(np.abs(out[0] / np.asarray(m0.flux_density_grid(t, nu).total) - 1)) > 0.5


def test_details_side_by_side():
    """Model.details(t_min, t_max): SimulationDetails of both modules (pybind/pymodel.cpp:315-348) for a tophat,
    an off-axis Gaussian with a reverse shock, a spreading jet and two inverse-Compton-cooled models."""
    ours, theirs = _both()

    def models(va):
        obs_on, obs_off = va.Observer(1e26, 0.1, 0), va.Observer(1e27, 0.3, 0.25)
        rad = va.Radiation(0.1, 1e-3, 2.3)
        return [
            va.Model(jet=va.TophatJet(0.1, 1e52, 300), medium=va.ISM(1), observer=obs_on, fwd_rad=rad),
            va.Model(jet=va.GaussianJet(0.1, 1e52, 300, duration=100), medium=va.Wind(0.1), observer=obs_off, fwd_rad=rad,
                     rvs_rad=va.Radiation(0.1, 1e-2, 2.5)),
            va.Model(jet=va.TophatJet(0.15, 1e52, 200, spreading=True), medium=va.ISM(0.1), observer=obs_off, fwd_rad=rad),
            # spreading with axisymmetric=False: one ODE row per (phi, theta) cell
            va.Model(jet=va.TophatJet(0.15, 1e52, 200, spreading=True), medium=va.ISM(0.1), observer=obs_off, fwd_rad=rad,
                     axisymmetric=False),
            # inverse-Compton cooling (Thomson and Klein-Nishina): electrons after cooling + the InverseComptonY record
            va.Model(jet=va.TophatJet(0.1, 1e53, 300), medium=va.ISM(1), observer=obs_on,
                     fwd_rad=va.Radiation(0.1, 1e-4, 2.3, ssc=True)),
            va.Model(jet=va.TophatJet(0.1, 1e53, 300, duration=50), medium=va.ISM(1), observer=obs_on,
                     fwd_rad=va.Radiation(0.1, 1e-4, 2.3, ssc=True, kn=True), rvs_rad=va.Radiation(0.1, 1e-3, 2.4, ssc=True, kn=True)),
        ]

    for mo, mr in zip(models(ours), models(theirs)):
        do, dr = mo.details(1e2, 1e7), mr.details(1e2, 1e7)
        np.testing.assert_allclose(do.theta, np.asarray(dr.theta), rtol=2e-4)
        np.testing.assert_allclose(do.phi, np.asarray(dr.phi), rtol=1e-6)
        np.testing.assert_allclose(do.t_src, np.asarray(dr.t_src), rtol=5e-3)
        shocks = [(do.fwd, dr.fwd)] + ([(do.rvs, dr.rvs)] if np.asarray(dr.rvs.Gamma).ndim == 3 else [])
        assert np.asarray(do.rvs.Gamma).ndim == np.asarray(dr.rvs.Gamma).ndim
        for so, sr in shocks:
            for name in ("t_comv", "r", "theta", "Gamma", "Gamma_th", "B_comv", "N_p", "t_obs", "Doppler", "nu_m", "nu_c",
                         "nu_a", "nu_M", "I_nu_max", "gamma_m", "gamma_c", "gamma_a", "gamma_M", "N_e", "gamma_m_hat",
                         "gamma_c_hat", "nu_m_hat", "nu_c_hat", "Y_T"):
                a, b = np.asarray(getattr(so, name)), np.asarray(getattr(sr, name))
                assert a.shape == b.shape, (name, a.shape, b.shape)
                ok = np.isfinite(b) & (np.abs(b) > 0) & np.isfinite(a)
                if not ok.any():  # e.g. Y_T without inverse-Compton cooling: identically zero in both
                    assert np.all(a[np.isfinite(b)] == b[np.isfinite(b)]), name
                    continue
                # stage tables follow the theta grid / time lattice, which carry the quadrature noise of the
                # grid builder: 1e-2 covers the noisiest early-time nodes
                med = np.median(np.abs(a[ok] - b[ok]) / np.abs(b[ok]))
                assert med < 1e-6, (name, med)
                np.testing.assert_allclose(a[ok], b[ok], rtol=2e-2, err_msg=name)
