"""CPU tier: host-side logic -- pybind11 mirror surface, likelihood glue, walker partition under
gloo (world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from vegasafterglow_b200 import abi, configs, fitting, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _va():
    from vegasafterglow_b200 import VegasAfterglowC_b200 as va

    return va


def test_pybind_mirror_has_the_reference_surface():
    va = _va()
    for name in ("Model", "TophatJet", "GaussianJet", "PowerLawJet", "TwoComponentJet", "StepPowerLawJet", "PowerLawWing",
                 "ISM", "Wind", "Observer", "Radiation", "FluxDict"):
        assert hasattr(va, name), name
    m = va.Model(jet=va.PowerLawJet(0.1, 1e52, 300, 2, 2, duration=10), medium=va.Wind(0.1, n0=1e3),
                 observer=va.Observer(1e28, 1.0, 0.3), fwd_rad=va.Radiation(0.1, 0.01, 2.3),
                 rvs_rad=va.Radiation(0.1, 0.01, 2.5, xi_e=0.5), resolutions=(0.1, 0.3, 8), rtol=1e-7,
                 radiative_fireball=False)
    p = np.frombuffer(m.params_bytes, dtype=abi.PARAMS_DTYPE)[0]
    ref = configs.make(jet="powerlaw", duration=10, medium="wind", A_star=0.1, n0=1e3, lumi_dist=1e28, z=1.0,
                       theta_obs=0.3, fwd=(0.1, 0.01, 2.3), rvs=(0.1, 0.01, 2.5), rvs_xi_e=0.5,
                       resolutions=(0.1, 0.3, 8), rtol=1e-7, radiative_fireball=False)[0]
    for k in abi.PARAMS_DTYPE.names:
        if k.startswith("pad"):
            continue
        assert np.all(p[k] == ref[k]) or (np.isinf(p[k]) and np.isinf(ref[k])), k
    assert m.rtol == 1e-7 and m.axisymmetric and not m.radiative_fireball


def test_pybind_mirror_error_conventions():
    va = _va()
    obs, rad = va.Observer(1e26, 0.1, 0), va.Radiation(0.1, 1e-3, 2.3)
    with pytest.raises(ValueError):
        va.TophatJet(0.0, 1e52, 300)
    with pytest.raises(ValueError):
        va.TophatJet(0.1, 1e52, 1.0)
    with pytest.raises(ValueError):
        va.Radiation(1.5, 1e-3, 2.3)
    with pytest.raises(ValueError):
        va.Observer(1e26, -1, 0)
    with pytest.raises(ValueError):
        va.Model(va.TophatJet(0.1, 1e52, 300), va.ISM(1), obs, rad, rtol=1.0)
    with pytest.raises(TypeError):
        va.Model(jet="tophat", medium=va.ISM(1), observer=obs, fwd_rad=rad)
    with pytest.raises(TypeError):
        va.Model(jet=va.TophatJet(0.1, 1e52, 300), medium=3, observer=obs, fwd_rad=rad)
    va.Wind(0.1, k_m=1.5)  # general wind slope: the reference's generic-Medium path, on the GPU path too
    with pytest.raises(ValueError, match="k_m"):
        va.Wind(0.1, k_m=-1.0)
    with pytest.raises(ValueError, match="theta_w"):
        va.TwoComponentJet(0.1, 1e52, 300, 0.05, 1e50, 50)
    j = va.TwoComponentJet(0.05, 1e52, 300, 0.3, 1e50, 50)
    mm = va.Model(j, va.ISM(1), obs, rad)
    pp = np.frombuffer(mm.params_bytes, dtype=abi.PARAMS_DTYPE)[0]
    assert pp["jet_type"] == abi.JET_TWO_COMPONENT and pp["theta_w"] == 0.3 and pp["Gamma0_w"] == 50
    m = va.Model(va.TophatJet(0.1, 1e52, 300), va.ISM(1), obs, rad)
    # Model.resolutions (pybind/pybind.cpp:453): the defaults in effect, or what was passed
    assert m.resolutions == (0.06, 0.15, 6.0)
    assert va.Model(va.TophatJet(0.1, 1e52, 300), va.ISM(1), obs, rad, rvs_rad=rad).resolutions == (0.06, 0.2, 10.0)
    assert va.Model(va.TophatJet(0.1, 1e52, 300), va.ISM(1), obs, rad, resolutions=(0.1, 0.3, 8)).resolutions == (0.1, 0.3, 8.0)
    with pytest.raises(ValueError, match="same size"):
        m.flux_density(np.array([1.0, 2.0]), np.array([1e9]))
    with pytest.raises(ValueError, match="non-empty"):
        m.flux_density_grid(np.array([]), np.array([1e9]))


def test_partition_covers_everything():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [parallel.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) <= -(-n // world) if n else True


def test_cost_balanced_assignment_of_a_mixed_ensemble():
    # a tophat walker is one ODE row and one phi row, a structured off-axis walker ~50 x ~15 of
    # them.  The cost proxy must see that, and the serpentine deal must level it: max / min per-rank cost <= 1.15.
    P = np.concatenate([configs.random_draw(3072, seed=1, rvs=True),
                        configs.random_draw(1024, seed=2, rvs=True, jet="gaussian", theta_obs_max=0.4)])
    P = P[np.random.default_rng(0).permutation(P.size)]
    cost = parallel.cost_proxy(P, 1e2, 1e7)
    g = P["jet_type"] == abi.JET_GAUSSIAN
    assert np.median(cost[g]) > 10 * np.median(cost[~g])
    for world in (2, 4, 8):
        sets = parallel.balanced_assignment(cost, world)
        assert sorted(np.concatenate(sets).tolist()) == list(range(P.size))
        per = np.array([cost[s].sum() for s in sets])
        assert per.max() / per.min() <= 1.15, (world, per)
        contiguous = np.array([cost[slice(*parallel.partition(P.size, world, r))].sum() for r in range(world)])
        assert per.max() <= contiguous.max()


def test_invalid_walkers_are_masked_before_the_split():
    # a walker the model constructor rejects (theta_w <= theta_c is routine in MCMC) must not change what any
    # rank sends into the collective.  The mask is pure host code on the full ensemble.
    from vegasafterglow_b200.engine import Engine

    P = configs.random_draw(9, seed=4)
    P["theta_c"][3] = -0.1
    P["fwd"]["eps_e"][7] = 2.0
    ok = Engine.valid_mask(P)
    assert ok.tolist() == [True, True, True, False, True, True, True, False, True]
    seen = []

    def evaluate(block):
        seen.append(len(block))
        return np.arange(len(block), dtype=float)

    chi2 = parallel.partitioned_chi2(None, P, np.array([1e3, 1e5]), None, None, None, None, evaluate=evaluate, valid=ok)
    assert seen == [7] and np.isinf(chi2[[3, 7]]).all() and np.isfinite(np.delete(chi2, [3, 7])).all()


def test_band_observation_record():
    b = fitting.band_obs(t=[2e4, 1e3], flux=[2e-12, 1e-12], err=[2e-13, 2e-13], nu_min=7.25e16, nu_max=2.4e18, num_points=9,
                         weights=[2.0, 1.0])
    np.testing.assert_array_equal(b["t"], [1e3, 2e4])
    np.testing.assert_allclose(b["lnF_obs"], np.log([1e-12, 2e-12]))
    np.testing.assert_allclose(b["sigma_ln"], [0.2, 0.1])
    np.testing.assert_array_equal(b["w"], [1.0, 2.0])  # band weights are NOT normalised (fitter.py:362-373)
    assert (b["nu_min"], b["nu_max"], b["num_nu"]) == (7.25e16, 2.4e18, 9)
    with pytest.raises(ValueError, match="positive"):
        fitting.band_obs([1.0], [0.0], [1.0], 1e9, 1e10)


def test_likelihood_parameter_mapping():
    lk = fitting.BatchedLikelihood(None, configs.make(rvs=(0.1, 0.01, 2.3)), ["E_iso", "theta_c", "p", "eps_B_r", "theta_v"],
                                   [True, False, False, True, False], [10.0, 5.0], [1e9, 1e14], [1.0, 2.0], [0.1, 0.2],
                                   weights=[1.0, 3.0])
    P = lk.to_params(np.array([[52.5, 0.2, 2.4, -3.0, 0.1]]))
    assert P["E_iso"][0] == 10 ** 52.5 and P["theta_c"][0] == 0.2 and P["fwd"]["p"][0] == 2.4
    assert P["rvs"]["eps_B"][0] == 1e-3 and P["theta_obs"][0] == 0.1 and P["has_rvs"][0] == 1
    # data consolidation (fitter.py:407-437): sorted by time, weights normalised to N
    np.testing.assert_array_equal(lk.t, [5.0, 10.0])
    np.testing.assert_allclose(lk.w, [1.5, 0.5])
    np.testing.assert_allclose(lk.lnF, np.log([2.0, 1.0]))
    np.testing.assert_allclose(lk.sig, [0.1, 0.1])


_GLOO_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch.distributed as dist
    from vegasafterglow_b200 import configs, parallel
    from oracle.hostemu import emu
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    P = configs.random_draw(7, seed=3, rvs=True)
    ts = np.sort(np.tile(np.logspace(3, 6, 4), 2)); nus = np.tile([1e9, 1e17], 4)
    lnF = np.log(np.full(ts.size, 1e-27)); sig = np.full(ts.size, 0.1); w = np.ones(ts.size)
    def host_eval(block):  # test-only evaluator: the kernel bodies executed on the host
        f, _ = emu.flux_density_series(block, ts, nus)
        d = (lnF - np.log(np.maximum(f[:, 0], 1e-300))) / sig
        return np.sum(w * d * d, axis=1)
    chi2 = parallel.partitioned_chi2(None, P, ts, nus, lnF, sig, w, evaluate=host_eval)
    full = host_eval(P)
    assert chi2.shape == (7,) and np.array_equal(chi2, full), (chi2, full)
    chi2c = parallel.partitioned_chi2(None, P, ts, nus, lnF, sig, w, evaluate=host_eval, balance=False)
    assert np.array_equal(chi2c, full)
    # a rejected walker on ONE rank's share: masked on every rank before the split, one collective of one shape
    from vegasafterglow_b200.engine import Engine
    Q = P.copy(); Q["theta_c"][5] = -1.0
    ok = Engine.valid_mask(Q)
    chi2m = parallel.partitioned_chi2(None, Q, ts, nus, lnF, sig, w, evaluate=host_eval, valid=ok)
    assert np.isinf(chi2m[5]) and np.array_equal(np.delete(chi2m, 5), np.delete(full, 5))
    lo, hi = parallel.partition(7, 2, dist.get_rank())
    assert (lo, hi) == ((0, 4) if dist.get_rank() == 0 else (4, 7))
    dist.destroy_process_group()
    print("ok", dist.get_rank() if False else sys.argv[1])
""")


def test_walker_partition_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=29517 + os.getpid() % 1000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_softplus_table_accuracy():
    """log2(1 + 2^x) table of the EATS kernel (vag_math.cuh): absolute error against long double over [-20, 20]."""
    import ctypes

    from oracle.hostemu import emu

    lib = emu.lib()
    lib.vagemu_softplus_lut_maxerr.restype = ctypes.c_double
    assert lib.vagemu_softplus_lut_maxerr(400001) < 5e-13


def test_params_layout_upgrade():
    """Fixtures written with an earlier, shorter vag_params layout load by field name; new fields keep defaults."""
    from vegasafterglow_b200 import abi, configs

    cur = configs.make(jet="gaussian", theta_obs=0.3, rvs=(0.1, 0.01, 2.4), duration=50.0)
    old_dtype = np.dtype({n: abi.PARAMS_DTYPE.fields[n] for n in abi.PARAMS_DTYPE.names[:-6]})  # before the magnetar / k_m fields
    old = np.zeros(1, dtype=old_dtype)
    for n in old_dtype.names:
        old[n] = cur[n]
    up = abi.upgrade_params(old)
    assert up.dtype == abi.PARAMS_DTYPE and up["wind_k_m"][0] == 2.0 and up["has_magnetar"][0] == 0
    for n in old_dtype.names:
        assert np.array_equal(up[n], cur[n]), n
    assert abi.upgrade_params(cur) is cur
