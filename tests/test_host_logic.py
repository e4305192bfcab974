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
