"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) -- run in the CPU container where /root/reference exists:

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs) and tests/golden/PINNING.json, which
records how closely the locally built reference reproduces the reference's OWN committed golden
vectors (/root/reference/tests/python/golden/*.npz) for the in-scope configurations.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from vegasafterglow_b200 import abi, configs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_GOLDEN = "/root/reference/tests/python/golden"


ONLY = tuple(sys.argv[1:])  # optional fixture-name prefixes: regenerate just those, keep the other files


def save(name, params, t, nu, series=False, **extra):
    path = os.path.join(OUT, name + ".npz")
    if ONLY and not name.startswith(ONLY) and os.path.exists(path):
        return np.load(path)["flux"]
    fn = ref.flux_density_series if series else ref.flux_density_grid
    f = fn(params, t, nu, n_threads=8)
    # the same unmodified reference built with different code generation (oracle/Makefile,
    # libvagref_alt.so): |flux - flux_alt| is the reference's own reproducibility floor
    with ref.use_variant("alt"):
        f_alt = fn(params, t, nu, n_threads=8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), params=params, t=t, nu=nu, flux=f, flux_alt=f_alt,
                        series=series, **extra)
    return f


def main():
    pin = {}
    # 1. the reference's own golden configurations that are in scope (typed jets, no SSC)
    for name in configs.GOLDEN:
        p = configs.golden(name)
        f = save("golden_" + name, p, configs.GOLDEN_T, configs.GOLDEN_NU)
        path = os.path.join(REF_GOLDEN, name + ".npz")
        if os.path.exists(path):
            g = np.load(path)
            dev = {}
            for ci, cname in enumerate(abi.COMPONENTS):
                b = np.asarray(g[cname])
                if b.ndim != 2 or not b.any():
                    continue
                a = f[0, ci]
                m = b > 1e-2 * b.max()
                dev[cname] = float(np.max(np.abs(a[m] - b[m]) / b[m]))
            pin[name] = dev
    # 2. BASELINE.json configs C1-C3 (+ dense C2)
    for name, (p, t, nu) in {"C1": configs.C1(), "C2": configs.C2(), "C2_dense": configs.C2(dense=True),
                             "C3": configs.C3(), "C4": configs.C4()}.items():
        save("config_" + name, p, t, nu)
    # 3. stage tables of C1, C2, C3 (Coord + Shock + observer grids)
    for name, (p, t, nu) in ({} if ONLY else {"C1": configs.C1(), "C2": configs.C2(), "C3": configs.C3()}).items():
        d = ref.details(p, t[0], t[-1])
        np.savez_compressed(os.path.join(OUT, "stages_" + name + ".npz"), params=p, t_min=t[0], t_max=t[-1],
                            info=np.array(tuple(d["info"].tolist())), **{k: v for k, v in d.items() if k != "info"})
    # 4. seeded random batches (SURVEY.md section 8d draw), grid and series (config C5 data layout)
    t, nu = configs.C1()[1:]
    save("batch_fs_tophat_ism", configs.random_draw(48, seed=11), t, nu)
    save("batch_rs_tophat_ism", configs.random_draw(48, seed=12, rvs=True), t, nu)
    save("batch_fs_gauss_offaxis", configs.random_draw(12, seed=13, jet="gaussian", theta_obs_max=0.4), t, nu)
    save("batch_fs_powerlaw_wind", configs.random_draw(12, seed=14, jet="powerlaw", medium="wind", theta_obs_max=0.3), t, nu)
    save("batch_rs_tophat_wind", configs.random_draw(24, seed=15, rvs=True, medium="wind"), t, nu)
    P2 = configs.random_draw(16, seed=31, jet="two_component", theta_obs_max=0.3)
    P2["theta_c"] = np.random.default_rng(31).uniform(0.03, 0.1, 16)
    P2["theta_w"] = P2["theta_c"] * np.random.default_rng(32).uniform(2, 5, 16)
    P2["E_iso_w"] = P2["E_iso"] * 10 ** np.random.default_rng(33).uniform(-3, -1, 16)
    P2["Gamma0_w"] = np.maximum(P2["Gamma0"] * 0.2, 5.0)
    save("batch_fs_two_component", P2, t, nu)
    P3 = configs.random_draw(8, seed=34, jet="step_powerlaw", rvs=True, theta_obs_max=0.2)
    P3["E_iso_w"], P3["Gamma0_w"], P3["k_e"], P3["k_g"] = P3["E_iso"] * 0.3, np.maximum(P3["Gamma0"] * 0.5, 5.0), 3.0, 1.5
    save("batch_rs_step_powerlaw", P3, t, nu)
    # PowerLawWing (pybind/pymodel.cpp:131-146, math::powerlaw_wing jet.h:403-416): a hollow core, the wing carries
    # E_iso_w / Gamma0_w; forward shock and forward + reverse shock, on and off axis
    rw = np.random.default_rng(61)
    PW = configs.random_draw(16, seed=60, jet="powerlaw_wing", theta_obs_max=0.4)
    PW["has_rvs"][8:] = 1
    PW["rvs"][8:] = PW["fwd"][8:]
    PW["duration"][8:] = 10 ** rw.uniform(0, 3, 8)
    PW["theta_c"] = rw.uniform(0.03, 0.15, 16)
    PW["E_iso_w"], PW["Gamma0_w"] = 10 ** rw.uniform(50, 53, 16), 10 ** rw.uniform(1.0, 2.5, 16)
    PW["k_e"], PW["k_g"] = rw.uniform(1.5, 4.5, 16), rw.uniform(1.0, 3.0, 16)
    save("batch_mixed_powerlaw_wing", PW, t, nu)
    P4 = configs.random_draw(16, seed=35, rvs=True)
    P4["sigma0"] = 10 ** np.random.default_rng(36).uniform(-2, 1, 16)
    save("batch_rs_magnetized_tophat", P4, t, nu)
    # narrow structured cores: the steep Gaussian trips find_jet_jumps' discontinuity test at up to 14 scan
    # intervals (grid-refinement.h:40-86), each of which adds refinement nodes
    PN = configs.random_draw(12, seed=54, jet="gaussian", theta_obs_max=0.1)
    PN["theta_c"] = 10 ** np.random.default_rng(55).uniform(-2.4, -1.5, PN.size)
    save("batch_fs_narrow_gauss", PN, t, nu)
    # magnetar energy injection (jet factories' magnetar=Magnetar(L0, t0, q); the jet takes the Ejecta path)
    rng = np.random.default_rng(41)
    for nm, P5 in (("batch_fs_magnetar_tophat", configs.random_draw(16, seed=37)),
                   ("batch_rs_magnetar_tophat", configs.random_draw(12, seed=38, rvs=True)),
                   ("batch_fs_magnetar_gauss_offaxis", configs.random_draw(8, seed=39, jet="gaussian", theta_obs_max=0.3))):
        P5["has_magnetar"] = 1
        P5["magnetar_L0"] = 10 ** rng.uniform(46, 49.5, P5.size)
        P5["magnetar_t0"] = 10 ** rng.uniform(2, 4.5, P5.size)
        P5["magnetar_q"] = rng.uniform(1.0, 3.0, P5.size)
        save(nm, P5, t, nu)
    # lateral spreading (spreading=True of the jet factories): theta(k) dynamics for forward-shock models,
    # Symmetry::structured lattices for both
    for nm, P6 in (("batch_fs_spreading_tophat", configs.random_draw(16, seed=42, theta_obs_max=0.3)),
                   ("batch_fs_spreading_gauss", configs.random_draw(8, seed=43, jet="gaussian", theta_obs_max=0.4)),
                   ("batch_fs_spreading_powerlaw_wind", configs.random_draw(8, seed=44, jet="powerlaw", medium="wind", theta_obs_max=0.2)),
                   ("batch_rs_spreading_tophat", configs.random_draw(8, seed=45, rvs=True, theta_obs_max=0.2))):
        P6["spreading"] = 1
        save(nm, P6, t, nu)
    # Model(axisymmetric=False): unmirrored phi grid, every phi row observed (jet_3d = 1)
    for nm, P7 in (("batch_fs_nonaxisym_mixed", np.concatenate([configs.random_draw(6, seed=46, theta_obs_max=0.3),
                                                                configs.random_draw(5, seed=47, jet="gaussian", theta_obs_max=0.4),
                                                                configs.random_draw(5, seed=48, jet="powerlaw", medium="wind", theta_obs_max=0.2)])),
                   ("batch_rs_nonaxisym_tophat", configs.random_draw(8, seed=49, rvs=True, theta_obs_max=0.2))):
        P7["axisymmetric"] = 0
        save(nm, P7, t, nu)
    # axisymmetric=False together with spreading=True: one ODE row per (phi, theta) cell on its own time lattice
    # (build_time_grid, grid-refinement.h:612-619)
    for nm, P7b in (("batch_fs_spreading_nonaxisym", np.concatenate([configs.random_draw(4, seed=60, theta_obs_max=0.3),
                                                                     configs.random_draw(3, seed=61, jet="gaussian", theta_obs_max=0.4),
                                                                     configs.random_draw(3, seed=62, jet="powerlaw", medium="wind", theta_obs_max=0.2)])),
                    ("batch_rs_spreading_nonaxisym", configs.random_draw(4, seed=63, rvs=True, theta_obs_max=0.2))):
        P7b["axisymmetric"], P7b["spreading"] = 0, 1
        save(nm, P7b, t, nu)
    # Wind(A_star, n_ism, n0, k_m != 2): the reference's generic-Medium path (pybind/pymodel.cpp:169-185)
    P8 = np.concatenate([configs.random_draw(6, seed=50, medium="wind", theta_obs_max=0.2),
                         configs.random_draw(5, seed=51, jet="gaussian", medium="wind", theta_obs_max=0.3),
                         configs.random_draw(5, seed=52, medium="wind", rvs=True)])
    P8["wind_k_m"] = np.random.default_rng(53).uniform(0.5, 2.8, P8.size)
    P8["n0"][::3] = 1e5
    P8["n_ism"][1::4] = 1e-3
    save("batch_mixed_wind_k", P8, t, nu)
    nu_ssc = np.array([1e9, 1e14, 1e17, 1e22, 1e25])
    save("batch_ssc_kn_tophat_ism", configs.random_draw(24, seed=21, ssc=True, kn=True), t, nu_ssc)
    save("batch_ssc_thomson_tophat_wind", configs.random_draw(12, seed=22, ssc=True, kn=False, medium="wind"), t, nu_ssc)
    save("batch_ssc_kn_rs_tophat_ism", configs.random_draw(12, seed=23, ssc=True, kn=True, rvs=True), t, nu_ssc)
    P9 = configs.random_draw(6, seed=24, ssc=True, kn=True, theta_obs_max=0.3)
    P9["spreading"] = 1  # SSC output band of a spreading jet: per-node Doppler extrema
    save("batch_ssc_spreading_tophat", P9, t, nu_ssc)
    ts = np.sort(np.tile(np.logspace(2.5, 6.5, 20), 5))
    nus = np.tile([1e9, 5e9, 4.84e14, 1e17, 1e18], 20)
    save("series_rs_tophat_ism", configs.random_draw(48, seed=16, rvs=True), ts, nus, series=True)
    save("series_rs_gauss", configs.random_draw(8, seed=17, rvs=True, jet="gaussian", theta_obs_max=0.4), ts, nus, series=True)
    # 5. edge cases the reference tests exercise (tests/python/test_parameter_corners.py,
    #    test_pybind_validation.py): single point, repeated times, extreme early/late windows,
    #    unsorted frequencies, p < 2, tiny / huge Gamma0, off-axis beyond the jet edge
    pe = configs.make()
    save("edge_single_point", pe, np.array([1e4]), np.array([1e14]))
    save("edge_repeated_times", pe, np.array([1e3, 1e3, 1e4, 1e4, 1e5]), np.array([1e17, 1e9]))
    save("edge_tiny_times", pe, np.array([1e-9, 1e-8]), np.array([1e9, 1e14, 1e17]))
    save("edge_late_times", pe, np.logspace(8, 11, 12), np.array([1e9, 1e14]))
    corners = np.concatenate([
        configs.make(fwd=(0.1, 1e-3, 1.8)),                       # 1 < p < 2
        configs.make(fwd=(0.1, 1e-3, 3.4)),                       # p > 3: cyclotron correction branch
        configs.make(Gamma0=2.0),                                 # barely relativistic
        configs.make(Gamma0=2000.0, E_iso=1e54),                  # ultra-relativistic
        configs.make(theta_obs=0.5),                              # tophat far off-axis
        configs.make(theta_c=1.5, theta_obs=0.0),                 # nearly spherical
        configs.make(n_ism=1e-5),                                 # tenuous medium
        configs.make(n_ism=1e4, fwd=(0.3, 0.1, 2.2)),             # dense, strongly self-absorbed
        configs.make(medium="wind", A_star=1.0, n0=1e3),          # wind with inner density floor
        configs.make(medium="wind_ism", A_star=0.01, n_ism=0.1),  # wind + ISM floor
        configs.make(radiative_fireball=False, rvs=(0.1, 0.01, 2.3), duration=100.0),
        configs.make(xi_e=0.1, rvs=(0.05, 0.001, 2.6), rvs_xi_e=0.3),
        configs.make(jet="powerlaw", k_e=4.0, k_g=1.0, theta_obs=0.05),
        configs.make(jet="gaussian", theta_c=0.05, theta_obs=0.0, resolutions=(0.1, 0.3, 5.0)),
        configs.make(z=3.0, lumi_dist=8e28, rtol=1e-8),
    ])
    save("edge_parameter_corners", corners, np.logspace(1, 8, 30), np.array([1e8, 1e11, 1e14, 1e18, 1e22]))
    if ONLY:
        return
    # 6. Model.flux (band integration) and Model.flux_density_exposures through the reference's own
    #    pybind11 module (oracle/_ref/VegasAfterglowC*.so)
    va = ref.pymodule()
    mdl = va.Model(jet=va.TophatJet(0.1, 1e52, 300, duration=1e3), medium=va.ISM(1), observer=va.Observer(1e27, 0.5, 0),
                   fwd_rad=va.Radiation(0.1, 1e-3, 2.3), rvs_rad=va.Radiation(0.1, 1e-2, 2.5))
    pb = configs.make(duration=1e3, lumi_dist=1e27, z=0.5, rvs=(0.1, 1e-2, 2.5))
    tb = np.logspace(2, 7, 25)
    band = {}
    for num in (2, 4, 5, 7, 12, 21):
        fb = mdl.flux(tb, 2.4e17 * 0.3, 2.4e17 * 10, num)
        band[str(num)] = np.stack([np.asarray(fb.total), np.asarray(fb.fwd.sync), np.asarray(fb.rvs.sync)])
    np.savez_compressed(os.path.join(OUT, "method_flux_band.npz"), params=pb, t=tb, nu_min=2.4e17 * 0.3, nu_max=2.4e17 * 10,
                        **{"num_" + k: v for k, v in band.items()})
    te = np.array([1e3, 1e3, 5e3, 2e4, 2e4, 1e5, 1e6])
    nue = np.array([1e9, 1e17, 4.84e14, 1e9, 1e17, 4.84e14, 1e17])
    ex = np.array([100.0, 500.0, 1e3, 2e4, 1e3, 5e4, 1e5])
    fe = mdl.flux_density_exposures(te, nue, ex, 10)
    np.savez_compressed(os.path.join(OUT, "method_exposures.npz"), params=pb, t=te, nu=nue, expo=ex, num_points=10,
                        flux=np.stack([np.asarray(fe.total), np.asarray(fe.fwd.sync), np.asarray(fe.rvs.sync)]))
    with open(os.path.join(OUT, "PINNING.json"), "w") as fh:
        json.dump({"what": "max relative deviation (bins > 1% of peak) of oracle/_ref (reference built here, "
                           "x86-64-v3, g++ 13) from the reference's own committed golden .npz",
                   "deviation": pin}, fh, indent=1)
    print(json.dumps(pin, indent=1))


if __name__ == "__main__":
    main()
