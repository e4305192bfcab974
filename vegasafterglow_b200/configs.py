"""Named parameter sets: BASELINE.json configs C1-C5 and the in-scope
golden configurations of the reference (tests/python/golden/regenerate.py:47-138)."""
from __future__ import annotations

import numpy as np

from . import abi


def make(jet="tophat", theta_c=0.1, E_iso=1e52, Gamma0=300.0, k_e=2.0, k_g=2.0, duration=1.0, medium="ism",
         n_ism=1.0, A_star=0.1, n0=np.inf, lumi_dist=1e26, z=0.1, theta_obs=0.0, fwd=(0.1, 1e-3, 2.3), rvs=None,
         resolutions=None, rtol=0.0, radiative_fireball=True, xi_e=1.0, rvs_xi_e=1.0, ssc=False, kn=False,
         rvs_ssc=False, rvs_kn=False, theta_w=0.3, E_iso_w=1e50, Gamma0_w=50.0, sigma0=0.0, magnetar=None, k_m=2.0):
    p = abi.default_params(1)
    p["jet_type"] = {"tophat": abi.JET_TOPHAT, "gaussian": abi.JET_GAUSSIAN, "powerlaw": abi.JET_POWERLAW,
                     "two_component": abi.JET_TWO_COMPONENT, "step_powerlaw": abi.JET_STEP_POWERLAW,
                     "powerlaw_wing": abi.JET_POWERLAW_WING}[jet]
    p["theta_w"], p["E_iso_w"], p["Gamma0_w"], p["sigma0"] = theta_w, E_iso_w, Gamma0_w, sigma0
    if magnetar is not None:  # (L0 [erg/s], t0 [s], q)
        p["has_magnetar"] = 1
        p["magnetar_L0"], p["magnetar_t0"], p["magnetar_q"] = magnetar
    p["theta_c"], p["E_iso"], p["Gamma0"], p["k_e"], p["k_g"], p["duration"] = theta_c, E_iso, Gamma0, k_e, k_g, duration
    if medium == "ism":
        p["medium_type"], p["n_ism"] = abi.MEDIUM_ISM, n_ism
    else:
        p["medium_type"], p["A_star"], p["n_ism"], p["n0"] = abi.MEDIUM_WIND, A_star, 0.0 if medium == "wind" else n_ism, n0
        p["wind_k_m"] = k_m
    p["lumi_dist"], p["z"], p["theta_obs"] = lumi_dist, z, theta_obs
    p["fwd"]["eps_e"], p["fwd"]["eps_B"], p["fwd"]["p"] = fwd
    p["fwd"]["xi_e"] = xi_e
    p["fwd"]["ssc"], p["fwd"]["kn"] = int(ssc), int(kn)
    if rvs is not None:
        p["has_rvs"] = 1
        p["rvs"]["eps_e"], p["rvs"]["eps_B"], p["rvs"]["p"] = rvs
        p["rvs"]["xi_e"] = rvs_xi_e
        p["rvs"]["ssc"], p["rvs"]["kn"] = int(rvs_ssc), int(rvs_kn)
    if resolutions is not None:
        p["phi_resol"], p["theta_resol"], p["t_resol"] = resolutions
    p["rtol"] = rtol
    p["radiative_fireball"] = 1 if radiative_fireball else 0
    return p


# BASELINE.json configs
def C1():
    return make(), np.logspace(2, 8, 100), np.array([1e9, 1e14, 1e17])


def C2(dense=False):
    p = make(jet="gaussian", theta_obs=0.3, resolutions=(0.3, 1.0, 10) if dense else None)
    return p, np.logspace(2, 8, 200), np.logspace(9, 18, 8)


def C3():
    p = make(duration=1e4, medium="wind", A_star=0.1, lumi_dist=1e28, z=1.0, fwd=(0.1, 0.01, 2.3), rvs=(0.1, 0.01, 2.3))
    return p, np.logspace(1, 7, 100), np.array([1e9, 4.84e14, 1e18])


def C4():
    p = make(jet="powerlaw", k_e=2.0, k_g=2.0, theta_obs=0.2, fwd=(0.1, 0.01, 2.3), ssc=True, kn=True)
    return p, np.logspace(2, 7, 50), np.logspace(9, 27, 40)


# in-scope goldens of the reference (typed jets, no magnetisation)
GOLDEN_T = np.logspace(2, 8, 40)
GOLDEN_NU = np.array([1e9, 1e14, 1e17, 1e22])
GOLDEN = {
    "tophat_ism_adiabatic": dict(E_iso=1e53, lumi_dist=3e28, z=0.5, radiative_fireball=False),
    "tophat_ism": dict(E_iso=1e53, lumi_dist=3e28, z=0.5),
    "rs_thick": dict(E_iso=1e53, Gamma0=100.0, duration=1000.0, lumi_dist=3e28, z=0.5, rvs=(0.1, 1e-2, 2.5)),
    "gauss_ism_rs": dict(jet="gaussian", lumi_dist=1e28, z=1.0, theta_obs=0.4, fwd=(0.1, 0.01, 2.3),
                         rvs=(0.1, 0.01, 2.3), resolutions=(0.1, 1.2, 10)),
    "powerlaw_wind_rs": dict(jet="powerlaw", medium="wind", A_star=0.1, lumi_dist=1e28, z=1.0, theta_obs=0.3,
                             fwd=(0.1, 0.01, 2.3), rvs=(0.1, 0.01, 2.3)),
    "two_component_ism": dict(jet="two_component", theta_c=0.05, theta_w=0.3, E_iso_w=1e50, Gamma0_w=50.0,
                              lumi_dist=1e28, z=1.0, theta_obs=0.15, fwd=(0.1, 0.01, 2.3)),
    "tophat_sigma_rs": dict(sigma0=0.1, lumi_dist=3e28, z=0.5, rvs=(0.1, 1e-2, 2.5)),
    "tophat_sigma1_rs": dict(sigma0=1.0, lumi_dist=3e28, z=0.5, rvs=(0.1, 1e-2, 2.5)),
    "tophat_sigma10_rs": dict(sigma0=10.0, lumi_dist=3e28, z=0.5, rvs=(0.1, 1e-2, 2.5)),
    "gauss_wind_ssc": dict(jet="gaussian", E_iso=1e53, medium="wind", A_star=0.1, lumi_dist=3e28, z=1.0, theta_obs=0.2,
                           fwd=(0.1, 1e-4, 2.3), ssc=True, kn=True),
    "dense_ism_ssa_ssc": dict(n_ism=1e5, fwd=(0.1, 3e-2, 2.5), ssc=True, kn=True),
    "ism_absorbed_slow_ssc": dict(n_ism=1e3, fwd=(0.1, 1e-6, 2.5), ssc=True, kn=True),
}


def golden(name):
    return make(**GOLDEN[name])


def random_draw(n, seed=0, rvs=False, jet="tophat", medium="ism", theta_obs_max=0.0, ssc=False, kn=False):
    """Common synthetic draw of SURVEY.md section 8d (numpy default_rng(seed))."""
    rng = np.random.default_rng(seed)
    p = np.repeat(make(jet=jet, medium=medium), n)
    p["E_iso"] = 10 ** rng.uniform(51, 54, n)
    p["Gamma0"] = 10 ** rng.uniform(np.log10(50), 3, n)
    p["theta_c"] = rng.uniform(0.03, 0.4, n)
    if medium == "ism":
        p["n_ism"] = 10 ** rng.uniform(-3, 1, n)
    else:
        p["A_star"] = 10 ** rng.uniform(-2, 0, n)
    p["fwd"]["eps_e"] = 10 ** rng.uniform(-2, -0.5, n)
    p["fwd"]["eps_B"] = 10 ** rng.uniform(-4, -1, n)
    p["fwd"]["p"] = rng.uniform(2.1, 2.8, n)
    if rvs:
        p["has_rvs"] = 1
        p["rvs"]["eps_e"] = 10 ** rng.uniform(-2, -0.5, n)
        p["rvs"]["eps_B"] = 10 ** rng.uniform(-4, -1, n)
        p["rvs"]["p"] = rng.uniform(2.1, 2.8, n)
        p["duration"] = 10 ** rng.uniform(0, 4, n)
    if theta_obs_max > 0:
        p["theta_obs"] = rng.uniform(0, theta_obs_max, n)
    if ssc:
        p["fwd"]["ssc"], p["fwd"]["kn"] = 1, int(kn)
        if rvs:
            p["rvs"]["ssc"], p["rvs"]["kn"] = 1, int(kn)
    return p
