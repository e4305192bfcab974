"""``Model.details(t_min, t_max)`` on the GPU path: the reference's ``SimulationDetails`` / ``ShockDetails``
(pybind/pymodel.cpp:213-348, pybind/pybind.cpp:522-584) assembled from the device stage tables.

The device keeps only the UNIQUE shock rows (symmetry collapse) and the photon coefficients; this module
broadcasts them over theta (and phi for ``axisymmetric=False``), converts to the reference's CGS / second
units and rebuilds the observer-frame grids (``t_obs``, ``Doppler``: observer.cpp:51-205) and the electron
break Lorentz factors from the break frequencies (gamma = sqrt(nu / (K B)), synchrotron.cpp:107-116).
For a shock with ``Radiation(ssc=True)`` the electron break Lorentz factors and the inverse-Compton bookkeeping
(``gamma_*``, ``gamma_m_hat``, ``gamma_c_hat``, ``Y_T``, ``nu_*_hat``) are the device's records after IC cooling
(``vag_details_ic``); without ssc the bookkeeping fields keep the defaults the reference leaves there.
"""
from __future__ import annotations

import types

import numpy as np

from . import abi

# unit system of the reference (src/util/macros.h:43-77), same expressions as csrc/vag_common.cuh
_LEN = 1.5e13
CM = 1 / _LEN
SEC = 3e10 / _LEN
G = 1 / 2e33
HZ = 1 / SEC
ERG = G * CM * CM / SEC / SEC
FLUX_DEN_CGS = ERG / (CM * CM) / SEC / HZ
GAUSS = 8.66e-11 / SEC
MP = 1.67e-24 * G
ME = MP / 1836
E_CHARGE = 4.8e-10 / 4.472136e16 / 5.809475e19 / SEC
C = 1.0
_K_SYN = 3 * E_CHARGE / (4 * np.pi * ME * C)  # nu = K B gamma^2


def _shock(tab, take):
    """[7, n_reps, n_t] device table -> dict of [n_ext, n_theta, n_t] arrays in the reference's units."""
    names = ("t_comv", "r", "theta", "Gamma", "Gamma_th", "B_comv", "N_p")
    scale = (1 / SEC, 1 / CM, 1.0, 1.0, 1.0, 1 / GAUSS, 1.0)
    return {n: take(tab[a]) * s for a, (n, s) in enumerate(zip(names, scale))}


def simulation_details(engine, param, t_min, t_max):
    p = np.ascontiguousarray(param, dtype=abi.PARAMS_DTYPE).reshape(-1)
    assert p.size == 1
    d = engine.details(p, float(t_min), float(t_max))
    n_theta, n_t = d["theta"].size, d["t_rows"].shape[1]
    reps = d["reps"]
    axis = bool(p["axisymmetric"][0])
    has_rvs = bool(p["has_rvs"][0])
    spreading = bool(p["spreading"][0]) and not has_rvs
    n_ext = 1 if axis else d["phi"].size
    # ODE row behind cell (i, j): the representative of theta_j's symmetry group -- or, for a structured model with
    # axisymmetric=False (one row per (phi, theta) cell, row = i n_theta + j), the cell's own row
    rows3d = (not axis) and bool(p["spreading"][0]) and reps.size == n_ext * n_theta and n_ext > 1
    if rows3d:
        row_of = np.arange(n_ext * n_theta).reshape(n_ext, n_theta)
    else:
        rep_of = np.searchsorted(reps, np.arange(n_theta), side="right") - 1
        row_of = np.broadcast_to(rep_of[None, :], (n_ext, n_theta))
    take = lambda a: np.ascontiguousarray(a[row_of])  # noqa: E731  [n_reps, n_t] -> [n_ext, n_theta, n_t]
    theta_v, z = float(p["theta_obs"][0]), float(p["z"][0])
    n_phi_eff = int(d["info"]["n_phi_eff"])
    ph_f, ph_r = engine.details_photons(p, float(t_min), float(t_max), reps.size, n_t)
    any_ssc = bool(p["fwd"]["ssc"][0]) or (has_rvs and bool(p["rvs"]["ssc"][0]))
    ic_f, ic_r = engine.details_ic(p, float(t_min), float(t_max), reps.size, n_t) if any_ssc else (None, None)

    out = types.SimpleNamespace()
    out.phi, out.theta = d["phi"].copy(), d["theta"].copy()
    t_code = take(d["t_rows"])                          # [n_ext, n_theta, n_t]
    out.t_src = t_code / SEC

    # observer grids (one EAT geometry serves both shocks: pybind/pymodel.cpp:337), [n_phi_eff, n_theta, n_t]
    fwd_tab = d["fwd_shock"]
    sel = (lambda a: a[:n_phi_eff]) if n_ext >= n_phi_eff else (lambda a: np.broadcast_to(a[:1], (n_phi_eff,) + a.shape[1:]))  # noqa: E731
    Gam, r, t_e = sel(take(fwd_tab[3])), sel(take(fwd_tab[1])), sel(t_code)
    th_k = sel(take(fwd_tab[2])) if spreading else np.broadcast_to(d["theta"][None, :, None], (n_phi_eff, n_theta, n_t))
    cos_phi = np.cos(d["phi"][:n_phi_eff])[:, None, None]
    cos_v = np.sin(th_k) * cos_phi * np.sin(theta_v) + np.cos(th_k) * np.cos(theta_v)
    u = np.sqrt((Gam - 1) * (Gam + 1))
    doppler = 1.0 / (Gam - u * cos_v)
    t_obs = (t_e + (1 - cos_v) * r / C) * (1 + z) / SEC

    def shock_details(tab, coef, rad, ic):
        s = types.SimpleNamespace(**_shock(tab, take))
        if not spreading:  # Shock::broadcast_groups gives every row its own coord.theta(j) (shock.cpp:41-88)
            s.theta = np.broadcast_to(d["theta"][None, :, None], (n_ext, n_theta, n_t)).copy()
        s.t_obs, s.Doppler = t_obs.copy(), doppler.copy()
        B = take(tab[5])
        with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
            nu = {k: np.exp2(take(coef[i])) for i, k in enumerate(("nu_m", "nu_c", "nu_a"))}
            nu["nu_M"] = 1.0 / take(coef[5])
            for k, v in nu.items():
                setattr(s, k, v / HZ)
                setattr(s, "gamma_" + k[3:], np.sqrt(v / (_K_SYN * B)))
            s.I_nu_max = np.exp2(take(coef[4])) / FLUX_DEN_CGS
            gm = np.sqrt(nu["nu_m"] / (_K_SYN * B))
            f_syn = (gm - 1) / gm
            if rad["p"] > 3:
                f_syn = f_syn ** ((rad["p"] - 1) / 2)
            s.N_e = take(tab[6]) * rad["xi_e"] * f_syn
        if ic is not None and bool(rad["ssc"]):
            # electrons after IC cooling and the InverseComptonY record (save_electron_details / save_photon_details)
            for a, k in enumerate(("gamma_m", "gamma_c", "gamma_a", "gamma_M", "gamma_m_hat", "gamma_c_hat", "Y_T")):
                setattr(s, k, take(ic[a]))
            with np.errstate(over="ignore", invalid="ignore"):
                s.nu_m_hat = _K_SYN * B * take(ic[4]) ** 2 / HZ  # compute_syn_freq(gamma_hat, B)
                s.nu_c_hat = _K_SYN * B * take(ic[5]) ** 2 / HZ
        else:  # InverseComptonY defaults (inverse-compton.cpp:38-44): gamma_hat = 1, Y_T = 0
            shape = (n_ext, n_theta, n_t)
            s.gamma_m_hat, s.gamma_c_hat, s.Y_T = np.ones(shape), np.ones(shape), np.zeros(shape)
            s.nu_m_hat = _K_SYN * B / HZ
            s.nu_c_hat = _K_SYN * B / HZ
        return s

    out.fwd = shock_details(fwd_tab, ph_f, p["fwd"][0], ic_f)
    if has_rvs:
        out.rvs = shock_details(d["rvs_shock"], ph_r, p["rvs"][0], ic_r)
    else:  # the reference leaves the absent shock's arrays default-constructed (0-d)
        out.rvs = types.SimpleNamespace(**{k: np.zeros(()) for k in vars(out.fwd)})
    return out
