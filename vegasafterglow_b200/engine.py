"""Batched model evaluation on one GPU through the C ABI (include/vag.h).

``Engine`` owns one ``vag_context`` (a CUDA stream + growable HBM workspaces).  Host-array
methods copy inputs/outputs inside the call; ``*_dev`` methods take raw device pointers (e.g.
``torch.Tensor.data_ptr()``) and only enqueue work.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, abi


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.vag_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vag_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers -------------------------------------------------------------------------
    @staticmethod
    def _params(params):
        p = np.ascontiguousarray(params, dtype=abi.PARAMS_DTYPE).reshape(-1)
        return p

    def flux_density_grid(self, params, t, nu, return_status=False):
        """Batched ``Model.flux_density_grid`` -> float64[n_models, 5, n_nu, n_t]
        (component order ``abi.COMPONENTS``)."""
        p, t, nu = self._params(params), _f64(t).reshape(-1), _f64(nu).reshape(-1)
        out = (np.zeros if getattr(self, "_present_only", False) else np.empty)((p.size, abi.NCOMP, nu.size, t.size))
        st = np.zeros(p.size, dtype=np.int32)
        _lib.check(self._lib.vag_flux_density_grid(self._h, p.ctypes.data, p.size, t.ctypes.data, t.size,
                                                   nu.ctypes.data, nu.size, out.ctypes.data, st.ctypes.data))
        return (out, st) if return_status else out

    def flux_density_series(self, params, t, nu, return_status=False):
        """Batched ``Model.flux_density`` -> float64[n_models, 5, n]."""
        p, t, nu = self._params(params), _f64(t).reshape(-1), _f64(nu).reshape(-1)
        if t.size != nu.size:  # pybind/pymodel.cpp:376-379
            raise ValueError("time and frequency arrays must have the same size\nIf you intend to get grid-like "
                             "output, use the generic `flux_density_grid` instead")
        out = (np.zeros if getattr(self, "_present_only", False) else np.empty)((p.size, abi.NCOMP, t.size))
        st = np.zeros(p.size, dtype=np.int32)
        _lib.check(self._lib.vag_flux_density_series(self._h, p.ctypes.data, p.size, t.ctypes.data, nu.ctypes.data,
                                                     t.size, out.ctypes.data, st.ctypes.data))
        return (out, st) if return_status else out

    def flux_band(self, params, t, nu_min, nu_max, num_nu, return_status=False):
        """Batched ``Model.flux`` (band-integrated, Boole rule) -> float64[n_models, 5, n_t] [erg cm^-2 s^-1]."""
        p, t = self._params(params), _f64(t).reshape(-1)
        out = np.empty((p.size, abi.NCOMP, t.size))
        st = np.zeros(p.size, dtype=np.int32)
        _lib.check(self._lib.vag_flux_band(self._h, p.ctypes.data, p.size, t.ctypes.data, t.size, float(nu_min),
                                           float(nu_max), int(num_nu), out.ctypes.data, st.ctypes.data))
        return (out, st) if return_status else out

    def flux_density_exposures(self, params, t, nu, expo_time, num_points=10, return_status=False):
        """Batched ``Model.flux_density_exposures`` (pybind/pymodel.cpp:412-496): each point is the
        average of ``num_points`` samples spread over its exposure window (host-side sampling and
        averaging exactly as the reference does; the series itself runs on the GPU)."""
        t, nu, expo_time = (_f64(a).reshape(-1) for a in (t, nu, expo_time))
        if not (t.size == nu.size == expo_time.size):
            raise ValueError("time, frequency, and exposure time arrays must have the same size")
        if num_points < 2:
            raise ValueError("num_points must be at least 2 to sample within each exposure time")
        if not np.all(np.isfinite(expo_time) & (expo_time > 0)):
            bad = int(np.nonzero(~(np.isfinite(expo_time) & (expo_time > 0)))[0][0])
            raise ValueError(f"expo_time[{bad}] must be finite and > 0, got {expo_time[bad]}")
        k = np.arange(num_points, dtype=np.float64)
        ts = (t[:, None] + k[None, :] * (expo_time / (num_points - 1))[:, None]).reshape(-1)
        nus = np.repeat(nu, num_points)
        idx = np.repeat(np.arange(t.size), num_points)
        order = np.argsort(ts, kind="stable")
        res = self.flux_density_series(params, ts[order], nus[order], return_status=True)
        f, st = res
        out = np.zeros((f.shape[0], abi.NCOMP, t.size))
        np.add.at(out, (slice(None), slice(None), idx[order]), f)
        out /= float(num_points)
        return (out, st) if return_status else out

    def chi2_series(self, params, t, nu, lnF_obs, sigma_ln, w, return_status=False):
        """Batched ``Fitter._evaluate`` for point data -> chi2[n_models] (+inf where non-finite)."""
        p = self._params(params)
        t, nu, lnF_obs, sigma_ln, w = (_f64(a).reshape(-1) for a in (t, nu, lnF_obs, sigma_ln, w))
        if not (t.size == nu.size == lnF_obs.size == sigma_ln.size == w.size):
            raise ValueError("t, nu, lnF_obs, sigma_ln and w must have the same size")
        chi2 = np.empty(p.size)
        st = np.zeros(p.size, dtype=np.int32)
        _lib.check(self._lib.vag_chi2_series(self._h, p.ctypes.data, p.size, t.ctypes.data, nu.ctypes.data,
                                             lnF_obs.ctypes.data, sigma_ln.ctypes.data, w.ctypes.data, t.size,
                                             chi2.ctypes.data, st.ctypes.data))
        return (chi2, st) if return_status else chi2

    def chi2(self, params, points=None, bands=(), return_status=False):
        """Batched ``Fitter._evaluate`` (fitter.py:503-533) -> chi2[n_models]: the point-data term (``points`` =
        (t, nu, lnF_obs, sigma_ln, w), times ascending, or None) plus one term per band-integrated data set
        (``bands``: dicts with t, lnF_obs, sigma_ln, w, nu_min, nu_max, num_nu -- ``Model.flux`` on each).
        +inf where the sum is not finite or the model's ODE failed (include/vag.h vag_chi2)."""
        p = self._params(params)
        keep = []  # the arrays must outlive the call
        if points is not None:
            pt = [_f64(a).reshape(-1) for a in points]
            if len(pt) != 5 or len({a.size for a in pt}) != 1:
                raise ValueError("points must be (t, nu, lnF_obs, sigma_ln, w) of one size")
            n_pt = pt[0].size
        else:
            pt, n_pt = [np.zeros(0)] * 5, 0
        arr = (abi.BandObs * max(len(bands), 1))()
        for i, b in enumerate(bands):
            cols = [_f64(b[k]).reshape(-1) for k in ("t", "lnF_obs", "sigma_ln", "w")]
            if len({a.size for a in cols}) != 1:
                raise ValueError("band arrays t, lnF_obs, sigma_ln, w must have the same size")
            keep.append(cols)
            arr[i] = abi.BandObs(cols[0].ctypes.data, cols[1].ctypes.data, cols[2].ctypes.data, cols[3].ctypes.data,
                                 cols[0].size, float(b["nu_min"]), float(b["nu_max"]), int(b["num_nu"]))
        out = np.empty(p.size)
        st = np.zeros(p.size, dtype=np.int32)
        _lib.check(self._lib.vag_chi2(self._h, p.ctypes.data, p.size, *[a.ctypes.data if n_pt else None for a in pt], n_pt,
                                      C.cast(arr, C.c_void_p) if len(bands) else None, len(bands), out.ctypes.data,
                                      st.ctypes.data))
        return (out, st) if return_status else out

    @staticmethod
    def valid_mask(params):
        """bool[n]: which parameter sets pass the constructor checks of the reference (vag_params_validate_batch)."""
        p = np.ascontiguousarray(params, dtype=abi.PARAMS_DTYPE).reshape(-1)
        ok = np.zeros(p.size, dtype=np.int32)
        _lib.check(_lib.load().vag_params_validate_batch(p.ctypes.data, p.size, ok.ctypes.data))
        return ok.astype(bool)

    # ---- device buffers (raw pointers) ---------------------------------------------------------
    def flux_density_grid_dev(self, d_params, n_models, d_t, n_t, d_nu, n_nu, d_out, d_status=0, stream=0):
        _lib.check(self._lib.vag_flux_density_grid_dev(self._h, d_params, n_models, d_t, n_t, d_nu, n_nu, d_out,
                                                       d_status or None, stream or None))

    def flux_density_series_dev(self, d_params, n_models, d_t, d_nu, n, d_out, d_status=0, stream=0):
        _lib.check(self._lib.vag_flux_density_series_dev(self._h, d_params, n_models, d_t, d_nu, n, d_out,
                                                         d_status or None, stream or None))

    def chi2_series_dev(self, d_params, n_models, d_t, d_nu, d_lnF, d_sig, d_w, n, d_chi2, d_status=0, stream=0):
        _lib.check(self._lib.vag_chi2_series_dev(self._h, d_params, n_models, d_t, d_nu, d_lnF, d_sig, d_w, n, d_chi2,
                                                 d_status or None, stream or None))

    def synchronize(self):
        _lib.check(self._lib.vag_synchronize(self._h))

    def set_capacity(self, cap_theta, cap_phi):
        _lib.check(self._lib.vag_set_capacity(self._h, int(cap_theta), int(cap_phi)))

    def set_output_mode(self, present_only: bool, alias_total: bool = False):
        """True: host-buffer calls skip the planes of components no model of the batch has (vag.h VAG_OUT_PRESENT);
        alias_total: additionally skip `total` when it equals the batch's only component (``last_total_alias()``)."""
        _lib.check(self._lib.vag_set_output_mode(self._h, (2 if alias_total else 1) if present_only else 0))
        self._present_only = bool(present_only)

    def set_series_mode(self, mode: int):
        """Evaluation of series requests with <= 8 distinct frequencies (vag.h vag_set_series_mode): 0 auto (default),
        1 per-point spectra always, 2 banded whenever possible."""
        _lib.check(self._lib.vag_set_series_mode(self._h, int(mode)))

    def last_total_alias(self):
        """Index into ``abi.COMPONENTS`` of the plane that holds `total` after the last host call, or -1."""
        return int(self._lib.vag_last_total_alias(self._h))

    def debug_set_ode_limits(self, max_steps=0, max_fails=0):
        """Test hook (include/vag.h vag_debug_set_ode_limits): 0 restores the reference limits."""
        _lib.check(self._lib.vag_debug_set_ode_limits(self._h, int(max_steps), int(max_fails)))

    def set_profiling(self, on=True):
        _lib.check(self._lib.vag_set_profiling(self._h, 1 if on else 0))

    def last_stage_ms(self):
        ms = (C.c_float * 8)()
        self._lib.vag_last_stage_ms(self._h, ms)
        names = ("grid", "dynamics", "radiation", "eats", "finish")
        return {k: float(ms[i]) for i, k in enumerate(names)}

    def last_work(self):
        """Work counters of the last profiled pass (include/vag.h vag_last_work)."""
        w = (C.c_double * 8)()
        self._lib.vag_last_work(self._h, w)
        names = ("rows_fwd", "rows_pair", "cells", "eats_cells", "eats_rows", "quad_attempts_theta", "quad_phi_evals", "sum_n_theta")
        return {k: float(w[i]) for i, k in enumerate(names)}

    def measure_fp64_peak(self):
        v = C.c_double()
        _lib.check(self._lib.vag_measure_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    LIBM_FUNCS = ("exp", "exp2", "log", "log2", "log10", "pow", "sin", "cos")

    def selftest_libm(self, fn, x, y=None):
        """Device evaluation of one libm-exact function of csrc/vag_libm.cuh (include/vag.h vag_selftest_libm)."""
        x = _f64(x).reshape(-1)
        out = np.empty_like(x)
        yv = None if y is None else _f64(y).reshape(-1)
        _lib.check(self._lib.vag_selftest_libm(self._h, self.LIBM_FUNCS.index(fn), x.ctypes.data,
                                               None if yv is None else yv.ctypes.data, out.ctypes.data, x.size))
        return out

    def last_launch_count(self):
        return int(self._lib.vag_last_launch_count(self._h))

    def details(self, param, t_min, t_max):
        """Stage tables of one model (code units) -- the analogue of ``Model.details``."""
        p = self._params(param)
        assert p.size == 1
        info = np.zeros(1, dtype=abi.GRID_INFO_DTYPE)
        f = self._lib.vag_details
        _lib.check(f(self._h, p.ctypes.data, t_min, t_max, info.ctypes.data, *([None] * 7)))
        i = info[0]
        n_phi, n_theta, n_t, n_reps = (int(i[k]) for k in ("n_phi", "n_theta", "n_t", "n_reps"))
        d = {
            "info": i,
            "theta": np.zeros(n_theta),
            "phi": np.zeros(n_phi),
            "reps": np.zeros(n_reps, dtype=np.int32),
            "t_rows": np.zeros((n_reps, n_t)),
            "fwd_shock": np.zeros((7, n_reps, n_t)),
            "rvs_shock": np.zeros((7, n_reps, n_t)),
            "inj_idx": np.zeros(n_reps, dtype=np.int32),
        }
        _lib.check(f(self._h, p.ctypes.data, t_min, t_max, info.ctypes.data, d["theta"].ctypes.data,
                     d["phi"].ctypes.data, d["reps"].ctypes.data, d["t_rows"].ctypes.data,
                     d["fwd_shock"].ctypes.data, d["rvs_shock"].ctypes.data, d["inj_idx"].ctypes.data))
        return d

    def details_ic(self, param, t_min, t_max, n_reps, n_t):
        """Electron / inverse-Compton records of the ssc=True shocks: (fwd, rvs) each [7, n_reps, n_t] -- gamma_m, gamma_c,
        gamma_a, gamma_M, gamma_m_hat, gamma_c_hat, Y_T (include/vag.h vag_details_ic); zero for shocks without ssc."""
        p = self._params(param)
        assert p.size == 1
        fwd = np.zeros((7, n_reps, n_t))
        rvs = np.zeros((7, n_reps, n_t))
        _lib.check(self._lib.vag_details_ic(self._h, p.ctypes.data, t_min, t_max, fwd.ctypes.data, rvs.ctypes.data))
        return fwd, rvs

    def details_photons(self, param, t_min, t_max, n_reps, n_t):
        """Photon tables of one model: (fwd, rvs) each [6, n_reps, n_t] -- log2 nu_m, nu_c, nu_a, nu_M, I_nu_max (code
        units) and 1/nu_M (include/vag.h vag_details_photons)."""
        p = self._params(param)
        assert p.size == 1
        fwd = np.zeros((6, n_reps, n_t))
        rvs = np.zeros((6, n_reps, n_t))
        _lib.check(self._lib.vag_details_photons(self._h, p.ctypes.data, t_min, t_max, fwd.ctypes.data, rvs.ctypes.data))
        return fwd, rvs
