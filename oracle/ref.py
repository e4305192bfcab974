"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libvagref.so (the unmodified
reference compiled by oracle/Makefile, driven by oracle/ref_driver.cpp) and a loader for the
reference's own pybind11 module oracle/_ref/VegasAfterglowC*.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.util
import os

import numpy as np

from vegasafterglow_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_libs = {}
_variant = "main"  # "main": reference flags (x86-64-v3, FMA, xsimd); "alt": -O2, no FMA, no xsimd


def _path(variant):
    return os.path.join(_REF_DIR, "libvagref.so" if variant == "main" else "libvagref_alt.so")


def available(variant="main") -> bool:
    return os.path.exists(_path(variant))


class use_variant:
    """Context manager selecting which build of the unmodified reference the calls below use."""

    def __init__(self, variant):
        self.variant = variant

    def __enter__(self):
        global _variant
        self._old, _variant = _variant, self.variant

    def __exit__(self, *a):
        global _variant
        _variant = self._old


def lib():
    if _variant not in _libs:
        path = _path(_variant)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` where /root/reference exists")
        l = C.CDLL(path)
        l.vagref_hardware_threads.restype = C.c_int
        _libs[_variant] = l
    return _libs[_variant]


_pymod = None


def pymodule():
    """The reference's own pybind11 module (Model, TophatJet, ...) built from its sources.  Loaded once per process
    (pybind11 registers its types globally): a copy imported as ``VegasAfterglow.VegasAfterglowC`` is reused."""
    global _pymod
    import sys

    if _pymod is None:
        _pymod = sys.modules.get("VegasAfterglow.VegasAfterglowC")
    if _pymod is not None:
        return _pymod
    cands = glob.glob(os.path.join(_REF_DIR, "VegasAfterglowC*.so"))
    if not cands:
        raise RuntimeError("oracle/_ref/VegasAfterglowC*.so missing: run `make -C oracle`")
    spec = importlib.util.spec_from_file_location("VegasAfterglowC", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _pymod = mod
    return mod


def hardware_threads() -> int:
    return int(lib().vagref_hardware_threads())


def _params(params):
    p = np.ascontiguousarray(params, dtype=abi.PARAMS_DTYPE).reshape(-1)
    return p, p.ctypes.data_as(C.c_void_p)


def flux_density_grid(params, t, nu, n_threads=1):
    """Reference ``Model.flux_density_grid`` for a batch -> [n, 5, n_nu, n_t]."""
    p, pp = _params(params)
    t = np.ascontiguousarray(t, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    out = np.zeros((p.size, abi.NCOMP, nu.size, t.size))
    lib().vagref_flux_density_grid(pp, C.c_size_t(p.size), abi.as_ptr(t), C.c_size_t(t.size), abi.as_ptr(nu),
                                   C.c_size_t(nu.size), abi.as_ptr(out), C.c_int(n_threads))
    return out


def flux_density_series(params, t, nu, n_threads=1):
    """Reference ``Model.flux_density`` for a batch -> [n, 5, n_pts]."""
    p, pp = _params(params)
    t = np.ascontiguousarray(t, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    assert t.size == nu.size
    out = np.zeros((p.size, abi.NCOMP, t.size))
    lib().vagref_flux_density_series(pp, C.c_size_t(p.size), abi.as_ptr(t), abi.as_ptr(nu), C.c_size_t(t.size),
                                     abi.as_ptr(out), C.c_int(n_threads))
    return out


def chi2_series(params, t, nu, lnF, sigma_ln, w, n_threads=1):
    p, pp = _params(params)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (t, nu, lnF, sigma_ln, w)]
    out = np.zeros(p.size)
    lib().vagref_chi2_series(pp, C.c_size_t(p.size), *[abi.as_ptr(a) for a in arrs], C.c_size_t(arrs[0].size),
                             abi.as_ptr(out), C.c_int(n_threads))
    return out


def details(param, t_min, t_max):
    """Stage tables of one model (code units): dict with info, theta, phi, reps, t_rows,
    fwd_shock[7,n_reps,n_t], rvs_shock, inj_idx, lg2_t, lg2_doppler, lg2_geom."""
    p, pp = _params(param)
    assert p.size == 1
    info = np.zeros(1, dtype=abi.GRID_INFO_DTYPE)
    null = None
    f = lib().vagref_details
    f(pp, C.c_double(t_min), C.c_double(t_max), info.ctypes.data_as(C.c_void_p), *([null] * 10))
    i = info[0]
    n_phi, n_theta, n_t, n_reps, n_pe = (int(i[k]) for k in ("n_phi", "n_theta", "n_t", "n_reps", "n_phi_eff"))
    d = {
        "info": i,
        "theta": np.zeros(n_theta),
        "phi": np.zeros(n_phi),
        "reps": np.zeros(n_reps, dtype=np.int32),
        "t_rows": np.zeros((n_reps, n_t)),
        "fwd_shock": np.zeros((7, n_reps, n_t)),
        "rvs_shock": np.zeros((7, n_reps, n_t)),
        "inj_idx": np.zeros(n_reps, dtype=np.int32),
        "lg2_t": np.zeros((n_pe, n_theta, n_t)),
        "lg2_doppler": np.zeros((n_pe, n_theta, n_t)),
        "lg2_geom": np.zeros((n_pe, n_theta, n_t)),
    }
    f(pp, C.c_double(t_min), C.c_double(t_max), info.ctypes.data_as(C.c_void_p), abi.as_ptr(d["theta"]),
      abi.as_ptr(d["phi"]), abi.as_ptr(d["reps"], abi.c_int32_p), abi.as_ptr(d["t_rows"]),
      abi.as_ptr(d["fwd_shock"]), abi.as_ptr(d["rvs_shock"]), abi.as_ptr(d["inj_idx"], abi.c_int32_p),
      abi.as_ptr(d["lg2_t"]), abi.as_ptr(d["lg2_doppler"]), abi.as_ptr(d["lg2_geom"]))
    return d
