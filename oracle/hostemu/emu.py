"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/hostemu/libvagemu.so (the kernel bodies of
vegasafterglow_b200/csrc executed sequentially on the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

from vegasafterglow_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
PH_NCOEF = 15


def build(force=False):
    so = os.path.join(_HERE, "libvagemu.so")
    src = os.path.join(_HERE, "hostemu.cpp")
    csrc = os.path.join(_HERE, "..", "..", "vegasafterglow_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=fast", "-fPIC",
                               "-shared", "-x", "c++", src, "-o", so])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _params(params):
    p = np.ascontiguousarray(params, dtype=abi.PARAMS_DTYPE).reshape(-1)
    return p, p.ctypes.data_as(C.c_void_p)


def flux_density_grid(params, t, nu):
    p, pp = _params(params)
    t = np.ascontiguousarray(t, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    out = np.zeros((p.size, abi.NCOMP, nu.size, t.size))
    st = np.zeros(p.size, dtype=np.int32)
    lib().vagemu_flux_density_grid(pp, C.c_size_t(p.size), abi.as_ptr(t), C.c_size_t(t.size), abi.as_ptr(nu),
                                   C.c_size_t(nu.size), abi.as_ptr(out), abi.as_ptr(st, abi.c_int32_p))
    return out, st


def flux_density_series(params, t, nu, series_mode=0):
    """series_mode as vag_set_series_mode (include/vag.h): 0 auto, 1 per-point spectra, 2 banded whenever possible."""
    lib().vagemu_set_series_mode(C.c_int(int(series_mode)))
    p, pp = _params(params)
    t = np.ascontiguousarray(t, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    out = np.zeros((p.size, abi.NCOMP, t.size))
    st = np.zeros(p.size, dtype=np.int32)
    lib().vagemu_flux_density_series(pp, C.c_size_t(p.size), abi.as_ptr(t), abi.as_ptr(nu), C.c_size_t(t.size),
                                     abi.as_ptr(out), abi.as_ptr(st, abi.c_int32_p))
    return out, st


def details(param, t_min, t_max):
    p, pp = _params(param)
    info = np.zeros(1, dtype=abi.GRID_INFO_DTYPE)
    f = lib().vagemu_details
    f(pp, C.c_double(t_min), C.c_double(t_max), info.ctypes.data_as(C.c_void_p), *([None] * 9))
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
        "coef_fwd": np.zeros((PH_NCOEF, n_reps, n_t)),
        "coef_rvs": np.zeros((PH_NCOEF, n_reps, n_t)),
    }
    f(pp, C.c_double(t_min), C.c_double(t_max), info.ctypes.data_as(C.c_void_p), abi.as_ptr(d["theta"]),
      abi.as_ptr(d["phi"]), abi.as_ptr(d["reps"], abi.c_int32_p), abi.as_ptr(d["t_rows"]),
      abi.as_ptr(d["fwd_shock"]), abi.as_ptr(d["rvs_shock"]), abi.as_ptr(d["inj_idx"], abi.c_int32_p),
      abi.as_ptr(d["coef_fwd"]), abi.as_ptr(d["coef_rvs"]))
    return d
