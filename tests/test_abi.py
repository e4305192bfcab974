"""C-ABI library: loads, exports every symbol include/vag.h declares, struct layout, loud failure
without a GPU, argument validation mirrors the reference's ValueError conventions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from vegasafterglow_b200 import _lib, abi, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "vag.h")).read()
    declared = set(re.findall(r"\b(vag_[a-z0-9_]+)\s*\(", header))
    assert {"vag_flux_density_grid", "vag_flux_density_series", "vag_chi2_series", "vag_create"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"libvag_b200.so does not export {name}"
    assert set(_lib.EXPORTS) == declared


def test_params_layout_matches_header(tmp_path):
    # compile include/vag.h with gcc and compare every field offset with the numpy mirror
    fields = [k for k in abi.PARAMS_DTYPE.names]
    src = tmp_path / "off.c"
    body = "".join(f'printf("{k} %zu\\n", offsetof(vag_params, {k}));' for k in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){%s printf("sizeof %%zu\\n", '
                   'sizeof(vag_params)); printf("info %%zu\\n", sizeof(vag_grid_info)); return 0;}'
                   % (os.path.join(ROOT, "include", "vag.h"), body))
    exe = tmp_path / "off"
    import subprocess

    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for k in fields:
        assert int(out[k]) == abi.PARAMS_DTYPE.fields[k][1], k
    assert int(out["sizeof"]) == abi.PARAMS_DTYPE.itemsize == 320
    assert int(out["info"]) == abi.GRID_INFO_DTYPE.itemsize


def test_defaults_match_reference_defaults():
    lib = _lib.load()
    p = np.zeros(1, dtype=abi.PARAMS_DTYPE)
    lib.vag_params_default(p.ctypes.data)
    q = abi.default_params(1)
    for k in ("k_e", "k_g", "duration", "n0", "axisymmetric", "radiative_fireball", "rtol", "phi_resol"):
        assert p[k][0] == q[k][0], k
    assert p["fwd"]["xi_e"][0] == 1.0


@pytest.mark.parametrize("field,value,msg", [
    ("theta_c", 0.0, "theta_c"), ("theta_c", 2.0, "theta_c"), ("E_iso", -1.0, "E_iso"), ("Gamma0", 1.0, "Gamma0"),
    ("duration", 0.0, "duration"), ("lumi_dist", 0.0, "lumi_dist"), ("z", -0.1, "z"), ("theta_obs", 4.0, "theta_obs"),
    ("n_ism", -1.0, "n_ism"), ("rtol", 1.5, "rtol"),
])
def test_validation_rejects_what_the_reference_rejects(field, value, msg):
    # pybind/pymodel.cpp:47-186, pymodel.h:190-204,642: AFTERGLOW_REQUIRE -> ValueError
    lib = _lib.load()
    p = configs.make()
    p[field] = value
    rc = lib.vag_params_validate(p.ctypes.data)
    assert rc == abi.VAG_ERR_INVALID
    assert msg in lib.vag_last_error().decode()


def test_validation_radiation_ranges():
    lib = _lib.load()
    for k, v in (("eps_e", 0.0), ("eps_e", 1.5), ("eps_B", 0.0), ("xi_e", 0.0), ("p", 1.0)):
        p = configs.make()
        p["fwd"][k] = v
        assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_ERR_INVALID, (k, v)
    p = configs.make()
    p["fwd"]["p"] = 1.8  # 1 < p < 2 is legal (pymodel.h:309-311)
    assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_OK


def test_grid_switches_are_accepted():
    lib = _lib.load()
    p = configs.make()
    p["spreading"] = 1  # lateral spreading is implemented
    assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_OK
    p = configs.make()
    p["axisymmetric"] = 0  # implemented ...
    assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_OK
    p["spreading"] = 1     # ... also together with spreading (one ODE row per (phi, theta) cell)
    assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_OK
    p = configs.make(ssc=True, kn=True)  # inverse Compton is implemented
    assert lib.vag_params_validate(p.ctypes.data) == abi.VAG_OK


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.vag_create(0, C.byref(h))
    assert rc == abi.VAG_ERR_CUDA and "CUDA" in lib.vag_last_error().decode()
    from vegasafterglow_b200.engine import Engine

    with pytest.raises(_lib.VagError):
        Engine(0)
