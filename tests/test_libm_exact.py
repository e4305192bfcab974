"""csrc/vag_libm.cuh must reproduce the host libm (glibc, the library the reference's std::exp / std::pow /
std::cos ... resolve to) BIT FOR BIT: the reference's theta / phi grids are the inverse CDF of an adaptive
quadrature that amplifies a one-ulp difference in any of these functions to ~1e-8 in the nodes
(src/core/grid-refinement.h:137-189).

CPU tier: the header compiled by gcc (oracle/hostemu) against the live libm.
GPU tier: the header compiled by nvcc, evaluated on the device through the C ABI (vag_selftest_libm), against the live
libm on the same arguments."""
import ctypes as C

import numpy as np
import pytest

from oracle.hostemu import emu
from vegasafterglow_b200 import abi

FN = {"exp": 0, "exp2": 1, "log": 2, "log2": 3, "log10": 4, "pow": 5, "sin": 6, "cos": 7}


def host_eval(fn, x, y=None, ref=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    yp = None if y is None else abi.as_ptr(np.ascontiguousarray(y, dtype=np.float64))
    rc = emu.lib().vagemu_libm_eval(C.c_int(FN[fn]), abi.as_ptr(x), yp, abi.as_ptr(out), C.c_size_t(x.size),
                                    C.c_int(1 if ref else 0))
    assert rc == 0
    return out


def cases(n, seed=1):
    """(name, function, x, y): the argument ranges the grid builder produces, and wide sweeps around them."""
    r = np.random.default_rng(seed)
    u, lu = r.uniform, lambda a, b: 10.0 ** r.uniform(a, b, n)
    return [
        ("exp gaussian-profile arguments", "exp", -lu(-12, 2.7), None),
        ("exp wide (main range |x| < 512; beyond it the platform function answers)", "exp", u(-511, 511, n), None),
        ("exp tiny", "exp", u(-1, 1, n) * lu(-20, 0), None),
        ("exp2 power-law arguments", "exp2", u(-60, 60, n), None),
        ("exp2 wide (main range)", "exp2", u(-511, 511, n), None),
        ("log wide", "log", lu(-300, 300), None),
        ("log near 1", "log", u(0.9, 1.1, n), None),
        ("log calibrate arguments", "log", 1.0 + lu(-6, 8), None),
        ("log2 wide", "log2", lu(-300, 300), None),
        ("log2 near 1", "log2", u(0.93, 1.07, n), None),
        ("log2 theta / theta_c", "log2", lu(-6, 2), None),
        ("log10 wide", "log10", lu(-300, 300), None),
        ("log10 grid bounds", "log10", lu(-7, 1), None),
        ("pow(10, x) sample abscissae", "pow", np.full(n, 10.0), u(-8, 3, n)),
        ("pow(err, -1/5) step increase", "pow", lu(-4, 0), np.full(n, -1.0 / 5.0)),
        ("pow(err, -1/3) step decrease", "pow", lu(0, 14), np.full(n, -1.0 / 3.0)),
        ("pow generic (|y ln x| < 350)", "pow", lu(-15, 15), u(-8, 8, n)),
        ("pow near-1 base", "pow", u(0.9, 1.1, n), u(-50, 50, n)),
        ("sin |x| < 0.9", "sin", u(-0.9, 0.9, n), None),
        ("sin |x| < 2.5", "sin", u(-2.5, 2.5, n), None),
        ("sin phi range", "sin", u(-7, 7, n), None),
        ("sin large", "sin", u(-1e6, 1e6, n), None),
        ("sin tiny", "sin", u(-1, 1, n) * lu(-12, 0), None),
        ("cos |x| < 0.9", "cos", u(-0.9, 0.9, n), None),
        ("cos |x| < 2.5", "cos", u(-2.5, 2.5, n), None),
        ("cos phi range", "cos", u(-7, 7, n), None),
        ("cos large", "cos", u(-1e6, 1e6, n), None),
        ("cos tiny", "cos", u(-1, 1, n) * lu(-12, 0), None),
    ]


def mismatches(a, b):
    bad = a.view(np.int64) != b.view(np.int64)
    return np.nonzero(bad & ~(np.isnan(a) & np.isnan(b)))[0]


def test_host_build_equals_live_libm():
    for name, fn, x, y in cases(200_000):
        got, want = host_eval(fn, x, y), host_eval(fn, x, y, ref=True)
        bad = mismatches(got, want)
        assert bad.size == 0, (name, bad.size, float(x[bad[0]]).hex(), got[bad[0]].hex(), want[bad[0]].hex())


def test_special_arguments_fall_back_to_the_platform():
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, -1.0, 1.0, 1e308, 750.0, -750.0, 1e-300])
    for fn in ("exp", "exp2", "log", "log2", "log10", "sin", "cos"):
        with np.errstate(all="ignore"):
            got, want = host_eval(fn, x), host_eval(fn, x, ref=True)
        assert mismatches(got, want).size == 0, fn
    got = host_eval("pow", x, np.full(x.size, -0.2))
    assert mismatches(got, host_eval("pow", x, np.full(x.size, -0.2), ref=True)).size == 0


def test_tables_match_the_installed_libm(tmp_path):
    """The committed table file is what scripts/gen_libm_tables.py extracts from this machine's libm.so.6."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "vegasafterglow_b200", "csrc", "vag_libm_tables.inc")
    before = open(inc).read()
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no glibc libm.so.6 at the usual place")
    try:
        subprocess.check_call([sys.executable, os.path.join(root, "scripts", "gen_libm_tables.py"), libm],
                              stdout=subprocess.DEVNULL)
        assert open(inc).read() == before
    finally:
        open(inc, "w").write(before)


@pytest.mark.gpu
def test_device_build_equals_live_libm(engine):
    total = 0
    for name, fn, x, y in cases(400_000, seed=7):
        got = engine.selftest_libm(fn, x, y)
        want = host_eval(fn, x, y, ref=True)
        bad = mismatches(got, want)
        assert bad.size == 0, (name, bad.size, float(x[bad[0]]).hex(), got[bad[0]].hex(), want[bad[0]].hex())
        total += x.size
    assert total > 1e7
