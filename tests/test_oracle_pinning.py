"""The oracle is pinned: (1) the committed fixtures were produced by the unmodified reference
(oracle/_ref) and that reference reproduces the reference repository's own golden .npz
(tests/golden/PINNING.json, written by make_golden.py); (2) wherever oracle/_ref is present it
still reproduces the fixtures bit-for-bit; (3) the reference's own pybind11 module agrees with
the C-ABI driver around it."""
import json
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN_DIR, golden_names, load_golden
from oracle import ref
from vegasafterglow_b200 import abi, configs

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_pinning_record_is_within_reference_tolerance():
    pin = json.load(open(os.path.join(GOLDEN_DIR, "PINNING.json")))["deviation"]
    # reference acceptance contract: rtol 2e-3 + 1e-2*peak (tests/python/golden/regenerate.py:29-30)
    assert set(pin) == set(configs.GOLDEN)
    for name, dev in pin.items():
        for comp, d in dev.items():
            assert d < 2e-3, (name, comp, d)
    # all but the structured-jet reverse shock (chaotic wing rows, test_golden.py:95) pin to < 1e-7
    for name in set(configs.GOLDEN) - {"gauss_ism_rs"}:
        assert max(pin[name].values()) < 1e-7, name
    assert len(pin) == 12  # every golden configuration of the reference is in scope


@needs_ref
@pytest.mark.parametrize("name", ["config_C1", "config_C3", "golden_tophat_ism", "golden_rs_thick",
                                  "batch_fs_tophat_ism", "series_rs_tophat_ism"])
def test_reference_reproduces_fixtures_exactly(name):
    g = load_golden(name)
    fn = ref.flux_density_series if bool(g["series"]) else ref.flux_density_grid
    f = fn(g["params"], g["t"], g["nu"], n_threads=4)
    np.testing.assert_array_equal(f, g["flux"])


@needs_ref
def test_pybind_module_matches_driver():
    va = ref.pymodule()
    p, t, nu = configs.C3()
    m = va.Model(jet=va.TophatJet(0.1, 1e52, 300, duration=1e4), medium=va.Wind(0.1), observer=va.Observer(1e28, 1.0, 0),
                 fwd_rad=va.Radiation(0.1, 0.01, 2.3), rvs_rad=va.Radiation(0.1, 0.01, 2.3))
    f = m.flux_density_grid(t, nu)
    d = ref.flux_density_grid(p, t, nu)
    np.testing.assert_array_equal(np.asarray(f.total), d[0, abi.COMPONENTS.index("total")])
    np.testing.assert_array_equal(np.asarray(f.fwd.sync), d[0, abi.COMPONENTS.index("fwd_sync")])
    np.testing.assert_array_equal(np.asarray(f.rvs.sync), d[0, abi.COMPONENTS.index("rvs_sync")])


@needs_ref
def test_chi2_definition_matches_fit_smoke_formula():
    # tests/python/test_fit_smoke.py:119-127: chi2 equals the manual ln-flux formula
    g = load_golden("series_rs_tophat_ism")
    P, t, nu = g["params"][:4], g["t"], g["nu"]
    rng = np.random.default_rng(42)
    truth = g["flux"][0, 0]
    obs = truth * (1 + 0.05 * rng.standard_normal(truth.size))
    err = 0.1 * obs
    w = np.ones_like(obs)
    chi2 = ref.chi2_series(P, t, nu, np.log(obs), err / obs, w)
    manual = [np.sum(w * ((np.log(obs) - np.log(np.maximum(g["flux"][i, 0], 1e-300))) / (err / obs)) ** 2) for i in range(4)]
    np.testing.assert_allclose(chi2, manual, rtol=1e-12)
