"""The sampler-side integration: integration/samplers_gpu.patch applied to a copy of the
reference's VegasAfterglow/fitting/samplers.py, driven through the reference's OWN ``Fitter`` class.

CPU tier, runs where /root/reference exists: the reference's Python package is assembled in a temp directory (symlinks to
its sources + the pybind module oracle/_ref built from them), ``emcee`` / ``bilby`` -- absent from this image -- are
replaced by minimal stand-ins (an ensemble "sampler" that just evaluates the vectorised log-probability on a few
proposal sets), and the GPU engine by a stand-in that answers ``chi2_series`` / ``chi2`` with the unmodified reference
(oracle/_ref).  What is tested is therefore exactly the glue: ModelParams -> vag_params (``Fitter._build_model`` and the
default factories), data consolidation, bounds / priors, band records, and the patched control flow of fit_emcee /
fit_bilby.  The device arithmetic behind the same entry points is covered by tests/test_gpu_semantics.py.
"""
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = "/root/reference/VegasAfterglow"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PKG), reason="needs the reference checkout")


def _stub_samplers():
    """Stand-ins for emcee / bilby: only what VegasAfterglow.fitting touches."""
    emcee = types.ModuleType("emcee")
    emcee.moves = types.SimpleNamespace(DEMove=lambda *a, **k: "de", DESnookerMove=lambda *a, **k: "snooker")

    class EnsembleSampler:
        def __init__(self, nwalkers, ndim, log_prob_fn, vectorize=False, moves=None):
            assert vectorize
            self.fn, self.chain, self.lp = log_prob_fn, [], []

        def run_mcmc(self, pos0, nsteps, progress=False):
            rng = np.random.default_rng(0)
            pos = np.array(pos0, dtype=float)
            for _ in range(nsteps):
                self.chain.append(pos.copy())
                self.lp.append(np.asarray(self.fn(pos)))
                pos = pos + 0.01 * rng.standard_normal(pos.shape)

        def get_chain(self, discard=0, thin=1, flat=False):
            return np.concatenate(self.chain[discard::thin])

        def get_log_prob(self, discard=0, thin=1, flat=False):
            return np.concatenate(self.lp[discard::thin])

    emcee.EnsembleSampler = EnsembleSampler
    bilby = types.ModuleType("bilby")

    class Likelihood:
        def __init__(self, parameters=None):
            self.parameters = parameters or {}

    class Uniform:
        def __init__(self, minimum, maximum, name=None, *a, **kw):
            self.minimum, self.maximum = minimum, maximum

        def ln_prob(self, v):
            v = np.asarray(v, dtype=float)
            return np.where((v >= self.minimum) & (v <= self.maximum), -np.log(self.maximum - self.minimum), -np.inf)

    class PriorDict(dict):
        pass

    bilby.Likelihood = Likelihood
    bilby.core = types.SimpleNamespace(prior=types.SimpleNamespace(Uniform=Uniform, PriorDict=PriorDict, Prior=object),
                                       result=types.SimpleNamespace(Result=object))
    mods = {"emcee": emcee, "bilby": bilby, "bilby.core": types.ModuleType("bilby.core"),
            "bilby.core.sampler": types.ModuleType("bilby.core.sampler"),
            "bilby.core.sampler.emcee": types.ModuleType("bilby.core.sampler.emcee")}
    mods["bilby.core.sampler.emcee"].Emcee = type("Emcee", (), {"default_kwargs": {}})
    return mods


@pytest.fixture(scope="module")
def refpkg(tmp_path_factory):
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref not built")
    base = tmp_path_factory.mktemp("refpkg")
    pkg = base / "VegasAfterglow"
    pkg.mkdir()
    for name in os.listdir(REF_PKG):
        if name not in ("fitting", "__pycache__"):
            os.symlink(os.path.join(REF_PKG, name), pkg / name)
    (pkg / "fitting").mkdir()
    for name in os.listdir(os.path.join(REF_PKG, "fitting")):
        if name.endswith(".py") and name != "samplers.py":
            os.symlink(os.path.join(REF_PKG, "fitting", name), pkg / "fitting" / name)
    shutil.copy(os.path.join(REF_PKG, "fitting", "samplers.py"), pkg / "fitting" / "samplers.py")
    so = [f for f in os.listdir(os.path.join(ROOT, "oracle", "_ref")) if f.startswith("VegasAfterglowC")][0]
    os.symlink(os.path.join(ROOT, "oracle", "_ref", so), pkg / so)
    # the shippable artefact: apply the patch with the stock `patch` tool
    out = subprocess.run(["patch", "-p1", "-d", str(base), "-i", os.path.join(ROOT, "integration", "samplers_gpu.patch")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    saved = {k: sys.modules.get(k) for k in list(_stub_samplers())}
    sys.modules.update(_stub_samplers())
    sys.path.insert(0, str(base))
    for k in [k for k in sys.modules if k == "VegasAfterglow" or k.startswith("VegasAfterglow.")]:
        del sys.modules[k]
    if ref._pymod is not None:  # the pybind module may be loaded only once per process: reuse an earlier load
        sys.modules["VegasAfterglow.VegasAfterglowC"] = ref._pymod
    import VegasAfterglow.fitting as fitting_mod

    ref._pymod = sys.modules["VegasAfterglow.VegasAfterglowC"]
    yield fitting_mod
    sys.path.remove(str(base))
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


class ReferenceBackedEngine:
    """Stand-in for vegasafterglow_b200.engine.Engine: same entry points, answered by the unmodified reference."""

    def __init__(self):
        self.calls = []

    def chi2_series(self, P, t, nu, lnF, sig, w):
        from oracle import ref

        self.calls.append(("chi2_series", len(P)))
        return ref.chi2_series(P, t, nu, lnF, sig, w, n_threads=4)

    def chi2(self, P, points, bands):
        self.calls.append(("chi2", len(P), len(bands)))
        out = self.chi2_series(P, *points) if points is not None else np.zeros(len(P))
        import VegasAfterglow as va  # the temp package of the `refpkg` fixture (same pybind module as oracle/_ref)

        for i, p in enumerate(P):  # band terms through the reference's own Model.flux
            jet = va.TophatJet(p["theta_c"], p["E_iso"], p["Gamma0"], duration=p["duration"])
            mdl = va.Model(jet, va.ISM(p["n_ism"]), va.Observer(p["lumi_dist"], p["z"], p["theta_obs"]),
                           va.Radiation(p["fwd"]["eps_e"], p["fwd"]["eps_B"], p["fwd"]["p"], p["fwd"]["xi_e"]))
            for b in bands:
                F = np.asarray(mdl.flux(b["t"], b["nu_min"], b["nu_max"], b["num_nu"]).total)
                out[i] += np.sum(b["w"] * ((b["lnF_obs"] - np.log(np.maximum(F, 1e-300))) / b["sigma_ln"]) ** 2)
        return out


def _fitter(fm, with_band):
    from VegasAfterglow import ParamDef, Scale

    f = fm.Fitter(z=0.1, lumi_dist=1e27, jet="tophat", medium="ism")
    t = np.logspace(3, 6, 6)
    rng = np.random.default_rng(3)
    for nu, amp in ((1e9, 1e-27), (4.84e14, 3e-28)):
        f.add_flux_density(nu, t, amp * (t / 1e4) ** -1.0 * (1 + 0.05 * rng.standard_normal(t.size)), amp * 0.1 * np.ones(t.size))
    if with_band:
        tb = np.array([3e3, 5e4])
        f.add_flux((7.25e16, 2.4e18), tb, np.array([2e-12, 1e-13]), np.array([2e-13, 2e-14]), num_points=5)
    defs = [ParamDef("E_iso", 1e51, 1e54, Scale.log), ParamDef("Gamma0", 50, 800, Scale.log),
            ParamDef("theta_c", 0.03, 0.4, Scale.linear), ParamDef("n_ism", 1e-3, 10, Scale.log),
            ParamDef("eps_e", 1e-2, 0.5, Scale.log), ParamDef("eps_B", 1e-4, 0.1, Scale.log),
            ParamDef("p", 2.1, 2.8, Scale.linear), ParamDef("theta_v", 0.0, 0.0, Scale.fixed)]
    return f, defs


@pytest.mark.parametrize("with_band", [False, True])
def test_patched_fit_emcee_equals_the_reference_thread_pool(refpkg, with_band):
    from vegasafterglow_b200.integration import GpuFitterEvaluator

    f_cpu, defs = _fitter(refpkg, with_band)
    f_gpu, _ = _fitter(refpkg, with_band)
    f_gpu.gpu_engine = ReferenceBackedEngine()
    kw = dict(sampler="emcee", nwalkers=12, nsteps=3, nburn=0, npool=2, top_k=3)
    np.random.seed(7)  # generate_initial_positions draws the walker seeds from numpy's global generator
    r_cpu = f_cpu.fit(defs, **kw)
    np.random.seed(7)
    r_gpu = f_gpu.fit(defs, **kw)
    assert f_gpu.gpu_engine.calls, "the patched sampler did not go through the GPU engine"
    assert all(c[1] == 12 for c in f_gpu.gpu_engine.calls if c[0] == ("chi2" if with_band else "chi2_series"))
    # same proposals (seeded stand-in sampler), same log-likelihoods: the GPU batch path reproduces the thread-pool path
    np.testing.assert_array_equal(r_cpu.samples, r_gpu.samples)
    np.testing.assert_allclose(r_gpu.log_probs, r_cpu.log_probs, rtol=1e-9)
    # and walker by walker against Fitter._evaluate
    ev = GpuFitterEvaluator(f_gpu, ReferenceBackedEngine())
    th = np.array([[52.3, 2.4, 0.12, -0.5, -1.2, -2.5, 2.35], [53.1, 2.1, 0.3, 0.5, -0.7, -3.0, 2.6]])
    want = np.array([f_cpu._evaluate(f_cpu._to_params(x)) for x in th])
    np.testing.assert_allclose(ev.chi2(th), want, rtol=1e-9)


def test_walkers_the_constructor_rejects_get_minus_inf(refpkg):
    from vegasafterglow_b200.integration import GpuFitterEvaluator, log_prob_batch_gpu

    f, defs = _fitter(refpkg, False)
    f.fit(defs, sampler="emcee", nwalkers=8, nsteps=1, nburn=0, npool=1, top_k=1)
    ev = GpuFitterEvaluator(f, ReferenceBackedEngine())
    th = np.array([[52.3, 2.4, 0.12, -0.5, -1.2, -2.5, 2.35], [52.3, 2.4, 0.12, -0.5, 0.5, -2.5, 2.35]])  # eps_e = 10^0.5 > 1
    chi2 = ev.chi2(th)
    assert np.isfinite(chi2[0]) and np.isinf(chi2[1])
    prior = {n: types.SimpleNamespace(ln_prob=lambda v: np.zeros(len(v))) for n in "abcdefg"}
    lp = log_prob_batch_gpu(ev, list("abcdefg"), np.full(7, -100.0), np.full(7, 100.0), prior, lambda c: -0.5 * c)(th)
    assert np.isfinite(lp[0]) and lp[1] == -np.inf


def test_map_pool_batches_the_queue():
    from vegasafterglow_b200.integration import GpuMapPool

    seen = []

    class Ev:
        def log_likelihoods(self, arr, fn):
            seen.append(arr.shape)
            return -arr.sum(axis=1)

    pool = GpuMapPool(Ev(), lambda c: -0.5 * c, ndim=3)
    out = pool.map(lambda x: 123.0, [np.array([1.0, 2, 3]), np.array([4.0, 5, 6])])
    assert out == [-6.0, -15.0] and seen == [(2, 3)]
    assert pool.map(lambda x: 123.0, ["not-a-vector"]) == [123.0]
