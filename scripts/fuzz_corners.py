"""Corner-range fuzz on the GPU (one-off robustness probe): parameters drawn from the edges of the physical
ranges the reference validates (pybind/pymodel.cpp:47-186), wide observation windows, series and grid mode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vegasafterglow_b200 import abi, configs
from vegasafterglow_b200.engine import Engine
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(77)
P = []
for i in range(n):
    jet = ["tophat", "gaussian", "powerlaw"][rng.integers(3)]
    med = ["ism", "wind"][rng.integers(2)]
    rvs = rng.random() < 0.4
    p = configs.make(jet=jet, medium=med)
    p["E_iso"] = 10 ** rng.uniform(48, 56)
    p["Gamma0"] = 1 + 10 ** rng.uniform(-0.5, 3.5)          # 1.3 .. 3000
    p["theta_c"] = 10 ** rng.uniform(-2.5, np.log10(1.57))
    p["n_ism"] = 10 ** rng.uniform(-6, 5) if med == "ism" else 0.0
    p["A_star"] = 10 ** rng.uniform(-4, 2)
    p["theta_obs"] = rng.uniform(0, 1.5) if rng.random() < 0.7 else 0.0
    p["z"] = 10 ** rng.uniform(-3, 1)
    p["lumi_dist"] = 10 ** rng.uniform(25, 29.5)
    for k in ("fwd", "rvs"):
        p[k]["eps_e"] = 10 ** rng.uniform(-4, 0)
        p[k]["eps_B"] = 10 ** rng.uniform(-8, 0)
        p[k]["p"] = rng.uniform(1.2, 3.8)
        p[k]["xi_e"] = 10 ** rng.uniform(-3, 0)
    if rvs:
        p["has_rvs"] = 1
        p["duration"] = 10 ** rng.uniform(-1, 5)
    if jet == "powerlaw":
        p["k_e"], p["k_g"] = rng.uniform(0.5, 8), rng.uniform(0.5, 8)
    if rng.random() < 0.2:
        p["spreading"] = 1
    if rng.random() < 0.3:
        p["phi_resol"], p["theta_resol"], p["t_resol"] = rng.uniform(0.02, 0.3), rng.uniform(0.05, 1.2), rng.uniform(2, 15)
    P.append(p)
P = np.concatenate(P)
eng = Engine(0)
for label, (t, nu) in {"wide grid": (np.logspace(-2, 10, 50), np.array([1e6, 1e10, 1e15, 1e19, 1e24])),
                        "late series": (np.sort(10 ** rng.uniform(3, 9, 64)), 10 ** rng.uniform(8, 20, 64))}.items():
    series = t.size == nu.size and label.endswith("series")
    t0 = time.time()
    fn = eng.flux_density_series if series else eng.flux_density_grid
    flux, st = fn(P, t, nu, return_status=True)
    bad = ~np.isfinite(flux).reshape(n, -1).all(axis=1)
    print(f"{label}: {n} models in {time.time() - t0:.2f} s; status bits { {int(b): int(np.count_nonzero(st & b)) for b in (1, 2, 4, 8, 16, 32)} }; "
          f"non-finite models {np.count_nonzero(bad)}; negative flux entries {np.count_nonzero(flux < 0)}")
    idx = rng.choice(n, size=40, replace=False)
    rfn = ref.flux_density_series if series else ref.flux_density_grid
    r = rfn(P[idx], t, nu, n_threads=ref.hardware_threads())
    errs = []
    for q, i in enumerate(idx):
        b, a = r[q, 0], flux[i, 0]
        if not np.isfinite(b).all():
            errs.append(np.nan)
            continue
        m = b > 1e-2 * b.max()
        errs.append(np.max(np.abs(a[m] - b[m]) / b[m]) if m.any() else 0.0)
    errs = np.array(errs)
    ok = np.isfinite(errs)
    print(f"  parity on 40 sampled: reference non-finite for {np.count_nonzero(~ok)}; median {np.median(errs[ok]):.2e}, 90% {np.percentile(errs[ok], 90):.2e}, max {errs[ok].max():.2e}")
    for q in np.argsort(-np.nan_to_num(errs))[:4]:
        i = idx[q]
        print(f"    err {errs[q]:.2e} model {i}: jet {int(P['jet_type'][i])} med {int(P['medium_type'][i])} rvs {int(P['has_rvs'][i])} spread {int(P['spreading'][i])} G0 {P['Gamma0'][i]:.3g} th_c {P['theta_c'][i]:.3g} th_obs {P['theta_obs'][i]:.3g} p {P['fwd']['p'][i]:.2f} status {st[i]}")
    for i in np.nonzero(bad)[0][:5]:
        print(f"    NON-FINITE model {i}: jet {int(P['jet_type'][i])} med {int(P['medium_type'][i])} rvs {int(P['has_rvs'][i])} G0 {P['Gamma0'][i]:.3g} th_c {P['theta_c'][i]:.3g} th_obs {P['theta_obs'][i]:.3g} status {st[i]}")
