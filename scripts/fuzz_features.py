"""Feature-combination fuzz on the GPU (not a test: a one-off robustness probe).  Random parameter sets over
jet type x medium x reverse shock x SSC x spreading x magnetar x wind slope x axisymmetric x magnetisation;
checks status words / finiteness for every model and parity against the unmodified reference for a sample."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vegasafterglow_b200 import abi, configs
from vegasafterglow_b200.engine import Engine
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(2026)
jets = ["tophat", "gaussian", "powerlaw", "two_component", "step_powerlaw", "powerlaw_wing"]
P = []
for i in range(n):
    jet = jets[rng.integers(len(jets))]
    med = ["ism", "wind", "wind_ism"][rng.integers(3)]
    rvs = rng.random() < 0.4
    p = configs.random_draw(1, seed=10000 + i, rvs=rvs, jet=jet, medium="ism" if med == "ism" else "wind",
                            theta_obs_max=0.4 if rng.random() < 0.6 else 0.0)
    if med == "wind_ism":
        p["n_ism"] = 10 ** rng.uniform(-3, 0)
    if med != "ism" and rng.random() < 0.4:
        p["wind_k_m"] = rng.uniform(0.5, 2.9)
    if med != "ism" and rng.random() < 0.3:
        p["n0"] = 10 ** rng.uniform(2, 6)
    if jet in ("two_component", "step_powerlaw", "powerlaw_wing"):
        p["theta_c"] = rng.uniform(0.03, 0.1)
        p["theta_w"] = p["theta_c"] * rng.uniform(2, 5)
        p["E_iso_w"] = p["E_iso"] * 10 ** rng.uniform(-3, -0.5)
        p["Gamma0_w"] = np.maximum(p["Gamma0"] * rng.uniform(0.1, 0.6), 3.0)
        p["k_e"], p["k_g"] = rng.uniform(1, 4), rng.uniform(1, 3)
    if jet == "powerlaw":
        p["k_e"], p["k_g"] = rng.uniform(1, 4), rng.uniform(1, 3)
    u_sw = rng.random()
    if u_sw < 0.25:
        p["spreading"] = 1
    elif u_sw < 0.36:
        p["axisymmetric"] = 0
    elif u_sw < 0.44:  # both: one ODE row per (phi, theta) cell
        p["spreading"], p["axisymmetric"] = 1, 0
    if jet != "powerlaw_wing" and rng.random() < 0.25:
        p["has_magnetar"] = 1
        p["magnetar_L0"], p["magnetar_t0"], p["magnetar_q"] = 10 ** rng.uniform(45, 49.5), 10 ** rng.uniform(1.5, 4.5), rng.uniform(1, 3)
    if rvs and rng.random() < 0.25:
        p["sigma0"] = 10 ** rng.uniform(-2, 1)
    if rng.random() < 0.15:
        p["fwd"]["ssc"], p["fwd"]["kn"] = 1, int(rng.random() < 0.5)
    if rng.random() < 0.1:
        p["radiative_fireball"] = 0
    P.append(p)
P = np.concatenate(P)
t, nu = np.logspace(1.5, 7.5, 40), np.array([1e9, 1e14, 1e17, 1e22])
eng = Engine(0)
t0 = time.time()
flux, st = eng.flux_density_grid(P, t, nu, return_status=True)
print(f"{n} mixed models in {time.time() - t0:.2f} s; status bits set: {np.count_nonzero(st)} "
      f"({ {int(b): int(np.count_nonzero(st & b)) for b in (1, 2, 4, 8, 16, 32)} }); non-finite models: "
      f"{np.count_nonzero(~np.isfinite(flux).all(axis=(1, 2, 3)))}")
bad = np.nonzero((st != 0) | ~np.isfinite(flux).all(axis=(1, 2, 3)))[0]
for i in bad[:10]:
    print("  model", i, "status", st[i], {k: P[k][i] for k in ("jet_type", "medium_type", "has_rvs", "spreading", "has_magnetar", "sigma0", "wind_k_m", "theta_c", "Gamma0")})
idx = rng.choice(n, size=min(n_ref, n), replace=False)
r = ref.flux_density_grid(P[idx], t, nu, n_threads=ref.hardware_threads())
errs = []
for q, i in enumerate(idx):
    b, a = r[q, 0], flux[i, 0]
    m = b > 1e-2 * b.max(axis=-1, keepdims=True)
    errs.append(np.max(np.abs(a[m] - b[m]) / b[m]) if m.any() else 0.0)
errs = np.array(errs)
print(f"parity vs the unmodified reference on {idx.size} sampled models: median {np.median(errs):.2e}, "
      f"90% {np.percentile(errs, 90):.2e}, max {errs.max():.2e}; > 1e-6: {np.count_nonzero(errs > 1e-6)}, > 1e-3: {np.count_nonzero(errs > 1e-3)}")
fs_only = np.array([P["has_rvs"][i] == 0 for i in idx])
print(f"  forward-shock-only models: {fs_only.sum()}, of which > 1e-6: {np.count_nonzero(errs[fs_only] > 1e-6)}, max {errs[fs_only].max() if fs_only.any() else 0:.2e}")
for q in np.argsort(-errs)[:8]:
    i = idx[q]
    print(f"  err {errs[q]:.2e} model {i}:", {k: (float(P[k][i]) if P[k][i].dtype.kind == 'f' else int(P[k][i])) for k in ("jet_type", "medium_type", "has_rvs", "spreading", "axisymmetric", "has_magnetar", "sigma0", "wind_k_m", "theta_obs")}, "ssc", int(P["fwd"]["ssc"][i]))
