"""GPU box: per-fixture parity table of the CUDA path (through the C ABI) against the committed reference fixtures,
plus bit-equality of the device grid (theta / phi nodes) against the host build of the same source and, where
oracle/_ref travelled, against the unmodified reference.  Writes profiles/parity_r02.json (or argv[1])."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from tests.helpers import golden_names, load_golden, model_errors
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine
from oracle.hostemu import emu

try:
    from oracle import ref
    HAVE_REF = ref.available()
except Exception:
    HAVE_REF = False

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "parity_r02.json")
eng = Engine(0)
rows = {}


def ulps(a, b):
    return int(np.abs(a.view(np.int64) - b.view(np.int64)).max()) if a.size else 0


for name in golden_names():
    g = load_golden(name)
    P, t, nu = g["params"], g["t"], g["nu"]
    fn = eng.flux_density_series if bool(g["series"]) else eng.flux_density_grid
    f, st = fn(P, t, nu, return_status=True)
    row = {"n_models": int(P.size), "status_or": int(np.bitwise_or.reduce(st)) if st.size else 0, "components": {}}
    for comp, cname in enumerate(abi.COMPONENTS):
        if not np.any(g["flux"][:, comp] > 0):
            continue
        err = model_errors(f, g["flux"], comp)
        spread = model_errors(g["flux_alt"], g["flux"], comp)
        row["components"][cname] = {"median": float(np.median(err)), "max": float(err.max()),
                                    "n_over_1e-6": int((err > 1e-6).sum()),
                                    "reference_cross_build_max": float(spread.max())}
    # grid bit-equality on up to 8 models of the fixture
    gb = {"models": 0, "theta_ulps_vs_host": 0, "phi_ulps_vs_host": 0, "t_ulps_vs_host": 0, "size_mismatch": 0}
    if HAVE_REF:
        gb.update({"theta_ulps_vs_reference": 0, "phi_ulps_vs_reference": 0})
    for i in range(min(8, P.size)):
        p = P[i:i + 1]
        d = eng.details(p, float(t.min()), float(t.max()))
        e = emu.details(p, float(t.min()), float(t.max()))
        gb["models"] += 1
        if d["theta"].shape != e["theta"].shape or d["phi"].shape != e["phi"].shape or d["t_rows"].shape != e["t_rows"].shape:
            gb["size_mismatch"] += 1
            continue
        gb["theta_ulps_vs_host"] = max(gb["theta_ulps_vs_host"], ulps(d["theta"], e["theta"]))
        gb["phi_ulps_vs_host"] = max(gb["phi_ulps_vs_host"], ulps(d["phi"], e["phi"]))
        gb["t_ulps_vs_host"] = max(gb["t_ulps_vs_host"], ulps(d["t_rows"].ravel(), e["t_rows"].ravel()))
        if HAVE_REF:
            r = ref.details(p, float(t.min()), float(t.max()))
            if r["theta"].shape == d["theta"].shape and r["phi"].shape == d["phi"].shape:
                gb["theta_ulps_vs_reference"] = max(gb["theta_ulps_vs_reference"], ulps(d["theta"], r["theta"]))
                gb["phi_ulps_vs_reference"] = max(gb["phi_ulps_vs_reference"], ulps(d["phi"], r["phi"]))
            else:
                gb["size_mismatch"] += 1
    row["grid"] = gb
    rows[name] = row
    c = row["components"]
    print(f"{name:38s} " + " | ".join(f"{k} med {v['median']:.1e} max {v['max']:.1e}" for k, v in c.items()) +
          f" | grid ulps host {gb['theta_ulps_vs_host']}/{gb['phi_ulps_vs_host']}/{gb['t_ulps_vs_host']}"
          + (f" ref {gb['theta_ulps_vs_reference']}/{gb['phi_ulps_vs_reference']}" if HAVE_REF else ""), flush=True)

os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump({"what": "GPU (C ABI) vs reference fixtures: per-model max relative flux error over bins above 1 % of the band peak; "
                   "grid: worst ulp distance of the device theta / phi / t nodes to the host build of the same source and to "
                   "the unmodified reference", "fixtures": rows}, open(out_path, "w"), indent=1)
print("wrote", out_path)
