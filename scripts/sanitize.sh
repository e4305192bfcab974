#!/bin/bash
# compute-sanitizer over the kernel set: memcheck and racecheck on one forward-shock, one reverse-shock, one SSC, one
# spreading batch, a structured off-axis batch and a < 148-model batch (row-split path).  Run on the GPU box:
#   bash scripts/sanitize.sh  ->  gpurun_out/sanitize_<tool>.log (copy the summaries to profiles/)
set -u
mkdir -p gpurun_out
cat > /tmp/sanitize_driver.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from vegasafterglow_b200 import configs
from vegasafterglow_b200.engine import Engine
eng = Engine(0)
t, nu = np.logspace(2, 7, 24), np.array([1e9, 1e14, 1e17])
def run(name, P, tt=t, nn=nu, series=False):
    f, st = (eng.flux_density_series if series else eng.flux_density_grid)(P, tt, nn, return_status=True)
    assert np.isfinite(f).all() and (st == 0).all(), name
    print(name, "ok", f.shape, flush=True)
run("fs tophat x200 (8 lanes/model grid path needs >= 8192: 16-lane path here)", configs.random_draw(200, seed=1))
run("fs tophat x3 (row-split slabs)", configs.random_draw(3, seed=2))
run("rs tophat x160", configs.random_draw(160, seed=3, rvs=True))
run("rs series x40", configs.random_draw(40, seed=4, rvs=True), np.sort(np.tile(np.logspace(3, 6, 8), 2)), np.tile([1e9, 1e17], 8), True)
eng.set_series_mode(2)  # banded series (k_series_bands + the (node, band) tile), forced
run("rs series x40, banded", configs.random_draw(40, seed=4, rvs=True), np.sort(np.tile(np.logspace(3, 6, 8), 2)), np.tile([1e9, 1e17], 8), True)
run("ssc series x6, banded", configs.random_draw(6, seed=12, ssc=True, kn=True), np.sort(np.tile(np.logspace(3, 6, 6), 3)), np.tile([1e9, 1e17, 1e24], 6), True)
run("series of 9 frequencies (not banded)", configs.random_draw(6, seed=13), np.logspace(3, 6, 9), np.logspace(9, 18, 9), True)
eng.set_series_mode(0)
run("gaussian off-axis x12", configs.random_draw(12, seed=5, jet="gaussian", theta_obs_max=0.4))
run("powerlaw wind rs x6", configs.random_draw(6, seed=6, jet="powerlaw", medium="wind", rvs=True, theta_obs_max=0.3))
P = configs.random_draw(6, seed=7, theta_obs_max=0.3); P["spreading"] = 1
run("spreading x6", P)
P = configs.random_draw(4, seed=15, theta_obs_max=0.3); P["spreading"] = 1; P["axisymmetric"] = 0
run("spreading, axisymmetric=False x4 (one ODE row per (phi, theta) cell)", P)
run("ssc kn x6", configs.random_draw(6, seed=8, ssc=True, kn=True), t, np.array([1e9, 1e17, 1e24]))
P = configs.random_draw(1100, seed=9); run("fs tophat x1100 (16 lanes / model)", P, t[:6], nu[:1])
P = configs.random_draw(8300, seed=10); run("fs tophat x8300 (8 lanes / model)", P, t[:3], nu[:1])
ts = np.sort(np.tile(np.logspace(3, 6, 6), 2)); nus = np.tile([1e9, 1e17], 6)
c = eng.chi2(configs.random_draw(20, seed=11, rvs=True), (ts, nus, np.full(12, -60.0), np.full(12, 0.1), np.ones(12)),
             [dict(t=np.array([1e4, 1e5]), lnF_obs=np.array([-30.0, -32.0]), sigma_ln=np.array([0.1, 0.1]), w=np.ones(2), nu_min=1e17, nu_max=1e18, num_nu=5)])
assert np.isfinite(c).all(); print("chi2 with band ok", flush=True)
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit |ok" gpurun_out/sanitize_$tool.log | tail -20
done
