"""First GPU contact: parity of every stage + rough timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vegasafterglow_b200 import configs
from vegasafterglow_b200.engine import Engine
from oracle import ref

eng = Engine(0)
def rel(a, b, floor=0.0):
    m = np.abs(b) > floor
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0

def cmp(name, p, t, nu):
    fr = ref.flux_density_grid(p, t, nu)
    t0 = time.time(); fg, st = eng.flux_density_grid(p, t, nu, return_status=True); dt = time.time() - t0
    out = []
    for c in (1, 3):
        b = fr[0, c]
        if not b.any(): continue
        peak = b.max(axis=1, keepdims=True)
        out.append((c, rel(fg[0, c], b), float(np.max(np.abs(fg[0,c]-b)[b > 1e-2*peak] / b[b > 1e-2*peak]))))
    d = eng.details(p, t[0], t[-1]); dr = ref.details(p, t[0], t[-1])
    print(name, 'status', st, 'gpu %.2f ms' % (dt*1e3), tuple(d['info'])[:5], tuple(dr['info'])[:5], out,
          'theta', rel(d['theta'], dr['theta']), 'Gamma', rel(d['fwd_shock'][3], dr['fwd_shock'][3]), flush=True)

cmp('C1', *configs.C1())
cmp('C2', *configs.C2())
cmp('C3', *configs.C3())
for g in configs.GOLDEN:
    cmp(g, configs.golden(g), configs.GOLDEN_T, configs.GOLDEN_NU)

# batch parity + timing
p, t, nu = configs.C1()
for n, kw in ((256, dict()), (256, dict(rvs=True)), (64, dict(jet='gaussian', theta_obs_max=0.4)), (64, dict(medium='wind', rvs=True))):
    P = configs.random_draw(n, seed=1, **kw)
    fr = ref.flux_density_grid(P, t, nu, n_threads=8)
    fg, st = eng.flux_density_grid(P, t, nu, return_status=True)
    errs = []
    for i in range(n):
        b = fr[i, 0]; m = b > 1e-2 * b.max(axis=1, keepdims=True)
        errs.append(np.max(np.abs(fg[i, 0] - b)[m] / b[m]))
    errs = np.array(errs)
    print('batch', n, kw, 'status', np.unique(st), 'max err', errs.max(), 'median', np.median(errs), 'n>1e-6', int((errs > 1e-6).sum()), flush=True)

eng.set_profiling(True)
for n in (1, 256, 4096, 32768):
    for kw in (dict(), dict(rvs=True)):
        P = configs.random_draw(n, seed=2, **kw)
        eng.flux_density_grid(P, t, nu)
        t0 = time.time(); eng.flux_density_grid(P, t, nu); dt = time.time() - t0
        print('timing n=%d %s: %.2f ms  %.0f evals/s' % (n, kw, dt*1e3, n/dt), eng.last_stage_ms(), flush=True)
ts = np.sort(np.tile(np.logspace(2.5, 6.5, 20), 5)); nus = np.tile([1e9, 5e9, 4.84e14, 1e17, 1e18], 20)
P = configs.random_draw(4096, seed=3, rvs=True)
eng.flux_density_series(P, ts, nus)
t0 = time.time(); eng.flux_density_series(P, ts, nus); dt = time.time() - t0
print('series 4096 rvs: %.2f ms %.0f evals/s' % (dt*1e3, 4096/dt), eng.last_stage_ms(), flush=True)
fr = ref.flux_density_series(P[:64], ts, nus, n_threads=8); fg = eng.flux_density_series(P[:64], ts, nus)
print('series parity', rel(fg[:, 0], fr[:, 0], 1e-300))
