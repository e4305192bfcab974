import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, bench
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine
eng = Engine(0); dev = torch.device('cuda:0'); eng.set_capacity(256, 128)
P, ts, nus = bench.loglike_workload(4096)
d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev); d_t, d_nu = torch.from_numpy(ts).to(dev), torch.from_numpy(nus).to(dev)
d_out = torch.empty((P.size, abi.NCOMP, ts.size), dtype=torch.float64, device=dev)
eng.set_profiling(True) if hasattr(eng, 'set_profiling') else None
outs = {}
for mode in (1, 2, 0):
    eng.set_series_mode(mode)
    for _ in range(3):
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr()); eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr()); eng.synchronize()
    dt = (time.perf_counter() - t0) / 5
    outs[mode] = d_out.cpu().numpy().copy()
    print('series_mode', mode, 'ms/batch %.3f' % (dt * 1e3), 'stage ms', eng.last_stage_ms() if hasattr(eng, 'last_stage_ms') else '')
a, b = outs[1], outs[2]
m = a > 0
print('banded vs per-point max rel', np.max(np.abs(a[m] - b[m]) / a[m]))
