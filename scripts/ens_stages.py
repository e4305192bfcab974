"""Stage times of one config-5 ensemble share: scripts/ens_stages.py [walkers ...] (device-resident series chi2-less flux call)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch, bench
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine
eng = Engine(0); dev = torch.device('cuda:0'); eng.set_capacity(256, 128); eng.set_profiling(True)
for n in [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096]:
    P, ts, nus = bench.loglike_workload(n)
    d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev); d_t, d_nu = torch.from_numpy(ts).to(dev), torch.from_numpy(nus).to(dev)
    d_out = torch.empty((P.size, abi.NCOMP, ts.size), dtype=torch.float64, device=dev)
    acc = {}
    for it in range(8):
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr()); eng.synchronize()
        if it >= 3:
            for k, v in eng.last_stage_ms().items(): acc[k] = acc.get(k, 0) + v / 5
    t0 = time.perf_counter()
    for _ in range(5):
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr()); eng.synchronize()
    print(n, 'walkers: wall %.3f ms' % ((time.perf_counter() - t0) / 5 * 1e3), {k: round(v, 3) for k, v in acc.items()}, 'sum %.3f' % sum(acc.values()))
