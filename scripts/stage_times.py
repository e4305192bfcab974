"""Per-stage device times (CUDA events inside the library) of the two bench workloads, un-overlapped.
usage: python scripts/stage_times.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = Engine(0)
dev = torch.device("cuda:0")
eng.set_capacity(256, 128)
eng.set_profiling(True)

def avg(fn, n=6):
    acc = {}
    for i in range(n + 2):
        fn()
        torch.cuda.synchronize()
        if i >= 2:
            for k, v in eng.last_stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / n
    return {k: round(v, 3) for k, v in acc.items()}

P, t, nu = bench.workload(batch)
d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
d_t, d_nu = torch.from_numpy(t).to(dev), torch.from_numpy(nu).to(dev)
d_out = torch.empty((P.size, abi.NCOMP, nu.size, t.size), dtype=torch.float64, device=dev)
r = avg(lambda: eng.flux_density_grid_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), t.size, d_nu.data_ptr(), nu.size, d_out.data_ptr()))
print("FS grid  ", batch, r, "total", round(sum(r.values()), 3))

P, ts, nus = bench.loglike_workload(batch)
d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
d_t, d_nu = torch.from_numpy(ts).to(dev), torch.from_numpy(nus).to(dev)
d_out = torch.empty((P.size, abi.NCOMP, ts.size), dtype=torch.float64, device=dev)
r = avg(lambda: eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr()))
print("RS series", batch, r, "total", round(sum(r.values()), 3))
