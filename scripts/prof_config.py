"""One of the BASELINE.json side configurations of bench.extra_workloads, a few device-resident steps -- run under ncu:
   python scripts/prof_config.py C4_powerlaw_ssc_kn_50x40 [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P, t, nu, series = bench.extra_workloads()[name]
eng = Engine(0)
dev = torch.device("cuda:0")
eng.set_capacity(256, 128)
d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
d_t, d_nu = torch.from_numpy(t).to(dev), torch.from_numpy(nu).to(dev)
shape = (P.size, abi.NCOMP, t.size) if series else (P.size, abi.NCOMP, nu.size, t.size)
d_out = torch.empty(shape, dtype=torch.float64, device=dev)
for _ in range(steps):
    if series:
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), t.size, d_out.data_ptr())
    else:
        eng.flux_density_grid_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), t.size, d_nu.data_ptr(), nu.size, d_out.data_ptr())
eng.synchronize()
print("done", name, steps, P.size)
