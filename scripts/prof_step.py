"""Short profiling target: a few device-resident steps of the bench workload (FS grid batch and the
config-5 FS+RS series batch) -- run under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vegasafterglow_b200 import abi
from vegasafterglow_b200.engine import Engine

which = sys.argv[1] if len(sys.argv) > 1 else "fs"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
eng = Engine(0)
dev = torch.device("cuda:0")
eng.set_capacity(256, 128)
if which == "fs":
    P, t, nu = bench.workload(batch)
    d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
    d_t, d_nu = torch.from_numpy(t).to(dev), torch.from_numpy(nu).to(dev)
    d_out = torch.empty((P.size, abi.NCOMP, nu.size, t.size), dtype=torch.float64, device=dev)
    for _ in range(steps):
        eng.flux_density_grid_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), t.size, d_nu.data_ptr(), nu.size, d_out.data_ptr())
else:
    P, ts, nus = bench.loglike_workload(batch)
    d_p = torch.from_numpy(P.view(np.uint8).copy()).to(dev)
    d_t, d_nu = torch.from_numpy(ts).to(dev), torch.from_numpy(nus).to(dev)
    d_out = torch.empty((P.size, abi.NCOMP, ts.size), dtype=torch.float64, device=dev)
    for _ in range(steps):
        eng.flux_density_series_dev(d_p.data_ptr(), P.size, d_t.data_ptr(), d_nu.data_ptr(), ts.size, d_out.data_ptr())
eng.synchronize()
print("done", which, steps, batch)
