"""D2H probe for the end-to-end arm: strided (VAG_OUT_PRESENT: 2 of 5 component planes per model) against
contiguous copies of the same byte count, pinned host memory.  usage: python scripts/d2h_probe.py [n_models]"""
import sys, time
import torch
from cuda.bindings import runtime as rt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
row = 300 * 8                       # one component plane of one model [n_nu=3][n_t=100] doubles
pitch = 5 * row
dev = torch.device("cuda:0")
d = torch.zeros(n * pitch, dtype=torch.uint8, device=dev)
h = torch.zeros(n * pitch, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
st = s.cuda_stream
K = rt.cudaMemcpyKind.cudaMemcpyDeviceToHost

def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

def c_full():
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), n * pitch, K, st)
def c_2d():
    rt.cudaMemcpy2DAsync(h.data_ptr(), pitch, d.data_ptr(), pitch, 2 * row, n, K, st)
def c_contig():
    rt.cudaMemcpyAsync(h.data_ptr(), d.data_ptr(), n * 2 * row, K, st)
def c_2d_dev_only():   # contiguous host, strided device
    rt.cudaMemcpy2DAsync(h.data_ptr(), 2 * row, d.data_ptr(), pitch, 2 * row, n, K, st)

for name, fn, nbytes in (("dense 5 planes contiguous", c_full, n * pitch), ("2 planes strided both sides", c_2d, n * 2 * row),
                         ("2 planes contiguous", c_contig, n * 2 * row), ("2 planes strided device / packed host", c_2d_dev_only, n * 2 * row)):
    t = timed(fn)
    print(f"{name:40s} {nbytes / 1e6:8.1f} MB  {t * 1e3:7.3f} ms  {nbytes / t / 1e9:6.1f} GB/s", flush=True)
