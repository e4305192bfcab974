"""GPU box: per-case mismatch counts of the device build of csrc/vag_libm.cuh against the live host libm."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_libm_exact import cases, host_eval, mismatches
from vegasafterglow_b200.engine import Engine
eng = Engine(0)
for name, fn, x, y in cases(400_000, seed=7):
    got = eng.selftest_libm(fn, x, y); want = host_eval(fn, x, y, ref=True); hb = host_eval(fn, x, y)
    bad = mismatches(got, want); badh = mismatches(hb, want)
    msg = f"{name:34s} device-vs-libm {bad.size:7d}  hostbuild-vs-libm {badh.size}"
    if bad.size:
        i = bad[0]; msg += f"  x={float(x[i]).hex()} y={(float(y[i]).hex() if y is not None else '-')} got={got[i].hex()} want={want[i].hex()}"
    print(msg, flush=True)
