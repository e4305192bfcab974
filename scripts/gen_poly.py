"""Generates the polynomial coefficients of vag::dexp2 / vag::dlog2 (vag_math.cuh): Chebyshev-node
interpolation in extended precision, printed as C hex-float literals, with the measured max error."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

ld = np.longdouble


def cheb_fit(fun, a, b, deg):
    k = np.arange(deg + 1, dtype=ld)
    x = np.cos(np.pi * (k + ld(0.5)) / (deg + 1))  # Chebyshev nodes on [-1,1]
    xx = (ld(0.5) * (b - a)) * x + ld(0.5) * (b + a)
    y = fun(xx)
    # solve Vandermonde in monomial basis of t=(x) scaled -> convert to monomials in xx
    V = np.vander(np.asarray(xx, dtype=ld), deg + 1, increasing=True)
    coef = np.linalg.solve(V.astype(np.float64), np.asarray(y, dtype=np.float64))  # float64 solve then refine
    # iterative refinement in long double
    c = coef.astype(ld)
    for _ in range(5):
        r = y - V @ c
        c = c + np.linalg.solve(V.astype(np.float64), np.asarray(r, dtype=np.float64)).astype(ld)
    return c


def horner(c, x):
    p = np.zeros_like(x) + c[-1]
    for ck in c[-2::-1]:
        p = p * x + ck
    return p


ln2 = np.log(ld(2))
# exp2 on [-0.5, 0.5]
for deg in (9, 10, 11, 12):
    c = cheb_fit(lambda f: np.exp(f * ln2), ld(-0.5), ld(0.5), deg)
    xs = np.linspace(-0.5, 0.5, 200001).astype(ld)
    cd = c.astype(np.float64)
    err = np.max(np.abs(horner(cd.astype(ld), xs) / np.exp(xs * ln2) - 1))
    print("exp2 deg", deg, "max rel err", float(err))
    if deg == 10:  # the degree vag_math.cuh uses (VAG_EXP2_DEG)
        print("EXP2 = {" + ", ".join(float(v).hex() for v in cd) + "};")
# log2(m) = s * q(s^2), s=(m-1)/(m+1), m in [sqrt(.5), sqrt(2)) -> z = s^2 in [0, 0.02944]
smax = (np.sqrt(ld(2)) - 1) / (np.sqrt(ld(2)) + 1)
zmax = smax * smax


def q(z):
    s = np.sqrt(z)
    out = np.where(z > 0, np.log((1 + s) / (1 - s)) / ln2 / np.where(s > 0, s, 1), 2 / ln2)
    return out


for deg in (8, 9, 10):
    c = cheb_fit(q, ld(0), zmax * ld(1.0001), deg)
    zs = np.linspace(1e-12, float(zmax), 200001).astype(ld)
    cd = c.astype(np.float64)
    err = np.max(np.abs(horner(cd.astype(ld), zs) / q(zs) - 1))
    print("log2 deg", deg, "max rel err", float(err))
    if deg == 9:
        print("LOG2 = {" + ", ".join(float(v).hex() for v in cd) + "};")
