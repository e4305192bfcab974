"""TEST INFRASTRUCTURE ONLY -- an independent numpy / pure-Python restatement of the reference's physics stages
for forward-shock synchrotron models: blast-wave ODE, synchrotron
electrons and photons, equal-arrival-time-surface flux integration.  It shares no code with the product
(`vegasafterglow_b200/csrc`): different language, different structure (row-vectorised numpy, scalar dopri5).

Scope: TophatJet / GaussianJet / PowerLawJet, ISM / Wind(k_m = 2), forward shock and forward + reverse shock
(unmagnetised), no inverse Compton, no spreading, axisymmetric.  `flux_density_grid` takes the adaptive
(phi, theta, t) grid as an input (e.g. the reference's own `details()` tables), which isolates the physics stages;
`auto_grid` restates the grid builder (row a1) and `model_flux` chains the two into the whole path.

Pinned by tests/test_restatement.py against the unmodified reference (oracle/_ref): physics stages on the
reference's grid <= 2e-10 (C1, Gaussian off-axis, power-law/wind, C3, config-5 truth model, thin shell); restated
grid = identical node counts / symmetry groups, nodes to the quadrature-noise level (2e-11 structured jets, 2e-6
tophat); whole path 2.6e-7 (C1), 1e-11 (C2), 4e-9 (C3), 5e-13 (power-law/wind).  Citations are `path:line` in the
reference repository.
"""
from __future__ import annotations

import math

import numpy as np

# ---- unit system and constants: src/util/macros.h:43-107 -----------------------------------------------------------
LEN = 1.5e13
CM = 1 / LEN
SEC = 3e10 / LEN
G = 1 / 2e33
HZ = 1 / SEC
ERG = G * CM * CM / SEC / SEC
FLUX_DEN_CGS = ERG / (CM * CM) / SEC / HZ
C = 1.0
MP = 1.67e-24 * G
ME = MP / 1836
E = 4.8e-10 / 4.472136e16 / 5.809475e19 / SEC
SIGMA_T = 6.65e-25 * CM * CM
PI = math.pi
GAMMA_CUT = 1 + 1e-6          # src/config/simulation-defaults.h:41
LN2 = math.log(2.0)
LOG2E = 1 / LN2


# ---- jets and media: src/environment/jet.h:84-257, medium.h:50-140 -------------------------------------------------
class Model:
    def __init__(self, p):
        g = lambda k: float(np.asarray(p[k]).reshape(-1)[0])  # noqa: E731
        self.jet, self.theta_c = int(g("jet_type")), g("theta_c")
        self.eps_k0 = g("E_iso") * ERG / (4 * PI)
        self.Gamma0, self.k_e, self.k_g = g("Gamma0"), g("k_e"), g("k_g")
        self.ism = int(g("medium_type")) == 0
        self.rho_ism = g("n_ism") / CM**3 * MP
        self.wind_A = g("A_star") * 5e11 * G / CM if not self.ism else 0.0
        n0 = g("n0")
        self.wind_r02 = self.wind_A / ((n0 / CM**3) * 1.3 * MP) if (not self.ism and math.isfinite(n0)) else 0.0
        f = np.asarray(p["fwd"]).reshape(-1)[0]
        self.eps_e, self.eps_B, self.p, self.xi_e = (float(f[k]) for k in ("eps_e", "eps_B", "p", "xi_e"))
        self.radiative = bool(g("radiative_fireball"))
        self.z, self.d_L, self.theta_v = g("z"), g("lumi_dist") * CM, g("theta_obs")
        rtol = g("rtol")
        self.rtol = rtol if rtol > 0 else 1e-6
        # RadiativeEfficiency coefficients: src/dynamics/shock-physics.h:251-261
        self.gm_coeff = (self.p - 2) / (self.p - 1) * self.eps_e * MP / ME / self.xi_e
        self.gc_coeff = 6 * PI * ME * C / SIGMA_T / (8 * PI * self.eps_B)

    def Gamma0_of(self, th):
        if self.jet == 0:
            return self.Gamma0 if th < self.theta_c else 1.0
        if self.jet == 1:
            return (self.Gamma0 - 1) * math.exp(th * th * (-1 / (2 * self.theta_c**2))) + 1
        return (self.Gamma0 - 1) / (1 + 2.0 ** (self.k_g * math.log2(th / self.theta_c))) + 1

    def eps_k_of(self, th):
        if self.jet == 0:
            return self.eps_k0 if th < self.theta_c else 0.0
        if self.jet == 1:
            return self.eps_k0 * math.exp(th * th * (-1 / (2 * self.theta_c**2)))
        return self.eps_k0 / (1 + 2.0 ** (self.k_e * math.log2(th / self.theta_c)))

    def rho(self, r):
        return self.rho_ism if self.ism else self.wind_A / (self.wind_r02 + r * r) + self.rho_ism

    def mass(self, r):  # enclosed mass per solid angle: medium.h:61,115-127
        m = self.rho_ism * r**3 / 3.0
        if not self.ism and self.wind_A != 0:
            if self.wind_r02 > 0:
                a = math.sqrt(self.wind_r02)
                m += self.wind_A * (r - a * math.atan(r / a))
            else:
                m += self.wind_A * r
        return m


def adiabatic_idx(g):  # src/core/physics.h:59-61
    return 4.0 / 3.0 + 1 / (3 * g)


def simpson_logspace(f, r):  # shock-physics.h:401-416
    n, u1 = 32, math.log(r)
    u0 = u1 - 18
    h = (u1 - u0) / n
    s = f(u0) + f(u1)
    s += sum(4 * f(u0 + i * h) for i in range(1, n, 2)) + sum(2 * f(u0 + i * h) for i in range(2, n, 2))
    return s * h / 3


def enclosed_thermal_energy(m: Model, r, Gamma, ad, eps_e):  # shock-physics.h:452-469
    cool = 3 * (ad - 1)
    if m.ism:
        pe = 3 + cool
        return (1 - eps_e) * (Gamma - 1) * C * C * m.rho_ism * r**3 * (1 - math.exp(-18.0) ** pe) / pe
    return (1 - eps_e) * (Gamma - 1) * C * C * simpson_logspace(
        lambda u: m.rho(math.exp(u)) * math.exp(u) ** 3 * (math.exp(u) / r) ** cool, r)


def estimate_t_dec(m: Model, th):  # src/core/grid-refinement.h:401-453
    g = m.Gamma0_of(th)
    beta = math.sqrt((g - 1) * (g + 1)) / g
    target = m.eps_k_of(th) / (g * C * C) / g
    r_min = 1e-3
    r_max = r_min * 10.0**40
    if target <= 0:
        return r_min * (1 - beta) / (beta * C)
    if m.ism:
        if m.rho_ism > 0:
            r_dec = np.cbrt(max(r_min**3 + 3 * target / m.rho_ism, 0.0))
            return min(float(r_dec), r_max) * (1 - beta) / (beta * C)
        return r_max * (1 - beta) / (beta * C)
    n, u0 = 256, math.log(1e-3)
    du = (u0 + 40 * math.log(10.0) - u0) / n
    mass, r_prev = 0.0, math.exp(u0)
    f_prev = m.rho(r_prev) * r_prev**2
    for i in range(1, n + 1):
        r_i = math.exp(u0 + i * du)
        f_i = m.rho(r_i) * r_i**2
        dr = r_i - r_prev
        mass += 0.5 * (f_prev + f_i) * dr
        if mass >= target:
            return (r_prev + (target - (mass - 0.5 * (f_prev + f_i) * dr)) / f_i) * (1 - beta) / (beta * C)
        f_prev, r_prev = f_i, r_i
    return math.exp(u0 + 40 * math.log(10.0)) * (1 - beta) / (beta * C)


# ---- Dormand-Prince 5(4) with the Boost.odeint 1.82 controller and dense output ------------------------------------
# tableau: external/boost/numeric/odeint/stepper/runge_kutta_dopri5.hpp:92-198, dense output :229-275,
# controller: stepper/controlled_runge_kutta.hpp:64-153
_A = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_B = ((1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9), (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
      (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656), (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_C = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
_DC = (35 / 384 - 5179 / 57600, 0.0, 500 / 1113 - 7571 / 16695, 125 / 192 - 393 / 640, -2187 / 6784 + 92097 / 339200,
       11 / 84 - 187 / 2100, -1 / 40)


class DenseDopri5:
    """make_dense_output(tol, tol, runge_kutta_dopri5): do_step() advances one accepted step, dense(tq) evaluates
    the continuous extension inside it."""

    def __init__(self, f, x0, t0, dt0, tol):
        self.f, self.x, self.t, self.dt, self.tol = f, np.array(x0, dtype=float), t0, dt0, tol
        self.k1 = f(self.x, t0)
        self.t_old = t0

    def do_step(self):
        f, tol = self.f, self.tol
        fails = 0
        while True:  # controlled_runge_kutta::try_step until success (max_step_checker: 500 failures)
            x, t, dt, k1 = self.x, self.t, self.dt, self.k1
            ks = [k1]
            for s in range(6):
                xt = 1.0 * x
                for j, b in enumerate(_B[s]):
                    if b != 0.0:
                        xt = xt + (dt * b) * ks[j]
                ks.append(f(xt, t + dt * _A[s]))
            err_vec = sum((dt * dc) * kk for dc, kk in zip(_DC, ks) if dc != 0.0)
            with np.errstate(invalid="ignore", divide="ignore"):
                err = float(np.max(np.abs(err_vec) / (tol + tol * (np.abs(x) + abs(dt) * np.abs(k1)))))
            if err > 1.0:
                self.dt = dt * max(0.9 * err ** (-1 / 3), 0.2)
                fails += 1
                if fails >= 500:
                    return False
                continue
            break
        self.x_old, self.k_old, self.t_old, self.ks = x, k1, t, ks
        self.x, self.k1, self.t = xt, ks[6], t + dt
        if err < 0.5:
            self.dt = dt * 0.9 * max(5.0**-5, err) ** (-1 / 5)
        return True

    def dense(self, tq):  # Hairer-Norsett-Wanner I, p. 191 as coded in runge_kutta_dopri5.hpp:229-275
        h = self.t - self.t_old
        th = (tq - self.t_old) / h
        X1 = 5.0 * (2558722523.0 - 31403016.0 * th) / 11282082432.0
        X3 = 100.0 * (882725551.0 - 15701508.0 * th) / 32700410799.0
        X4 = 25.0 * (443332067.0 - 31403016.0 * th) / 1880347072.0
        X5 = 32805.0 * (23143187.0 - 3489224.0 * th) / 199316789632.0
        X6 = 55.0 * (29972135.0 - 7076736.0 * th) / 822651844.0
        X7 = 10.0 * (7414447.0 - 829305.0 * th) / 29380423.0
        A, Bc, Cc, D = th * th * (3 - 2 * th), th * th * (th - 1), th * th * (th - 1) ** 2, th * (th - 1) ** 2
        ks = self.ks
        return (self.x_old + h * (A * _C[0] - Cc * X1 + D) * self.k_old + h * (A * _C[2] + Cc * X3) * ks[2]
                + h * (A * _C[3] - Cc * X4) * ks[3] + h * (A * _C[4] + Cc * X5) * ks[4]
                + h * (A * _C[5] - Cc * X6) * ks[5] + h * (Bc + Cc * X7) * ks[6])


def integrate_dense(f, x0, t0, dt0, tol, t_out, max_steps=100000):
    """Stepped until every t_out has been passed; states at t_out (NaN rows where the solve gave up)."""
    st = DenseDopri5(f, x0, t0, dt0, tol)
    out = np.full((len(t_out), len(x0)), np.nan)
    k_out, steps = 0, 0
    while st.t <= t_out[-1]:
        if not st.do_step():
            break
        steps += 1
        if steps > max_steps:
            break
        while k_out < len(t_out) and st.t > t_out[k_out]:
            out[k_out] = st.dense(t_out[k_out])
            k_out += 1
    return out


# ---- forward shock: src/dynamics/forward-shock.tpp:17-208 ----------------------------------------------------------
def radiative_efficiency(m: Model, t_comv, Gamma, e_th):  # shock-physics.h:247-288
    eps = m.eps_e if m.radiative else 0.0
    if eps == 0:
        return 0.0
    gm = m.gm_coeff * (Gamma - 1) + 1
    with np.errstate(divide="ignore"):
        gbar = m.gc_coeff / (e_th * t_comv) if e_th * t_comv != 0 else math.inf
    gc = 0.5 * (gbar + math.sqrt(gbar * gbar + 4)) if math.isfinite(gbar) else math.inf
    ratio = gm / gc
    if ratio < 1 and m.p > 2:
        return eps * (2.0 ** ((m.p - 2) * math.log2(ratio)) if ratio > 0 else 0.0)
    return eps


def solve_forward_row(m: Model, theta, t_lat):
    """Shock table of one row: t_comv, r, Gamma, Gamma_th, B, N_p at the lattice t_lat (code units)."""
    G4 = m.Gamma0_of(theta)
    m_jet0 = m.eps_k_of(theta) / G4 / (C * C)
    n = len(t_lat)
    t0 = min(t_lat[0], 0.1 * SEC, 0.1 * estimate_t_dec(m, theta))
    beta4 = math.sqrt((G4 - 1) * (G4 + 1)) / G4
    r0 = beta4 * C * t0 * G4 * G4 * (1 + beta4)
    tc0 = r0 / math.sqrt((G4 - 1) * (G4 + 1)) / C if G4 > 1 else math.inf
    if G4 <= GAMMA_CUT:  # set_stopping_shock: shock-physics.h:388-397
        tc0 = r0 / math.sqrt((G4 - 1) * (G4 + 1)) / C if G4 > 1 else math.nan
        return dict(t_comv=np.full(n, tc0), r=np.full(n, r0), Gamma=np.ones(n), Gamma_th=np.ones(n), B=np.zeros(n), N_p=np.zeros(n))
    U0 = enclosed_thermal_energy(m, r0, G4, adiabatic_idx(G4), m.eps_e if m.radiative else 0.0)
    x0 = [G4, m.mass(r0), U0, r0, tc0]  # Gamma, m2, U2_th, r, t_comv

    def rhs(x, _t):  # ForwardShockEqn::operator(): forward-shock.tpp:27-118
        Gm, m2, U, r, tc = x
        u2 = (Gm - 1) * (Gm + 1)
        u = math.sqrt(u2) if u2 >= 0 else math.nan
        dr = u * (Gm + u) * C
        rho = m.rho(r)
        dm2 = r * r * rho * dr
        e_th = (Gm - 1) * 4 * Gm * rho * C * C
        eps_rad = radiative_efficiency(m, tc, Gm, e_th)
        ad = adiabatic_idx(Gm)
        G2 = Gm * Gm
        Geff = (ad * (G2 - 1) + 1) / Gm
        dGeff = (ad * (G2 + 1) - 1) / G2
        dlnV = 3 / r * dr
        dG = (-(Gm - 1) * (Geff + 1) * C * C * dm2 + (ad - 1) * Geff * U * dlnV) / (
            (m_jet0 + m2) * C * C + (dGeff + Geff * (ad - 1) / Gm) * U)
        dU = (1 - eps_rad) * (Gm - 1) * C * C * dm2 - (ad - 1) * (dlnV - dG / Gm) * U
        return np.array([dG, dm2, dU, dr, Gm + u])

    X = integrate_dense(rhs, x0, t0, 0.01 * t0, m.rtol, list(t_lat))
    tab = dict(t_comv=np.zeros(n), r=np.zeros(n), Gamma=np.ones(n), Gamma_th=np.ones(n), B=np.zeros(n), N_p=np.zeros(n))
    for k in range(n):  # save_fwd_shock_state: forward-shock.tpp:151-173
        if not np.isfinite(X[k, 0]):
            continue
        Gm, m2, U, r, tc = X[k]
        ad = adiabatic_idx(Gm)  # compute_compression(1, Gamma, 0): shock.cpp:90-138, shock-physics.h:40-66,352-355
        u_down = math.sqrt(max((Gm - 1) * (ad - 1) ** 2 / (-ad * (ad - 2) * (Gm - 1) + 2), 0.0))
        u_up = math.sqrt((1 + u_down**2) * max((Gm - 1) * (Gm + 1), 0.0)) + u_down * Gm
        comp = u_up / u_down if u_down != 0 else 4 * Gm
        Gth = U / (m2 * C * C) + 1 if m2 != 0 else 1.0
        e_th = (Gth - 1) * m.rho(r) * comp * C * C
        tab["t_comv"][k], tab["r"][k], tab["Gamma"][k], tab["Gamma_th"][k] = tc, r, Gm, Gth
        tab["B"][k], tab["N_p"][k] = math.sqrt(8 * PI * m.eps_B * e_th), m2 / MP
    return tab


# ---- forward + reverse shock pair: src/dynamics/reverse-shock.tpp:11-591 (unmagnetised shells) ------------------------
def _compression(Gamma_rel):  # compute_4vel_jump(gamma_rel, sigma = 0): shock.cpp:90-138, shock-physics.h:40-66
    ad = adiabatic_idx(Gamma_rel)
    u_down = math.sqrt(max((Gamma_rel - 1) * (ad - 1) ** 2 / (-ad * (ad - 2) * (Gamma_rel - 1) + 2), 0.0))
    u_up = math.sqrt((1 + u_down**2) * max((Gamma_rel - 1) * (Gamma_rel + 1), 0.0)) + u_down * Gamma_rel
    return u_up / u_down if u_down != 0 else 4 * Gamma_rel


def _rel_Gamma(g1, g2):  # shock-physics.h:192-204
    u1u2 = math.sqrt(max((g1 - 1) * (g1 + 1) * (g2 - 1) * (g2 + 1), 0.0))
    den = g1 * g2 - 1 + u1u2
    return 1.0 if den <= 0 else 1 + (g1 - g2) ** 2 / den


def _sound_speed(G):  # shock-physics.h:75-78
    ad = adiabatic_idx(G)
    return math.sqrt(max(ad * (ad - 1) * (G - 1) / (1 + (G - 1) * ad), 0.0)) * C


def _smoothstep(e0, e1, x):  # reverse-shock.tpp:11-20
    t = min(max((x - e0) / (e1 - e0), 0.0), 1.0)
    return t * t * (3.0 - 2.0 * t)


def solve_pair_row(m: Model, rvs_eps_B, theta, T0, t_lat):
    """(forward table, reverse table, injection_idx) of one row: grid_solve_shock_pair (reverse-shock.tpp:511-591).
    State: Gamma, x4, x3, m2, m3, U2_th, U3_th, r, t_comv, eps4, m4 (reverse-shock.hpp:29-45 without theta)."""
    iG, iX4, iX3, iM2, iM3, iU2, iU3, iR, iT, iE4, iM4 = range(11)
    n = len(t_lat)
    G4 = m.Gamma0_of(theta)
    deps0 = m.eps_k_of(theta) / T0
    dm0 = deps0 / (G4 * C * C)
    u4 = math.sqrt(G4 * G4 - 1) * C
    cs4 = _sound_speed(G4)
    beta4 = math.sqrt((G4 - 1) * (G4 + 1)) / G4
    eps_e_rad = m.eps_e if m.radiative else 0.0

    def init_state(t0):  # set_init_state: reverse-shock.tpp:314-357
        x = np.zeros(11)
        x[iR] = beta4 * C * t0 * G4 * G4 * (1 + beta4)
        x[iT] = x[iR] / math.sqrt((G4 - 1) * (G4 + 1)) / C
        dt = min(t0, T0)
        x[iE4], x[iM4] = deps0 * dt, dm0 * dt
        x[iX4] = G4 * t0 * beta4 * C if t0 < T0 else G4 * T0 * beta4 * C + cs4 * (t0 - T0) * G4
        x[iM2] = simpson_logspace(lambda u: m.rho(math.exp(u)) * math.exp(u) ** 3, x[iR])
        mj = dm0 * T0
        x[iG] = G4 / (1 + x[iM2] / mj) if (mj > 0 and x[iM2] > 0) else G4
        ad = adiabatic_idx(x[iG])
        cool = 3 * (ad - 1)
        x[iU2] = (1 - eps_e_rad) * (x[iG] - 1) * C * C * simpson_logspace(
            lambda u: m.rho(math.exp(u)) * math.exp(u) ** 3 * (math.exp(u) / x[iR]) ** cool, x[iR])
        G34 = _rel_Gamma(G4, x[iG])
        if G34 > 1 and x[iM4] > 0 and x[iX4] > 0:
            x[iX3] = x[iX4] * 1e-8
            x[iM3] = x[iM4] * _compression(G34) * x[iX3] / x[iX4]
            x[iU3] = (G34 - 1) * x[iM3] * C * C
        return x

    def rhs(xr, t):  # FRShockEqn::operator(): reverse-shock.tpp:252-294 and the rate terms :62-250
        Gm = min(max(xr[iG], 1.0), G4)
        m4 = xr[iM4]
        m3 = min(max(xr[iM3], 0.0), max(m4, 0.0))
        x3, U3 = max(xr[iX3], 0.0), max(xr[iU3], 0.0)
        x4, m2, U2, r, tc = xr[iX4], xr[iM2], xr[iU2], xr[iR], xr[iT]
        d = np.zeros(11)
        u3 = math.sqrt((Gm - 1) * (Gm + 1))
        dr, dtc = u3 * (Gm + u3) * C, Gm + u3
        d[iR], d[iT] = dr, dtc
        rho = m.rho(r)
        dm2 = r * r * rho * dr
        d[iM2] = dm2
        w = _smoothstep(T0 * 1.5, T0 * 0.5, t)
        deps4, dm4 = (w * deps0, w * dm0) if w > 1e-6 else (0.0, 0.0)
        d[iE4], d[iM4] = deps4, dm4
        G34 = _rel_Gamma(G4, Gm)
        comp = _compression(G34)
        f = min(dm4 / dm0, 1.0) if (dm0 > 0 and dm4 > 0) else 0.0
        se4 = cs4 * dtc
        dx4 = f * u4 + (1 - f) * se4 if f > 1e-6 else se4
        d[iX4] = dx4
        se3 = _sound_speed(G34) * dtc
        dx3 = se3
        remaining = max(m4 - m3, 0.0)
        if not (m4 <= 0):
            cw = f + (1.0 - f) * remaining / m4
            if not (cw < 1e-6):
                pen = Gm * comp / G4 - 1
                if not (pen <= 0):
                    beta3 = math.sqrt((Gm - 1) * (Gm + 1)) / Gm
                    dx3dt = (G4 - Gm) * (G4 + Gm) * (1 + beta3) * C / (G4 * G4 * (beta3 + beta4) * pen)
                    crossing = abs(dx3dt * Gm)
                    if pen < 1:
                        cs = _sound_speed(G34)
                        crossing = min(crossing, math.sqrt(cs * cs / (C * C)) * C * dtc)
                    dx3 = cw * crossing + (1.0 - cw) * se3
        d[iX3] = dx3
        dm3 = 0.0
        if not (m4 <= 0) and not (remaining <= 0 and f < 1e-6):
            dm3dt = (f * m4 + (1.0 - f) * remaining) * comp / x4 * dx3
            if f > 1e-6:
                cap = _smoothstep(0, 1.0, m3 / m4)
                dm3 = (1.0 - cap) * dm3dt + cap * min(dm3dt, dm4)
            else:
                dm3 = dm3dt
        d[iM3] = dm3
        ad2, ad3 = adiabatic_idx(Gm), adiabatic_idx(G34)
        eps_rad = radiative_efficiency(m, tc, Gm, (Gm - 1) * 4 * Gm * rho * C * C)
        dlnv2 = 2 * dr / r + (dx4 / x4 if x4 > 0 else 0.0)
        dU2 = (1 - eps_rad) * dm2 * (Gm - 1) * C * C - (ad2 - 1) * dlnv2 * U2
        dlnv3 = 2 * dr / r + (dx3 / x3 if x3 > 0 else 0.0)
        dU3 = dm3 * (G34 - 1) * C * C - (ad3 - 1) * dlnv3 * U3
        d[iU2], d[iU3] = dU2, dU3
        Ge = lambda ad: (ad * Gm * Gm - ad + 1) / Gm          # noqa: E731  compute_effective_Gamma
        dGe = lambda ad: (ad * Gm * Gm + ad - 1) / (Gm * Gm)  # noqa: E731
        a = (Gm - 1) * C * C * dm2 + (Gm - G4) * C * C * dm3 + Ge(ad2) * dU2 + Ge(ad3) * dU3
        b = (m2 + m3) * C * C + dGe(ad2) * U2 + dGe(ad3) * U3
        q = -a / b if b != 0 else math.nan
        d[iG] = q if (b != 0 and math.isfinite(q)) else 0.0
        return d

    def complete(x, t):  # crossing_complete: reverse-shock.tpp:49-60
        return not (x[iM3] < 0.999 * x[iM4]) and not (_smoothstep(T0 * 1.5, T0 * 0.5, t) > 1e-6)

    t0 = min(t_lat[0], 0.01 * SEC, 0.1 * estimate_t_dec(m, theta))
    x0 = init_state(t0)
    blank = lambda: dict(t_comv=np.zeros(n), r=np.zeros(n), Gamma=np.ones(n), Gamma_th=np.ones(n), B=np.zeros(n), N_p=np.zeros(n))  # noqa: E731
    F, R = blank(), blank()
    if x0[iG] <= 1.03:  # RS_Gamma_limit: stopping shock
        for tb in (F, R):
            tb["t_comv"][:], tb["r"][:] = x0[iT], x0[iR]
        return F, R, n
    st = DenseDopri5(rhs, x0, t0, 1e-9 * t0, m.rtol)
    crossing, pending, inj, t_cross, t_start = True, False, n, 0.0, t0
    cross = {}
    k, steps = 0, 0

    def save(k, x):  # save_fwd_shock_state + save_rvs_shock_state: forward-shock.tpp:151-173, reverse-shock.tpp:403-426
        comp2 = _compression(_rel_Gamma(1.0, x[iG]))
        Gth2 = x[iU2] / (x[iM2] * C * C) + 1 if x[iM2] != 0 else 1.0
        F["t_comv"][k], F["r"][k], F["Gamma"][k], F["Gamma_th"][k] = x[iT], x[iR], x[iG], Gth2
        F["B"][k] = math.sqrt(8 * PI * m.eps_B * (Gth2 - 1) * m.rho(x[iR]) * comp2 * C * C)
        F["N_p"][k] = x[iM2] / MP
        if k <= inj:
            comp34 = _compression(_rel_Gamma(G4, x[iG]))
            rho4 = x[iM4] / (x[iR] * x[iR] * x[iX4])
            Gth3 = x[iU3] / (x[iM3] * C * C) + 1 if x[iM3] != 0 else 1.0
            if Gth3 < 1 + 1e-6:
                Gth3 = 1.0
            B3 = math.sqrt(8 * PI * rvs_eps_B * (Gth3 - 1) * rho4 * comp34 * C * C)
        else:
            comp = cross["V3"] / (x[iR] * x[iR] * x[iX3])
            Gth3 = x[iU3] / (x[iM3] * C * C) + 1 if x[iM3] != 0 else 1.0
            B3 = math.sqrt(8 * PI * rvs_eps_B * (Gth3 - 1) * cross["rho3"] * comp * C * C)
        R["t_comv"][k], R["r"][k], R["Gamma"][k], R["Gamma_th"][k], R["B"][k], R["N_p"][k] = x[iT], x[iR], x[iG], Gth3, B3, x[iM3] / MP

    while st.t <= t_lat[-1]:
        if not st.do_step():
            break
        steps += 1
        if steps > 100000 or st.t + st.dt == st.t:
            break
        if crossing and complete(st.x, st.t):  # locate_crossing_time: reverse-shock.tpp:482-495
            lo, hi = t_start, st.t
            for _ in range(100):
                if not ((hi - lo) > 1e-12 * hi):
                    break
                mid = 0.5 * (lo + hi)
                if complete(st.dense(mid), mid):
                    hi = mid
                else:
                    lo = mid
            xc = st.dense(hi)
            t_cross = hi
            rho4 = xc[iM4] / (xc[iR] * xc[iR] * xc[iX4])  # save_cross_state: reverse-shock.tpp:298-312
            cross = dict(V3=xc[iR] * xc[iR] * xc[iX3], rho3=rho4 * _compression(_rel_Gamma(G4, xc[iG])))
            crossing, pending = False, True
        t_start = st.t
        while k < n and st.t > t_lat[k]:
            x = st.dense(t_lat[k])
            if pending and t_lat[k] >= t_cross:
                inj, pending = (k if k > 0 else 1), False
            save(k, x)
            k += 1
    # reverse_shock_early_extrap: reverse-shock.tpp:428-467
    cut = next((q for q in range(n) if R["Gamma_th"][q] > 1 + 1e-6), n)
    if not (cut == 0 or cut >= n - 2 or cut >= inj):
        l2r = math.log2(R["r"][cut])
        dl = math.log2(R["r"][cut + 2]) - l2r
        for key, off in (("Gamma_th", 1.0), ("B", 0.0), ("N_p", 0.0)):
            l0 = math.log2(R[key][cut] - off)
            slope = (math.log2(R[key][cut + 2] - off) - l0) / dl
            for q in range(cut):
                R[key][q] = off + 2.0 ** (l0 + slope * (math.log2(R["r"][q]) - l2r))
    return F, R, inj


# ---- synchrotron electrons and photons: src/radiation/synchrotron.cpp:45-408, smooth-power-law-syn.cpp:26-166 ------
_KS = 3 * E / (4 * PI * ME * C)


def _softplus2(x):  # src/util/fast-math.h:179-185
    x = np.asarray(x, dtype=float)
    with np.errstate(over="ignore", invalid="ignore"):
        mid = np.log2(1.0 + np.exp2(np.clip(x, -20, 20)))
    return np.where(x > 20, x, np.where(x < -20, 0.0, mid))


def photon_tables(m: Model, tab, rad=None, inj=None):
    """Per-cell spectral coefficients (generate_syn_electrons + generate_syn_photons + SmoothPowerLawSyn::build).
    `rad` = (eps_e, eps_B, p, xi_e) of the shock (default: the forward shock's); cells k >= inj are relic cells whose
    gamma_c / gamma_M cool adiabatically from the crossing cell inj-1 (synchrotron.cpp:190-195, synchrotron.h:187-201)."""
    eps_e, _eps_B, p, xi_e = rad if rad is not None else (m.eps_e, m.eps_B, m.p, m.xi_e)
    B, r, Gth, Np, tc = tab["B"], tab["r"], tab["Gamma_th"], tab["N_p"], tab["t_comv"]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        gM = np.where(B == 0, np.inf, np.sqrt(6 * PI * E / SIGMA_T / B))
        gave = eps_e * (Gth - 1) * (MP / ME) / xi_e
        if p > 2:
            gm = (p - 2) / (p - 1) * gave + 1
        else:  # 1 < p < 2 (p == 2 root-finding is not restated)
            gm = ((2 - p) / (p - 1) * gave * gM ** (p - 2)) ** (1 / (p - 1)) + 1
        fsyn = (gm - 1) / gm
        if p > 3:
            fsyn = np.exp2((p - 1) / 2 * np.log2(fsyn))
        Ne = Np * xi_e * fsyn
        col = Ne / (r * r)
        gbar = (6 * PI * ME * C / SIGMA_T) / (B * B * tc)
        gc = (gbar + np.sqrt(gbar * gbar + 4)) / 2
        if inj is not None and 0 < inj < len(B):  # relic cells: cool_after_crossing from cell inj - 1
            f_ad = (gm[inj:] - 1) / (gm[inj - 1] - 1)
            gc = np.concatenate([gc[:inj], (gc[inj - 1] - 1) * f_ad + 1])
            gM = np.concatenate([gM[:inj], (gM[inj - 1] - 1) * f_ad + 1])
        I_peak = B * ((PI / 4) * 0.92 * math.sqrt(3.0) * E**3 / (ME * C * C)) * col / (4 * PI)
        freq = lambda g: np.where((B == 0) | ~np.isfinite(g), 0.0, _KS * B * g * g)  # noqa: E731
        # compute_syn_gamma_a: synchrotron.cpp:212-246 (no inverse Compton: every ic factor is 1)
        gpk = np.minimum(gm, gc)
        nu_pk, kT = freq(gpk), (gpk - 1) * (ME * C * C) / 3
        pw = lambda a, b: np.exp2(b * np.log2(a))  # noqa: E731  fast_pow
        base = I_peak * C * C / (2 * kT)
        nu_m, nu_c = freq(gm), freq(gc)
        nu_a = pw(I_peak * C * C / (np.cbrt(nu_pk) * 2 * kT), 0.6)
        slow = gc > gm
        a_mid_slow = pw(base * pw(nu_m, p / 2), 2 / (p + 4))
        a_hi = pw(base * np.sqrt(nu_c) * pw(nu_m, p / 2), 2 / (p + 5))
        a_mid_fast = pw(base * np.sqrt(nu_c), 0.4)
        over = nu_a > nu_pk
        nu_a_slow = np.where(a_mid_slow > nu_c, a_hi, a_mid_slow)
        nu_a_fast = np.where(a_mid_fast > nu_m, a_hi, a_mid_fast)
        nu_a = np.where(over, np.where(slow, nu_a_slow, nu_a_fast), nu_a)
        ga = np.sqrt((4 * PI * ME * C / (3 * E)) * (nu_a / B)) + 1
        nu_a = freq(ga)
        nu_M = freq(gM)
        l2m, l2c, l2a = np.log2(nu_m), np.log2(nu_c), np.log2(nu_a)
        c = dict(l2I=np.log2(I_peak), l2m=l2m, l2M=np.log2(nu_M), inv_M=1.0 / nu_M)
        sig = lambda x: 1.0 / (1.0 + np.exp2(-x))  # noqa: E731
        w_slow = sig(4.0 * (l2c - l2m))
        soft = _softplus2(-4.0 * np.abs(l2c - l2m)) / 4.0
        c["lo"], c["hi"] = np.minimum(l2m, l2c) - soft, np.maximum(l2m, l2c) + soft
        blend = lambda w, a, b: w * a + (1 - w) * b  # noqa: E731
        s_lo = blend(w_slow, max(1.84 - 0.40 * p, 0.1), 0.597)
        s_hi = blend(w_slow, max(1.15 - 0.06 * p, 0.1), max(3.34 - 0.82 * p, 0.1))
        a_mid = blend(w_slow, -0.5 * (p - 1.0), -0.5)
        c["s_lo"], c["s_hi"] = s_lo, s_hi
        c["d_lo"], c["d_hi"] = s_lo * (1 / 3 - a_mid), s_hi * (a_mid + 0.5 * p)
        u, v = sig(4.0 * (l2a - l2m)), sig(4.0 * (l2a - l2c))
        wb, wa = (1 - u) * (1 - v), u * v
        c["s_a"] = wb * 1.64 + wa * max(0.94 - 0.14 * p, 0.1) + (1 - wb - wa) * max(1.47 - 0.21 * p, 0.1)
        c["norm"] = 1.0 / s_lo
        # sharp forms for the thick normalisation: smooth-power-law-syn.cpp:48-74
        thick_sharp = np.where(l2a < l2m, 2.0 * (l2a - l2m), 2.5 * (l2a - l2m))
        thin_slow = np.where(l2a < l2m, (l2a - l2m) / 3, np.where(l2a < l2c, 0.5 * (1 - p) * (l2a - l2m),
                             0.5 * (1 - p) * (l2c - l2m) - 0.5 * p * (l2a - l2c)))
        thin_fast = np.where(l2a < l2c, (l2a - l2c) / 3, np.where(l2a < l2m, -0.5 * (l2a - l2c),
                             -0.5 * (l2m - l2c) - 0.5 * p * (l2a - l2m)))
        c["thick_norm"] = np.where(l2m < l2c, thin_slow, thin_fast) - thick_sharp
    c["p"] = p
    return c


def log2_I_nu(m: Model, c, l2nu):
    """SmoothPowerLawSyn::compute_log2_I_nu (smooth-power-law-syn.cpp:26-46,80-92,159-166); arrays broadcast."""
    smooth_thick = (3.44 * c["p"] - 1.41) / LN2
    x_far = 1.5 * math.log2(20.0 / smooth_thick)
    with np.errstate(over="ignore", invalid="ignore"):
        thin = ((l2nu - c["lo"]) / 3.0 - _softplus2(c["d_lo"] * (l2nu - c["lo"])) / c["s_lo"]
                - _softplus2(c["d_hi"] * (l2nu - c["hi"])) / c["s_hi"])
        lx = l2nu - c["l2m"]
        s = -smooth_thick * np.exp2(2.0 / 3 * lx)
        thick = np.where(lx > x_far, 2.5 * lx, 2.5 * lx + _softplus2(-0.5 * lx + s))
        b = thick + c["thick_norm"]
        smooth = thin - _softplus2(c["s_a"] * (thin - b)) / c["s_a"]
        spec = c["l2I"] + (c["norm"] + smooth)
        return np.where(l2nu - c["l2M"] < -20, spec, spec - LOG2E * c["inv_M"] * np.exp2(l2nu))


# ---- observer: src/core/observer.cpp:17-37,143-205,439-454, observer.h:355-445 -------------------------------------
def flux_density_grid(p, theta, phi, t_rows, reps, phi_mirrored, n_phi_eff, t_obs, nu_obs):
    """F_nu[n_nu, n_t] (erg cm^-2 s^-1 Hz^-1) of one model on the GIVEN grid: theta[N_theta], phi[N_phi],
    t_rows[n_reps, N_t] (engine-frame lattice of each representative row, code units), reps.
    Forward-shock model: returns the forward synchrotron flux.  With a reverse shock (has_rvs): returns
    (forward, reverse); one EAT geometry serves both (pybind/pymodel.h:943-950)."""
    m = Model(p)
    theta, phi = np.asarray(theta, float), np.asarray(phi, float)
    n_th = theta.size
    has_rvs = bool(np.asarray(p["has_rvs"]).reshape(-1)[0])
    shocks = []  # per shock: (tables per rep, coefficient tables per rep)
    if has_rvs:
        rv = np.asarray(p["rvs"]).reshape(-1)[0]
        rad_r = tuple(float(rv[k]) for k in ("eps_e", "eps_B", "p", "xi_e"))
        T0 = float(np.asarray(p["duration"]).reshape(-1)[0]) * SEC
        rows = [solve_pair_row(m, rad_r[1], theta[j0], T0, t_rows[r]) for r, j0 in enumerate(reps)]
        tabs = [r[0] for r in rows]
        shocks.append((tabs, [photon_tables(m, tb) for tb in tabs]))
        shocks.append(([r[1] for r in rows], [photon_tables(m, r[1], rad_r, r[2]) for r in rows]))
    else:
        tabs = [solve_forward_row(m, theta[j0], t_rows[r]) for r, j0 in enumerate(reps)]
        shocks.append((tabs, [photon_tables(m, tb) for tb in tabs]))
    rep_of = np.searchsorted(np.asarray(reps), np.arange(n_th), side="right") - 1
    l2t_obs = np.log2(np.asarray(t_obs, float) * SEC)
    l2nu = np.log2(np.asarray(nu_obs, float) * HZ) + math.log2(1 + m.z)
    F = [np.zeros((l2nu.size, l2t_obs.size)) for _ in shocks]
    cos_o, sin_o = math.cos(m.theta_v), math.sin(m.theta_v)
    last = n_phi_eff - 1
    for i in range(n_phi_eff):
        if n_phi_eff == 1:
            dphi = 2 * PI
        elif phi_mirrored:
            left = 0.5 * (phi[i - 1] + phi[i]) if i > 0 else 0.0
            right = 0.5 * (phi[i] + phi[i + 1]) if i < last else PI
            dphi = 2 * (right - left)
        else:
            dphi = 0.5 * (phi[min(i + 1, last)] - phi[max(i - 1, 0)])
        for j in range(n_th):
            tb = tabs[rep_of[j]]
            t_eng = np.asarray(t_rows[rep_of[j]], float)
            cos_v = math.sin(theta[j]) * math.cos(phi[i]) * sin_o + math.cos(theta[j]) * cos_o
            c_lo = math.cos(theta[j]) if j == 0 else math.cos(0.5 * (theta[j - 1] + theta[j]))
            c_hi = math.cos(theta[j]) if j == n_th - 1 else math.cos(0.5 * (theta[j] + theta[j + 1]))
            dOmega = abs((c_hi - c_lo) * dphi)
            Gm, r = tb["Gamma"], tb["r"]
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                l2dop = -np.log2(Gm - np.sqrt((Gm - 1) * (Gm + 1)) * cos_v)
                l2t = np.log2(t_eng * (1 + m.z) + (1 - cos_v) / C * (1 + m.z) * r)
                l2geom = (np.log2(np.float64(dOmega)) + 2 * np.log2(r)) + 3 * l2dop
                # iterate_to: t_row[k] <= x < t_row[k+1] (observer.h:309-313,405-433)
                k = np.searchsorted(l2t, l2t_obs, side="right") - 1
                ok = (k >= 0) & (k <= l2t.size - 2)
                kk = np.clip(k, 0, l2t.size - 2)
                for si, (_tabs, coefs) in enumerate(shocks):
                    L = log2_I_nu(m, coefs[rep_of[j]], l2nu[:, None] - l2dop[None, :]) + l2geom[None, :]  # [n_nu, N_t]
                    slope = (L[:, kk + 1] - L[:, kk]) / (l2t[kk + 1] - l2t[kk])[None, :]
                    val = np.exp2(L[:, kk] + (l2t_obs - l2t[kk])[None, :] * slope)
                    F[si] += np.where(ok[None, :] & np.isfinite(slope), val, 0.0)
    scale = ((1 + m.z) / (m.d_L * m.d_L)) / FLUX_DEN_CGS
    return tuple(f * scale for f in F) if has_rvs else F[0] * scale


# =====================================================================================================================
# Row a1: the adaptive (phi, theta, t) grid -- auto_grid and its callees (src/core/grid-refinement.h:40-706,
# src/core/grid-refinement.cpp:140-197, Coord::detect_symmetry src/core/mesh.h:120-185) for the typed jets,
# isotropic media, axisymmetric, non-spreading models.
# =====================================================================================================================
THETA_MIN = 1e-6          # defaults::grid::theta_min
N_SAMPLES = 200           # defaults::sampling::theta_samples


def _linspace_at(a, b, n, i):  # xt::linspace element (xbuilder.hpp:199-222,460-468): a + step i, last forced to b
    step = (b - a) / max(1.0, float(n - 1))
    return b if (n > 1 and i == n - 1) else a + step * float(i)


def _structure_weight(G):  # grid-refinement.h:11-13
    return G * math.sqrt(max((G - 1) * G, 0.0))


def _beta(G):
    return math.sqrt((G - 1) * (G + 1)) / G


def find_jet_jumps(m: Model):  # grid-refinement.h:40-86
    n_scan, lo, hi = 512, THETA_MIN, PI / 2
    dth = (hi - lo) / (n_scan - 1)
    if m.Gamma0_of(hi) >= GAMMA_CUT:
        return [hi]
    jumps, prev_th, prev_G = [], lo, m.Gamma0_of(lo)
    for j in range(1, n_scan):
        cur_th = lo + dth * float(j)
        cur_G = m.Gamma0_of(cur_th)
        if prev_G >= GAMMA_CUT or cur_G >= GAMMA_CUT:
            scale = max(prev_G - 1, cur_G - 1)
            if scale > 0 and abs(cur_G - prev_G) > 0.5 * scale:
                a, b = prev_th, cur_th
                while b - a > 1e-9:
                    mid = 0.5 * (a + b)
                    Gm = m.Gamma0_of(mid)
                    if abs(Gm - prev_G) < abs(Gm - cur_G):
                        a = mid
                    else:
                        b = mid
                jumps.append(a if prev_G > cur_G else b)
        prev_th, prev_G = cur_th, cur_G
    return jumps


def find_theta_range(m: Model):  # grid-refinement.h:88-111
    lo, hi = THETA_MIN, PI / 2
    step = (hi - lo) / 512
    th_max, th_min = hi, lo
    th = hi
    while th >= lo:
        if m.Gamma0_of(th) >= GAMMA_CUT:
            th_max = th
            break
        th -= step
    th = lo
    while th <= hi:
        if m.Gamma0_of(th) >= GAMMA_CUT:
            th_min = th
            break
        th += step
    return th_min, th_max


def inverse_cdf_sampling(pdf, lo, hi, num, log_sample, midpoint):  # grid-refinement.h:137-189
    if log_sample:
        a, b = math.log10(lo), math.log10(hi)
        x_i = [10.0 ** _linspace_at(a, b, N_SAMPLES, i) for i in range(N_SAMPLES)]
    else:
        x_i = [_linspace_at(lo, hi, N_SAMPLES, i) for i in range(N_SAMPLES)]
    cdf = [0.0] * N_SAMPLES
    st = DenseDopri5(lambda _x, t: np.array([pdf(t)]), [0.0], lo, (hi - lo) / 1e3, 1e-6)
    k, steps = 1, 0
    while st.t <= hi:
        if not st.do_step():
            break
        steps += 1
        if steps > 100000:
            break
        while k < N_SAMPLES and st.t > x_i[k]:
            cdf[k] = float(st.dense(x_i[k])[0])
            k += 1
    c0, c1 = cdf[0], cdf[-1]
    out = []
    for q in range(num):
        target = c0 + (c1 - c0) * (q + 0.5) / num if midpoint else _linspace_at(c0, c1, num, q)
        j = 0
        while j < N_SAMPLES and not (target <= cdf[j]):
            j += 1
        if j >= N_SAMPLES:
            out.append(0.0)
        elif j == 0:
            out.append(x_i[0])
        else:
            den = cdf[j] - cdf[j - 1]
            out.append(x_i[j - 1] + (x_i[j] - x_i[j - 1]) / den * (target - cdf[j - 1]) if den > 0 else x_i[j - 1])
    return out


def adaptive_theta_grid(m: Model, th_min, th_max, base_pts, theta_v, theta_resol):  # grid-refinement.h:199-291
    scan, ext = 100, th_max - th_min
    peak_w, G_peak, s_sum, last_bright, G_v = 0.0, 1.0, 0.0, 0, 1.0
    for i in range(scan + 1):
        th = th_min + ext * i / scan
        Gm = m.Gamma0_of(th)
        w = _structure_weight(Gm)
        s_sum += w
        if w > peak_w:
            peak_w, G_peak, last_bright = w, Gm, i
        elif w > 0.01 * peak_w:
            last_bright = i
        dth = th - theta_v
        G_v = max(G_v, Gm / math.sqrt(1.0 + Gm * Gm * dth * dth))
    floor_w = 0.25 * peak_w
    cdf_est = (s_sum / scan + floor_w) * ext
    th_bright = th_min + ext * last_bright / scan
    G_peak = max(G_peak, G_v)
    dop_alpha = 12.0 * math.sqrt(peak_w / max(_structure_weight(G_v), 1.0))
    Gp2, Gv2 = G_peak * G_peak, G_v * G_v
    beam = lambda dec, coeff, off: int(max(0.0, dec - off) * theta_resol * coeff)  # noqa: E731
    core_pts = beam(math.log10(max(1.0, G_peak * (th_bright - th_min))), 55.0, 1.0)
    view_pts = beam(math.log10(max(1.0, G_v * max(theta_v - th_min, th_max - theta_v))), 25.0, 0.0) if theta_v * G_peak > 3.0 else 0
    total = base_pts + core_pts + view_pts
    cal = lambda n, c: (float(n) / base_pts * cdf_est / c) if (n > 0 and c > 0) else 0.0  # noqa: E731
    core_w = cal(core_pts, 0.5 * math.log((1.0 + Gp2 * th_max * th_max) / (1.0 + Gp2 * th_min * th_min)))
    vl, vr = theta_v - th_min, th_max - theta_v
    view_w = cal(view_pts, 0.5 * (math.log(1.0 + Gv2 * vl * vl) + math.log(1.0 + Gv2 * vr * vr)))

    def pdf(th):
        Gm = m.Gamma0_of(th)
        b = _beta(Gm)
        dop = (1 - b) / (1 - b * math.cos(th - theta_v))
        d = th - theta_v
        return (core_w * Gp2 * th / (1.0 + Gp2 * th * th) + view_w * Gv2 * abs(d) / (1.0 + Gv2 * d * d)
                + (1 + dop_alpha * dop) * _structure_weight(Gm) + floor_w)

    return inverse_cdf_sampling(pdf, th_min, th_max, total, True, False)


def adaptive_phi_grid(m: Model, phi_num, theta_v, theta, phi_max, boost_cap):  # grid-refinement.h:295-360
    if theta_v == 0:
        return [_linspace_at(0.0, 2 * PI, phi_num, i) for i in range(phi_num)]
    n = len(theta)
    half = phi_max < 2 * PI
    dcos, beta, sw, ct, st = [], [], [], [], []
    for it in range(n):
        left = 0.0 if it == 0 else 0.5 * (theta[it - 1] + theta[it])
        right = theta[it] if it == n - 1 else 0.5 * (theta[it] + theta[it + 1])
        dcos.append(abs(math.cos(left) - math.cos(right)))
        Gm = m.Gamma0_of(theta[it])
        beta.append(_beta(Gm))
        sw.append(_structure_weight(Gm))
        ct.append(math.cos(theta[it]))
        st.append(math.sin(theta[it]))
    cos_tv, sin_tv = math.cos(theta_v), math.sin(theta_v)

    def weight(phi):
        cp, w = math.cos(phi), 0.0
        for it in range(n):
            ca = ct[it] * cos_tv + st[it] * sin_tv * cp
            w += (1 - beta[it]) / (1 - beta[it] * ca) * sw[it] * dcos[it]
        return w

    A = [weight(phi_max * float(s) / 100) for s in range(101)]
    peak, tot = 0.0, 0.0
    for a in A:
        peak = max(peak, a)
        tot += a
    floor_w = 0.05 * peak
    if boost_cap > 0 and peak > 0:
        conc = (peak + floor_w) / (tot / 101 + floor_w)
        phi_num = int(float(phi_num) * min(max(conc / 5, 1.0), boost_cap))
    return inverse_cdf_sampling(lambda ph: weight(ph) + floor_w, 0, phi_max, phi_num, False, half)


def _band_lattice(ts, t_end, b_lo, b_hi, n, factor):  # logspace_with_band_refinement: grid-refinement.h:533-569
    b_lo, b_hi = max(b_lo, ts), min(b_hi, t_end)
    if not (b_hi > b_lo) or n < 8:
        la, lb = math.log10(ts), math.log10(t_end)
        return [10.0 ** _linspace_at(la, lb, n, i) for i in range(n)]
    l0, l1, l2, l3 = math.log10(ts), math.log10(b_lo), math.log10(b_hi), math.log10(t_end)
    w1, w2, w3 = l1 - l0, factor * (l2 - l1), l3 - l2
    segs = n - 1
    rnd = lambda v: int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))  # noqa: E731  std::round
    n1 = min(rnd(float(segs) * w1 / (w1 + w2 + w3)), segs - 2)
    n3 = min(rnd(float(segs) * w3 / (w1 + w2 + w3)), segs - 1 - n1 - 1)
    n2 = segs - n1 - n3
    g = [10.0 ** (l0 + (l1 - l0) * float(k) / float(n1)) for k in range(n1)]
    g += [10.0 ** (l1 + (l2 - l1) * float(k) / float(n2)) for k in range(n2)]
    g += [10.0 ** ((l2 + (l3 - l2) * float(k) / float(n3)) if n3 > 0 else l3) for k in range(n3 + 1)]
    return g


def _cross_lattice(t_start, t_end, t_refine, t_num, base_num):  # grid-refinement.cpp:166-197
    t_refine = min(max(t_refine, t_start), t_end)
    if t_refine <= t_start or t_refine >= t_end:
        la, lb = math.log10(t_start), math.log10(t_end)
        return [10.0 ** _linspace_at(la, lb, t_num, i) for i in range(t_num)]
    n_post = int(float(base_num) * math.log10(t_end / t_refine) / math.log10(t_end / t_start))
    n_post = max(n_post, 2)
    if n_post >= t_num:
        n_post = t_num // 2
    n_pre = t_num + 1 - n_post
    la, lb = math.log10(t_start), math.log10(t_refine)
    g = [10.0 ** _linspace_at(la, lb, n_pre, k) for k in range(n_pre)]
    la, lb = math.log10(t_refine), math.log10(t_end)
    g += [10.0 ** _linspace_at(la, lb, n_post, k) for k in range(1, n_post)]
    return g[:t_num]


def auto_grid(p, t_obs_min, t_obs_max):
    """auto_grid (grid-refinement.h:638-706): theta[N_theta], phi[N_phi], reps, t_rows[n_reps][N_t] (code units),
    phi_mirrored, n_phi_eff for one axisymmetric, non-spreading typed-jet model; t_obs in seconds."""
    m = Model(p)
    g = lambda k: float(np.asarray(p[k]).reshape(-1)[0])  # noqa: E731
    is_rvs = bool(g("has_rvs"))
    phi_res = g("phi_resol") if g("phi_resol") > 0 else 0.06
    th_res = g("theta_resol") if g("theta_resol") > 0 else (0.2 if is_rvs else 0.15)
    t_res = g("t_resol") if g("t_resol") > 0 else (10.0 if is_rvs else 6.0)
    T0 = g("duration") * SEC
    tv = m.theta_v
    t_min, t_max = t_obs_min * SEC, t_obs_max * SEC
    jumps = find_jet_jumps(m)
    inner, outer = find_theta_range(m)
    for j in jumps:
        outer = max(outer, j)
    th_min, th_max = max(THETA_MIN, inner), min(outer, PI / 2)
    theta_num = 36 + int((th_max - th_min) * 180 / PI * th_res)
    base = adaptive_theta_grid(m, th_min, th_max, theta_num, tv, th_res)
    # jump_refinement_grid (grid-refinement.cpp:140-164) + merge_grids (grid-refinement.h:362-393)
    tight = (th_max - th_min) / len(base) / 8
    feat = []
    for jt in jumps:
        if jt >= PI / 2 - 0.01:
            continue
        if jt - tight >= th_min:
            feat.append(jt - tight)
        if jt + tight <= th_max:
            feat.append(jt + tight)
        if th_min <= jt <= th_max:
            feat.append(jt)
    feat = sorted(set(feat))
    theta, i, j = [], 0, 0
    add = lambda v: theta.append(v) if (not theta or theta[-1] != v) else None  # noqa: E731
    while i < len(base) and j < len(feat):
        if base[i] <= feat[j]:
            add(base[i])
            i += 1
            if base[i - 1] == feat[j]:
                j += 1
        else:
            add(feat[j])
            j += 1
    for v in base[i:]:
        add(v)
    for v in feat[j:]:
        add(v)
    n_th = len(theta)
    # phi grid (grid-refinement.h:664-693)
    phi_base = max(int(360 * phi_res), 1)
    mirror = tv != 0 and phi_base > 4
    if mirror:
        phi = adaptive_phi_grid(m, (phi_base + 1) // 2, tv, theta, PI, 5.0)
    else:
        boost = math.sqrt(max(m.Gamma0_of(tv) * math.sin(tv) / (2 * PI), 1.0))
        phi_num = min(max(int(phi_base * boost), 1), phi_base * 5)
        if phi_num <= 2:
            phi = [_linspace_at(0.0, 2 * PI, phi_num, i) for i in range(phi_num)]
        else:
            phi = adaptive_phi_grid(m, phi_num, tv, theta, 2 * PI, 0.0)
        if len(phi) >= 2:
            shift = 0.5 * (phi[1] - phi[0])
            phi = [v + shift for v in phi]
    n_phi_eff = 1 if tv == 0 else len(phi)
    # detect_symmetry (mesh.h:120-185)
    reps = [0] + [j for j in range(1, n_th) if m.eps_k_of(theta[j - 1]) != m.eps_k_of(theta[j])
                  or m.Gamma0_of(theta[j - 1]) != m.Gamma0_of(theta[j])]
    # build_time_grid (grid-refinement.h:471-636) with phi_size = 1
    t_end = 1.01 * t_max / (1 + m.z)
    cos_tv, sin_tv, cos_p0 = math.cos(tv), math.sin(tv), math.cos(phi[0])
    t_dec = {r: estimate_t_dec(m, theta[r]) for r in reps}
    min_raw = min_guard = min_cut = t_end
    max_ref, td, ri = 0.0, 0.0, -1
    for j in range(n_th):
        if ri + 1 < len(reps) and reps[ri + 1] == j:
            ri += 1
            td = t_dec[reps[ri]]
        b = _beta(m.Gamma0_of(theta[j]))
        cos_a = math.cos(theta[j]) * cos_tv + math.sin(theta[j]) * sin_tv * cos_p0
        ts = 0.99 * t_min * (1 - b) / (1 - cos_a * b) / (1 + m.z)
        cut = min(0.01 * td, 1e-2 * SEC)
        if is_rvs:
            cut = min(cut, 0.01 * T0)
            max_ref = max(max_ref, 10.0 * max(td, T0))
        min_raw, min_guard, min_cut = min(min_raw, ts), min(min_guard, max(ts, cut)), min(min_cut, cut)
    has_early = min_raw < min_cut
    n_base = int(max(math.log10(t_end / min_guard), 1.0) * t_res)
    extra = int(1.0 * math.log10(min(max_ref, t_end) / min_guard) * t_res) if (is_rvs and max_ref > min_guard) else 0
    n_tot = n_base + extra
    t_rows = []
    for r in reps:
        if is_rvs:
            lat = _cross_lattice(min_guard, t_end, 10 * max(t_dec[r], T0), n_tot, n_base)
        else:
            lat = _band_lattice(min_guard, t_end, t_dec[r] / 3, 3 * t_dec[r], n_tot, 3.0)
        t_rows.append(([min_raw] if has_early else []) + lat)
    return dict(theta=np.array(theta), phi=np.array(phi), reps=np.array(reps), t_rows=np.array(t_rows),
                phi_mirrored=mirror, n_phi_eff=n_phi_eff)


def model_flux(p, t_obs, nu_obs):
    """The whole path (rows a1-a6, a9) restated: grid, dynamics, radiation, EATS.  Returns the forward
    synchrotron flux [n_nu, n_t], or (forward, reverse) with a reverse shock."""
    t_obs = np.asarray(t_obs, float)
    g = auto_grid(p, float(t_obs.min()), float(t_obs.max()))
    return flux_density_grid(p, g["theta"], g["phi"], g["t_rows"], g["reps"], g["phi_mirrored"], g["n_phi_eff"], t_obs, nu_obs)
