// vag_radiation.cuh -- K2: per-cell synchrotron electron population and photon spectrum.
//
// Restates generate_syn_electrons (src/radiation/synchrotron.cpp:315-360 and callees :45-254),
// generate_syn_photons (:376-408), SmoothPowerLawSyn::build / compute_log2_I_nu
// (src/radiation/smooth-power-law-syn.cpp:80-166) for the no-inverse-Compton case
// (Radiation(ssc=False): Ys is default-constructed, Y_c = 0, every IC correction is exactly 1).
//
// Device layout: the reference keeps a 408-byte AoS photon object per (phi,theta,t) cell and
// broadcasts it over the symmetric directions; here only the unique (row,k) cells are stored,
// as a SoA of the PH_NCOEF hot coefficients that compute_log2_I_nu reads.
#pragma once

#include "vag_math.cuh"
#include "vag_model.cuh"

namespace vag {

// radiation-side log2 / exp2 (vag_math.cuh); the dynamics and grid kernels keep libdevice
VAG_HD double rlog2(double x) { return dlog2(x); }
VAG_HD double rexp2(double x) { return dexp2(x); }
// src/util/fast-math.h:179-185
VAG_HD double log2_softplus(double x) {
    if (x > 20.0) return x;
    if (x < -20.0) return 0.0;
    return rlog2(1.0 + rexp2(x));
}
// src/util/fast-math.h:199-202
VAG_HD double log2_broken_power_ratio(double log2_x, double log2_x_break, double s_delta_beta, double s) {
    return -log2_softplus(s_delta_beta * (log2_x - log2_x_break)) / s;
}

// ---- electrons ------------------------------------------------------------------------------
struct SynElectrons {
    double gamma_m, gamma_c, gamma_a, gamma_M, N_e, column_den;
    int regime;
};

// synchrotron.cpp:45-60
VAG_HD bool order3(double a, double b, double c) { return a <= b && b <= c; }
VAG_HD int determine_regime(double a, double c, double m) {
    if (order3(a, m, c)) return 1;
    if (order3(m, a, c)) return 2;
    if (order3(a, c, m)) return 3;
    if (order3(c, a, m)) return 4;
    if (order3(m, c, a)) return 5;
    if (order3(c, m, a)) return 6;
    return 0;
}
// synchrotron.cpp:76-95
VAG_HD double compute_syn_I_peak(double B, double column_den) {
    constexpr double sin_angle_ave = con::pi / 4;
    constexpr double Fx_max = 0.92;
    const double P = B * (sin_angle_ave * Fx_max * con::sqrt3 * con::e3 / (con::me * con::c2));
    return P * column_den / (4 * con::pi);
}
// synchrotron.cpp:107-116
VAG_HD double compute_syn_freq(double gamma, double B) {
    if (B == 0 || !isfinite(gamma)) return 0;
    return 3 * con::e / (4 * con::pi * con::me * con::c) * B * gamma * gamma;
}
VAG_HD double compute_syn_gamma(double nu, double B) {
    return sqrt((4 * con::pi * con::me * con::c / (3 * con::e)) * (nu / B));
}
// synchrotron.cpp:143-148
VAG_HD double compute_syn_gamma_M(double B, double Y) {
    if (B == 0) return kInf;
    return sqrt(6 * con::pi * con::e / con::sigmaT / (B * (1 + Y)));
}
// synchrotron.cpp:164-179 (root_bisect: src/util/utilities.h:201-212)
VAG_HD double compute_syn_gamma_m(double Gamma_th, double gamma_M, double eps_e, double p, double xi) {
    const double gamma_ave_minus_1 = eps_e * (Gamma_th - 1) * (con::mp / con::me) / xi;
    double gamma_m_minus_1 = 1;
    if (p > 2) {
        gamma_m_minus_1 = (p - 2) / (p - 1) * gamma_ave_minus_1;
    } else if (p < 2) {
        gamma_m_minus_1 = pow((2 - p) / (p - 1) * gamma_ave_minus_1 * pow(gamma_M, p - 2), 1 / (p - 1));
    } else {
        auto f = [=](double x) -> double {
            return (x * log(gamma_M) - (x + 1) * log(x) - gamma_ave_minus_1 - log(gamma_M));
        };
        double low = 0, high = gamma_M;
        const double eps = 1e-6;
        for (int iter = 0; iter < 1000 && (high - low) > fabs((high + low) * 0.5) * eps; ++iter) {
            const double mid = 0.5 * (high + low);
            if (f(mid) * f(high) > 0)
                high = mid;
            else
                low = mid;
        }
        gamma_m_minus_1 = 0.5 * (high + low);
    }
    return gamma_m_minus_1 + 1;
}
// synchrotron.cpp:181-188
VAG_HD double compute_gamma_c(double t_comv, double B, double Y) {
    const double gamma_bar = (6 * con::pi * con::me * con::c / con::sigmaT) / (B * B * (1 + Y) * t_comv) * 1;
    return (gamma_bar + sqrt(gamma_bar * gamma_bar + 4)) / 2;
}
// synchrotron.cpp:190-195
VAG_HD double cool_after_crossing(double gamma_x, double gamma_m_x, double gamma_m) {
    const double f_ad = (gamma_m - 1) / (gamma_m_x - 1);
    return (gamma_x - 1) * f_ad + 1;
}
// synchrotron.cpp:212-246 with Ys = InverseComptonY{} and Y_c = 0: every `ic` factor is exactly
// (1+0)/(1+0) = 1 and fast_pow(1, x) = exp2(x*log2(1)) = 1, so those multiplications are dropped.
VAG_HD double compute_syn_gamma_a(double B, double I_syn_peak, double gamma_m, double gamma_c, double p) {
    const double gamma_peak = vmin(gamma_m, gamma_c);
    const double nu_peak = compute_syn_freq(gamma_peak, B);
    const double kT = (gamma_peak - 1) * (con::me * con::c2) / 3;
    double nu_a = fast_pow(I_syn_peak * con::c2 / (cbrt(nu_peak) * 2 * kT), 0.6);
    if (nu_a > nu_peak) {
        if (gamma_c > gamma_m) {
            const double nu_m = compute_syn_freq(gamma_m, B);
            nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * fast_pow(nu_m, p / 2), 2 / (p + 4));
            const double nu_c = compute_syn_freq(gamma_c, B);
            if (nu_a > nu_c) {
                nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * sqrt(nu_c) * fast_pow(nu_m, p / 2), 2 / (p + 5));
            }
        } else {
            const double nu_c = compute_syn_freq(gamma_c, B);
            nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * sqrt(nu_c), 0.4);
            const double nu_m = compute_syn_freq(gamma_m, B);
            if (nu_a > nu_m) {
                nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * sqrt(nu_c) * fast_pow(nu_m, p / 2), 2 / (p + 5));
            }
        }
    }
    return compute_syn_gamma(nu_a, B) + 1;
}
// synchrotron.cpp:248-254
VAG_HD double cyclotron_correction(double gamma_m, double p) {
    double f = (gamma_m - 1) / gamma_m;
    if (p > 3) f = fast_pow(f, (p - 1) / 2);
    return f;
}

// Non-relic part of one cell of generate_syn_electrons (synchrotron.cpp:328-346): gamma_M, gamma_m,
// N_e, column density and the *un-cooled* gamma_c.
VAG_HD void electrons_injection(const RadCfg& rad, double t_comv, double B, double r, double Gamma_th, double N_p,
                                SynElectrons& e) {
    e.gamma_M = compute_syn_gamma_M(B, 0.);
    e.gamma_m = compute_syn_gamma_m(Gamma_th, e.gamma_M, rad.eps_e, rad.p, rad.xi_e);
    const double f_syn = cyclotron_correction(e.gamma_m, rad.p);
    e.N_e = N_p * rad.xi_e * f_syn;
    e.column_den = e.N_e / (r * r);
    e.gamma_c = compute_gamma_c(t_comv, B, 0.);
}

// ---- photons --------------------------------------------------------------------------------
// Hot coefficients of SmoothPowerLawSyn (smooth-power-law-syn.h:24-47), SoA index.
enum PhCoef {
    PH_LOG2_I_MAX = 0,
    PH_LOG2_NU_M,
    PH_LOG2_NU_C,  // kept for introspection (not read by the spectrum without IC)
    PH_LOG2_NU_A,  // idem
    PH_LOG2_NU_M_MAX,  // log2_nu_M
    PH_INV_NU_M_MAX,   // 1/nu_M
    PH_LOG2_NORM,
    PH_LOG2_THICK_NORM,
    PH_S_A_BLEND,
    PH_LOG2_NU_LO,
    PH_LOG2_NU_HI,
    PH_SMOOTH_LO,
    PH_SMOOTH_HI,
    PH_DIFF_LO,
    PH_DIFF_HI,
    PH_NCOEF
};

struct SynPhoton {
    double c[PH_NCOEF];
    double smooth_thick, log2_x_far;  // functions of p only
};

VAG_HD double sigmoid2(double x) { return 1.0 / (1.0 + rexp2(-x)); }
VAG_HD double blend(double w, double a, double b) { return w * a + (1.0 - w) * b; }
VAG_HD double log2_smooth_one(double log2_a, double log2_b, double s) {
    return log2_a - log2_softplus(s * (log2_a - log2_b)) / s;
}

// p-only constants of build(): smooth-power-law-syn.cpp:102-110
VAG_HD void photon_p_consts(double p, double& smooth_thick, double& log2_x_far) {
    smooth_thick = (3.44 * p - 1.41) / con::ln2;
    log2_x_far = 1.5 * rlog2(20.0 / smooth_thick);
}

// sharp forms used for the thick normalisation: smooth-power-law-syn.cpp:48-74
VAG_HD double log2_optical_thick_sharp(double log2_nu, double log2_nu_m) {
    if (log2_nu < log2_nu_m) return 2. * (log2_nu - log2_nu_m);
    return 2.5 * (log2_nu - log2_nu_m);
}
VAG_HD double log2_optical_thin_sharp(double log2_nu, double log2_nu_m, double log2_nu_c, double p) {
    if (log2_nu_m < log2_nu_c) {
        if (log2_nu < log2_nu_m) return (log2_nu - log2_nu_m) / 3.0;
        if (log2_nu < log2_nu_c) return 0.5 * (1.0 - p) * (log2_nu - log2_nu_m);
        return 0.5 * (1.0 - p) * (log2_nu_c - log2_nu_m) - 0.5 * p * (log2_nu - log2_nu_c);
    }
    if (log2_nu < log2_nu_c) return (log2_nu - log2_nu_c) / 3.0;
    if (log2_nu < log2_nu_m) return -0.5 * (log2_nu - log2_nu_c);
    return -0.5 * (log2_nu_m - log2_nu_c) - 0.5 * p * (log2_nu - log2_nu_m);
}

// generate_syn_photons cell body + SmoothPowerLawSyn::build (synchrotron.cpp:388-403,
// smooth-power-law-syn.cpp:94-153)
VAG_HD void build_photon(const SynElectrons& e, double B, double p, double* c /*[PH_NCOEF]*/) {
    const double nu_M = compute_syn_freq(e.gamma_M, B);
    const double nu_m = compute_syn_freq(e.gamma_m, B);
    const double nu_c = compute_syn_freq(e.gamma_c, B);
    const double nu_a = compute_syn_freq(e.gamma_a, B);
    const double I_nu_max = compute_syn_I_peak(B, e.column_den);

    const double log2_nu_m = rlog2(nu_m);
    const double log2_nu_c = rlog2(nu_c);
    const double log2_nu_a = rlog2(nu_a);
    c[PH_LOG2_I_MAX] = rlog2(I_nu_max);
    c[PH_LOG2_NU_M] = log2_nu_m;
    c[PH_LOG2_NU_C] = log2_nu_c;
    c[PH_LOG2_NU_A] = log2_nu_a;
    c[PH_LOG2_NU_M_MAX] = rlog2(nu_M);
    c[PH_INV_NU_M_MAX] = 1.0 / nu_M;

    constexpr double s_swap = 4.0;
    constexpr double s_floor = 0.1;
    const double w_slow = sigmoid2(s_swap * (log2_nu_c - log2_nu_m));
    const double soft_offset = log2_softplus(-s_swap * fabs(log2_nu_c - log2_nu_m)) / s_swap;
    c[PH_LOG2_NU_LO] = vmin(log2_nu_m, log2_nu_c) - soft_offset;
    c[PH_LOG2_NU_HI] = vmax(log2_nu_m, log2_nu_c) + soft_offset;

    const double s_m_slow = vmax(1.84 - 0.40 * p, s_floor);
    const double s_c_slow = vmax(1.15 - 0.06 * p, s_floor);
    constexpr double s_c_fast = 0.597;
    const double s_m_fast = vmax(3.34 - 0.82 * p, s_floor);
    const double smooth_lo = blend(w_slow, s_m_slow, s_c_fast);
    const double smooth_hi = blend(w_slow, s_c_slow, s_m_fast);
    const double alpha_mid = blend(w_slow, -0.5 * (p - 1.0), -0.5);
    c[PH_SMOOTH_LO] = smooth_lo;
    c[PH_SMOOTH_HI] = smooth_hi;
    c[PH_DIFF_LO] = smooth_lo * (1.0 / 3.0 - alpha_mid);
    c[PH_DIFF_HI] = smooth_hi * (alpha_mid + 0.5 * p);

    const double u = sigmoid2(s_swap * (log2_nu_a - log2_nu_m));
    const double v = sigmoid2(s_swap * (log2_nu_a - log2_nu_c));
    const double w_below = (1.0 - u) * (1.0 - v);
    const double w_above = u * v;
    constexpr double s_a_below = 1.64;
    const double s_a_mid = vmax(1.47 - 0.21 * p, s_floor);
    const double s_a_above = vmax(0.94 - 0.14 * p, s_floor);
    c[PH_S_A_BLEND] = w_below * s_a_below + w_above * s_a_above + (1.0 - w_below - w_above) * s_a_mid;

    c[PH_LOG2_NORM] = 1.0 / smooth_lo;
    c[PH_LOG2_THICK_NORM] =
        log2_optical_thin_sharp(log2_nu_a, log2_nu_m, log2_nu_c, p) - log2_optical_thick_sharp(log2_nu_a, log2_nu_m);
}

// SmoothPowerLawSyn::compute_log2_I_nu (smooth-power-law-syn.cpp:26-46,80-92,159-166) without IC.
// `get(i)` returns coefficient i of the cell (lets the caller read registers, shared or global).
template <class Get>
VAG_HD double photon_log2_I_nu(const Get& get, double smooth_thick, double log2_x_far, double log2_nu) {
    // log2_optical_thin
    const double lo = get(PH_LOG2_NU_LO);
    const double thin = (log2_nu - lo) / 3.0 + log2_broken_power_ratio(log2_nu, lo, get(PH_DIFF_LO), get(PH_SMOOTH_LO)) +
                        log2_broken_power_ratio(log2_nu, get(PH_LOG2_NU_HI), get(PH_DIFF_HI), get(PH_SMOOTH_HI));
    // log2_optical_thick
    const double log2_x = log2_nu - get(PH_LOG2_NU_M);
    double thick;
    if (log2_x > log2_x_far) {
        thick = 2.5 * log2_x;
    } else {
        const double s = -smooth_thick * rexp2(2. / 3 * log2_x);
        thick = 2.5 * log2_x + log2_softplus(-0.5 * log2_x + s);
    }
    const double spec = get(PH_LOG2_I_MAX) +
                        (get(PH_LOG2_NORM) + log2_smooth_one(thin, thick + get(PH_LOG2_THICK_NORM), get(PH_S_A_BLEND)));
    if (log2_nu - get(PH_LOG2_NU_M_MAX) < -20) return spec;
    return spec - con::log2e * get(PH_INV_NU_M_MAX) * rexp2(log2_nu);
}

// The same spectrum for the EATS hot loop: coefficients already in registers (`c`, indexed by PhCoef),
// log2_softplus through the shared-memory table, divisions by cell constants as multiplications by
// their reciprocals (<= 1 ulp per term against photon_log2_I_nu).
struct SynCoefRegs {
    double log2_I_max, log2_nu_m, log2_nu_M, inv_nu_M, inv_smooth_lo, log2_thick_norm, s_a, log2_nu_lo, log2_nu_hi,
        smooth_hi, diff_lo, diff_hi;
    double inv_smooth_hi, inv_s_a;
};
// TILE: the record serves a whole frequency tile (phase 1 of the grid / banded kernels): the two reciprocals then use
// the branch-free form (-1 % on the grid kernel).  The per-point series path evaluates one spectrum per load and is
// MUFU-bound: there the IEEE divisions are the faster choice (branch-free: +15 % on the series kernel).
template <bool TILE = false, class Get>
VAG_HD SynCoefRegs load_syn_coefs(const Get& get) {
    SynCoefRegs r;
    r.log2_I_max = get(PH_LOG2_I_MAX);
    r.log2_nu_m = get(PH_LOG2_NU_M);
    r.log2_nu_M = get(PH_LOG2_NU_M_MAX);
    r.inv_nu_M = get(PH_INV_NU_M_MAX);
    r.inv_smooth_lo = get(PH_LOG2_NORM);
    r.log2_thick_norm = get(PH_LOG2_THICK_NORM);
    r.s_a = get(PH_S_A_BLEND);
    r.log2_nu_lo = get(PH_LOG2_NU_LO);
    r.log2_nu_hi = get(PH_LOG2_NU_HI);
    r.smooth_hi = get(PH_SMOOTH_HI);
    r.diff_lo = get(PH_DIFF_LO);
    r.diff_hi = get(PH_DIFF_HI);
    r.inv_smooth_hi = TILE ? vdiv(1.0, r.smooth_hi) : 1.0 / r.smooth_hi;  // both >= s_floor = 0.1
    r.inv_s_a = TILE ? vdiv(1.0, r.s_a) : 1.0 / r.s_a;
    return r;
}
VAG_HD double photon_log2_I_nu_fast(const SynCoefRegs& c, const double* __restrict__ sp_lut, double smooth_thick,
                                    double log2_x_far, double log2_nu) {
    const double dlo = log2_nu - c.log2_nu_lo;
    const double thin = dlo * (1.0 / 3.0) - log2_softplus_lut(sp_lut, c.diff_lo * dlo) * c.inv_smooth_lo -
                        log2_softplus_lut(sp_lut, c.diff_hi * (log2_nu - c.log2_nu_hi)) * c.inv_smooth_hi;
    const double log2_x = log2_nu - c.log2_nu_m;
    double thick = 2.5 * log2_x;
    if (!(log2_x > log2_x_far)) {
        const double s = -smooth_thick * rexp2((2. / 3) * log2_x);
        thick += log2_softplus_lut(sp_lut, -0.5 * log2_x + s);
    }
    const double b = thick + c.log2_thick_norm;
    const double smooth = thin - log2_softplus_lut(sp_lut, c.s_a * (thin - b)) * c.inv_s_a;
    const double spec = c.log2_I_max + (c.inv_smooth_lo + smooth);
    if (log2_nu - c.log2_nu_M < -20) return spec;
    return spec - con::log2e * c.inv_nu_M * rexp2(log2_nu);
}

// photon_log2_I_nu_fast with the two exponentials of the point supplied by the caller: x23 = 2^(2/3 log2_x) and
// cut = log2(e) 2^log2_nu / nu_M.  A frequency tile of one cell shares them up to a per-frequency factor
// (eats_phase1), so they cost one multiplication each instead of one exp2 each (identities; <= 2 ulp per factor).
VAG_HD double photon_log2_I_nu_tile(const SynCoefRegs& c, const double* __restrict__ sp_lut, double smooth_thick,
                                    double log2_x_far, double log2_nu, double x23, double cut) {
    const double dlo = log2_nu - c.log2_nu_lo;
    const double thin = dlo * (1.0 / 3.0) - log2_softplus_lut(sp_lut, c.diff_lo * dlo) * c.inv_smooth_lo -
                        log2_softplus_lut(sp_lut, c.diff_hi * (log2_nu - c.log2_nu_hi)) * c.inv_smooth_hi;
    const double log2_x = log2_nu - c.log2_nu_m;
    double thick = 2.5 * log2_x;
    if (!(log2_x > log2_x_far)) thick += log2_softplus_lut(sp_lut, -0.5 * log2_x - smooth_thick * x23);
    const double b = thick + c.log2_thick_norm;
    const double smooth = thin - log2_softplus_lut(sp_lut, c.s_a * (thin - b)) * c.inv_s_a;
    const double spec = c.log2_I_max + (c.inv_smooth_lo + smooth);
    return (log2_nu - c.log2_nu_M < -20) ? spec : spec - cut;
}

// One cell of K2: shock state -> photon coefficients.  For relic cells (k >= injection_idx,
// synchrotron.h:187-201) the injection-time cell k_inj-1 is re-derived from its shock state
// instead of being read from a neighbour's output, so every cell is independent.
struct CellShock {
    double t_comv, r, Gamma_th, B, N_p;
};

VAG_HD void radiation_cell(const RadCfg& rad, const CellShock& cs, bool relic, const CellShock& inj_cs,
                           double* coef /*[PH_NCOEF]*/, SynElectrons* e_out = nullptr) {
    SynElectrons e;
    electrons_injection(rad, cs.t_comv, cs.B, cs.r, cs.Gamma_th, cs.N_p, e);
    const double I_nu_peak = compute_syn_I_peak(cs.B, e.column_den);
    if (relic) {
        SynElectrons inj;
        electrons_injection(rad, inj_cs.t_comv, inj_cs.B, inj_cs.r, inj_cs.Gamma_th, inj_cs.N_p, inj);
        e.gamma_c = cool_after_crossing(inj.gamma_c, inj.gamma_m, e.gamma_m);
        e.gamma_M = cool_after_crossing(inj.gamma_M, inj.gamma_m, e.gamma_m);
    }
    e.gamma_a = compute_syn_gamma_a(cs.B, I_nu_peak, e.gamma_m, e.gamma_c, rad.p);
    e.regime = determine_regime(e.gamma_a, e.gamma_c, e.gamma_m);
    build_photon(e, cs.B, rad.p, coef);
    if (e_out) *e_out = e;
}

}  // namespace vag
