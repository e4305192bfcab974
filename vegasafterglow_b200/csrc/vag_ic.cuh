// vag_ic.cuh -- inverse-Compton cooling of the electron population and the per-cell
// synchrotron-self-Compton (SSC) spectrum, Thomson or Klein-Nishina.
//
// Restates (from scratch):
//   InverseComptonY (Y(gamma) broken power law)      src/radiation/inverse-compton.h:28-91,
//                                                    src/radiation/inverse-compton.cpp:18-187
//   BrokenPowerLaw<5>                                src/util/utilities.h:23-69
//   IC_cooling + update_gamma_c_{Thomson,KN}, update_gamma_M
//                                                    inverse-compton.h:729-764, inverse-compton.cpp:192-251
//   compute_syn_gamma_a with IC                      src/radiation/synchrotron.cpp:212-246
//   SynElectrons::compute_spectrum/column_den        synchrotron.cpp:260-299
//   IC-corrected synchrotron spectrum                smooth-power-law-syn.cpp:80-92, inverse-compton.h:781-792
//   KN cross-section + 128-entry LUT                 inverse-compton.cpp:257-378
//   ICPhoton spectrum on commensurate log lattices   inverse-compton.h:270-647
//
// Device structure: IC cooling is sequential in k along a row (cell k starts its fixed point from
// gamma_c of cell k-1, inverse-compton.h:746) -> one thread per (row, shock).  The SSC spectrum of a
// cell is an O(n_gamma (n_seed + n_out)) accumulation over lattices that differ per cell -> one warp
// per cell with the `Par` pattern of vag_grid.cuh (uniform code + lane-strided regions on scratch).
#pragma once

#include "vag_grid.cuh"  // Par executors
#include "vag_radiation.cuh"

namespace vag {

namespace con {
constexpr double h = 6.63e-27 * unit::erg * unit::sec;  // src/util/macros.h:93
}

// ---------------------------------------------------------------------------------------------
// InverseComptonY
// ---------------------------------------------------------------------------------------------
struct ICY {
    double gamma_m_hat, gamma_c_hat, gamma_self3, Y_T;
    double gamma_m_, B_, p_;
    int regime, nseg;
    double slope[5], log2_lower[5], log2_const[5];  // BrokenPowerLaw<5> segments
};

VAG_HD void icy_default(ICY& y) {  // InverseComptonY::InverseComptonY() inverse-compton.cpp:38-44
    y.gamma_m_hat = 1.0;
    y.gamma_c_hat = 1.0;
    y.gamma_self3 = 1.0;
    y.Y_T = 0.0;
    y.gamma_m_ = 1.0;
    y.B_ = 0.0;
    y.p_ = 2.3;
    y.regime = 0;
    y.nseg = 0;
}
// BrokenPowerLaw::first_segment / add_segment: utilities.h:27-47
VAG_HD void bpl_first(ICY& y, double norm, double lower, double slope) {
    const double l = fast_log2(lower), v = fast_log2(norm);
    y.nseg = 1;
    y.slope[0] = slope;
    y.log2_lower[0] = l;
    y.log2_const[0] = v - slope * l;
}
VAG_HD void bpl_add(ICY& y, double lower, double slope) {
    const double l = fast_log2(lower);
    const int prev = y.nseg - 1;
    const double v = y.log2_const[prev] + y.slope[prev] * l;
    const int i = y.nseg++;
    y.slope[i] = slope;
    y.log2_lower[i] = l;
    y.log2_const[i] = v - slope * l;
}
// BrokenPowerLaw::eval: utilities.h:49-57
VAG_HD double icy_gamma_spectrum(const ICY& y, double x) {
    const double log2_x = fast_log2(x);
    for (int i = y.nseg - 1; i > 0; --i)
        if (log2_x >= y.log2_lower[i]) return fast_exp2(y.log2_const[i] + y.slope[i] * log2_x);
    return (y.nseg > 0) ? fast_exp2(y.log2_const[0] + y.slope[0] * log2_x) : 0.0;
}
VAG_HD double icy_nu_spectrum(const ICY& y, double nu) { return icy_gamma_spectrum(y, compute_syn_gamma(nu, y.B_)); }
VAG_HD double icy_gamma_hat(const ICY& y, double gamma) { return vmax(y.gamma_self3 / (gamma * gamma), 1.0); }

// build_segments: inverse-compton.cpp:117-175 (regimes 0-2 are the reachable ones: :95-115)
VAG_HD void icy_build_segments(ICY& y) {
    switch (y.regime) {
        case 0:
            bpl_first(y, y.Y_T, 1.0, 0.0);
            break;
        case 1:
            bpl_first(y, y.Y_T, 1.0, 0.0);
            bpl_add(y, y.gamma_c_hat, 0.5 * (y.p_ - 3.0));
            bpl_add(y, y.gamma_m_hat, -4.0 / 3.0);
            break;
        default:
            bpl_first(y, y.Y_T, 1.0, 0.0);
            bpl_add(y, y.gamma_m_hat, -0.5);
            bpl_add(y, y.gamma_c_hat, -4.0 / 3.0);
            break;
    }
}
// update_cooling_breaks: inverse-compton.cpp:90-115
VAG_HD void icy_update_cooling_breaks(ICY& y, double gamma_c, double Y_T) {
    y.gamma_c_hat = icy_gamma_hat(y, gamma_c);
    y.Y_T = Y_T;
    y.regime = (y.gamma_m_ < gamma_c) ? 1 : 2;
    icy_build_segments(y);
}
// InverseComptonY(gamma_m, gamma_c, p, B, Y_T, is_KN): inverse-compton.cpp:18-36
VAG_HD void icy_init(ICY& y, double gamma_m, double gamma_c, double p, double B, double Y_T, bool is_KN) {
    const double nu_m = compute_syn_freq(gamma_m, B);
    y.gamma_m_hat = vmax(con::me * con::c2 / con::h / nu_m, 1.0);
    const double gamma_self = fast_pow(y.gamma_m_hat * gamma_m * gamma_m, 1.0 / 3.0);
    y.gamma_self3 = gamma_self * gamma_self * gamma_self;
    y.B_ = B;
    y.gamma_m_ = gamma_m;
    y.p_ = p;
    y.nseg = 0;
    if (is_KN) {
        icy_update_cooling_breaks(y, gamma_c, Y_T);
    } else {
        y.gamma_c_hat = icy_gamma_hat(y, gamma_c);
        y.Y_T = Y_T;
        y.regime = 0;
        icy_build_segments(y);
    }
}

// eta_rad_Thomson / compute_Thomson_Y: inverse-compton.h:692-704
VAG_HD double compute_Thomson_Y(const RadCfg& rad, double gamma_m, double gamma_c) {
    const double eta_e = (gamma_c < gamma_m) ? 1 : fast_pow(gamma_c / gamma_m, 2 - rad.p);
    const double b = eta_e * rad.eps_e / rad.eps_B;
    return 0.5 * (sqrt(1. + 4. * b) - 1.);
}
// update_gamma_c_Thomson: inverse-compton.cpp:192-204
VAG_HD void update_gamma_c_Thomson(double& gamma_c, ICY& Ys, const RadCfg& rad, double B, double t_com, double gamma_m,
                                   double gamma_c_last) {
    double Y_T = compute_Thomson_Y(rad, gamma_m, gamma_c);
    double gamma_c_new = gamma_c_last;
    while (fabs((gamma_c_new - gamma_c) / gamma_c) > 1e-3) {
        gamma_c = gamma_c_new;
        Y_T = compute_Thomson_Y(rad, gamma_m, gamma_c);
        gamma_c_new = compute_gamma_c(t_com, B, Y_T);
    }
    gamma_c = gamma_c_new;
    icy_init(Ys, gamma_m, gamma_c, rad.p, B, Y_T, false);
}
// update_gamma_c_KN: inverse-compton.cpp:206-235
VAG_HD void update_gamma_c_KN(double& gamma_c, ICY& Ys, const RadCfg& rad, double B, double t_com, double gamma_m,
                              double gamma_c_last) {
    double gamma_c_new = gamma_c_last;
    double Y_T = compute_Thomson_Y(rad, gamma_m, gamma_c_new);
    icy_init(Ys, gamma_m, gamma_c_new, rad.p, B, Y_T, true);
    int iter = 0;
    do {
        gamma_c = gamma_c_new;
        Y_T = compute_Thomson_Y(rad, gamma_m, gamma_c);
        icy_update_cooling_breaks(Ys, gamma_c, Y_T);
        const double Y_c = icy_gamma_spectrum(Ys, gamma_c);
        gamma_c_new = compute_gamma_c(t_com, B, Y_c);
        iter++;
    } while (fabs((gamma_c_new - gamma_c) / gamma_c) > 1e-3 && iter < 100);
    gamma_c = gamma_c_new;
}
// update_gamma_M: inverse-compton.cpp:237-251
VAG_HD void update_gamma_M(double& gamma_M, const ICY& Ys, double B) {
    if (B == 0) {
        gamma_M = kInf;
        return;
    }
    double Y_M = icy_gamma_spectrum(Ys, gamma_M);
    double gamma_M_new = compute_syn_gamma_M(B, Y_M);
    while (fabs((gamma_M - gamma_M_new) / gamma_M_new) > 1e-3) {
        gamma_M = gamma_M_new;
        Y_M = icy_gamma_spectrum(Ys, gamma_M);
        gamma_M_new = compute_syn_gamma_M(B, Y_M);
    }
}
// compute_syn_gamma_a with a populated Ys: synchrotron.cpp:212-246
VAG_HD double compute_syn_gamma_a_ic(double B, double I_syn_peak, double gamma_m, double gamma_c, double p, const ICY& Ys,
                                     double Y_c) {
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
                const double ic = (1 + Y_c) / (1 + icy_nu_spectrum(Ys, nu_a));
                nu_a *= fast_pow(ic, 2 / (p + 5));
            }
        } else {
            const double nu_c = compute_syn_freq(gamma_c, B);
            nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * sqrt(nu_c), 0.4);
            double ic = (1 + Y_c) / (1 + icy_nu_spectrum(Ys, nu_a));
            nu_a *= fast_pow(ic, 0.4);
            const double nu_m = compute_syn_freq(gamma_m, B);
            if (nu_a > nu_m) {
                nu_a = fast_pow(I_syn_peak * con::c2 / (2 * kT) * sqrt(nu_c) * fast_pow(nu_m, p / 2), 2 / (p + 5));
                ic = (1 + Y_c) / (1 + icy_nu_spectrum(Ys, nu_a));
                nu_a *= fast_pow(ic, 2 / (p + 5));
            }
        }
    }
    return compute_syn_gamma(nu_a, B) + 1;
}

// ---------------------------------------------------------------------------------------------
// per-cell record of a shock with ssc=True (cold data: only read by the IC paths)
// ---------------------------------------------------------------------------------------------
struct IcCell {
    double gamma_m, gamma_c, gamma_a, gamma_M, column_den, Y_c, p;
    double nu_a, nu_m, nu_M;  // linear photon breaks (ICPhoton grid bounds)
    int regime, pad_;
    ICY ys;
};

// SynElectrons::compute_spectrum / compute_column_den: synchrotron.cpp:260-299
VAG_HD double electron_spectrum(const IcCell& e, double gamma) {
    switch (e.regime) {
        case 1:
        case 2:
        case 5:
            return (e.p - 1) / e.gamma_m * fast_exp(-gamma / e.gamma_M - e.gamma_m / gamma) *
                   fast_pow(gamma / e.gamma_m, -e.p) * e.gamma_c / (gamma + e.gamma_c);
        case 3:
        case 4:
        case 6:
            return fast_exp(-gamma / e.gamma_M - e.gamma_c / gamma) * e.gamma_c / (gamma * gamma) /
                   (1.0 + fast_pow(gamma / e.gamma_m, e.p - 1));
        default:
            return 0;
    }
}
VAG_HD double electron_column_den(const IcCell& e, double gamma) {
    if (gamma <= e.gamma_c) return e.column_den * electron_spectrum(e, gamma);
    return e.column_den * electron_spectrum(e, gamma) * (1 + e.Y_c) / (1 + icy_gamma_spectrum(e.ys, gamma));
}

// SmoothPowerLawSyn::compute_log2_I_nu WITH the IC correction of the thin branch
// (smooth-power-law-syn.cpp:80-92,159-166; has_IC_correction / inverse_compton_correction
// inverse-compton.h:773-792)
template <class Get>
VAG_HD double photon_log2_I_nu_ic(const Get& get, double smooth_thick, double log2_x_far, const IcCell& c,
                                  double log2_nu) {
    const double lo = get(PH_LOG2_NU_LO);
    double thin = (log2_nu - lo) / 3.0 + log2_broken_power_ratio(log2_nu, lo, get(PH_DIFF_LO), get(PH_SMOOTH_LO)) +
                  log2_broken_power_ratio(log2_nu, get(PH_LOG2_NU_HI), get(PH_DIFF_HI), get(PH_SMOOTH_HI));
    const double log2_x = log2_nu - get(PH_LOG2_NU_M);
    double thick;
    if (log2_x > log2_x_far) {
        thick = 2.5 * log2_x;
    } else {
        const double s = -smooth_thick * rexp2(2. / 3 * log2_x);
        thick = 2.5 * log2_x + log2_softplus(-0.5 * log2_x + s);
    }
    if (log2_nu > get(PH_LOG2_NU_C) && (c.Y_c > 0 || c.ys.Y_T > 0)) {
        const double nu = rexp2(log2_nu);
        thin += rlog2((1. + c.Y_c) / (1 + icy_nu_spectrum(c.ys, nu)));
    }
    const double spec = get(PH_LOG2_I_MAX) +
                        (get(PH_LOG2_NORM) + log2_smooth_one(thin, thick + get(PH_LOG2_THICK_NORM), get(PH_S_A_BLEND)));
    if (log2_nu - get(PH_LOG2_NU_M_MAX) < -20) return spec;
    return spec - con::log2e * get(PH_INV_NU_M_MAX) * rexp2(log2_nu);
}

// The same spectrum for a frequency tile of one cell (eats_phase1): coefficients in registers, log2_softplus through the
// shared-memory table, the two exponentials supplied as per-cell x per-frequency products (photon_log2_I_nu_tile);
// nu_lin = 2^log2_nu is the comoving frequency the IC correction is evaluated at.
VAG_HD double photon_log2_I_nu_ic_tile(const SynCoefRegs& c, double log2_nu_c, const double* __restrict__ sp_lut,
                                       double smooth_thick, double log2_x_far, const IcCell& ic, double log2_nu, double x23,
                                       double cut, double nu_lin) {
    const double dlo = log2_nu - c.log2_nu_lo;
    double thin = dlo * (1.0 / 3.0) - log2_softplus_lut(sp_lut, c.diff_lo * dlo) * c.inv_smooth_lo -
                  log2_softplus_lut(sp_lut, c.diff_hi * (log2_nu - c.log2_nu_hi)) * c.inv_smooth_hi;
    const double log2_x = log2_nu - c.log2_nu_m;
    double thick = 2.5 * log2_x;
    if (!(log2_x > log2_x_far)) thick += log2_softplus_lut(sp_lut, -0.5 * log2_x - smooth_thick * x23);
    if (log2_nu > log2_nu_c && (ic.Y_c > 0 || ic.ys.Y_T > 0)) thin += rlog2((1. + ic.Y_c) / (1 + icy_nu_spectrum(ic.ys, nu_lin)));
    const double b = thick + c.log2_thick_norm;
    const double smooth = thin - log2_softplus_lut(sp_lut, c.s_a * (thin - b)) * c.inv_s_a;
    const double spec = c.log2_I_max + (c.inv_smooth_lo + smooth);
    return (log2_nu - c.log2_nu_M < -20) ? spec : spec - cut;
}

// ---------------------------------------------------------------------------------------------
// One row of generate_syn_electrons + IC_cooling + generate_syn_photons for a shock with ssc
// (synchrotron.cpp:315-408, inverse-compton.h:729-764).  Sequential in k.
//   sh_*: the row's shock planes; coef_out(k, c) stores photon coefficient c of cell k.
// ---------------------------------------------------------------------------------------------
template <class CoefOut>
VAG_HD void ic_cool_row(const RadCfg& rad, int n_t, int inj_idx, const double* t_comv, const double* r,
                        const double* Gamma_th, const double* B_arr, const double* N_p, IcCell* cells,
                        const CoefOut& coef_out) {
    const bool kn = rad.kn != 0;
    // first-pass (no IC) values of the injection-time cell k_inj-1, needed by first-pass relic cooling
    SynElectrons inj0;
    inj0.gamma_c = inj0.gamma_m = inj0.gamma_M = 1;
    if (inj_idx < n_t && inj_idx >= 1) {
        const int ki = inj_idx - 1;
        electrons_injection(rad, t_comv[ki], B_arr[ki], r[ki], Gamma_th[ki], N_p[ki], inj0);
    }
    double gamma_c_prev = 0;                      // updated gamma_c of cell k-1
    double injc = 1, injm = 1, injM = 1;          // IC-updated values of cell k_inj-1
    for (int k = 0; k < n_t; ++k) {
        const double B = B_arr[k], t_com = t_comv[k];
        // ---- generate_syn_electrons (first pass) ------------------------------------------------
        SynElectrons e;
        electrons_injection(rad, t_com, B, r[k], Gamma_th[k], N_p[k], e);
        const bool relic = k >= inj_idx;
        if (relic) {
            e.gamma_c = cool_after_crossing(inj0.gamma_c, inj0.gamma_m, e.gamma_m);
            e.gamma_M = cool_after_crossing(inj0.gamma_M, inj0.gamma_m, e.gamma_m);
        }
        // ---- IC_cooling ---------------------------------------------------------------------------
        IcCell& c = cells[k];
        const double gamma_c_last = (k > 0) ? gamma_c_prev : e.gamma_c;
        if (kn)
            update_gamma_c_KN(e.gamma_c, c.ys, rad, B, t_com, e.gamma_m, gamma_c_last);
        else
            update_gamma_c_Thomson(e.gamma_c, c.ys, rad, B, t_com, e.gamma_m, gamma_c_last);
        update_gamma_M(e.gamma_M, c.ys, B);
        if (relic) {  // cool_relic_electrons with the already-updated injection cell
            e.gamma_c = cool_after_crossing(injc, injm, e.gamma_m);
            e.gamma_M = cool_after_crossing(injM, injm, e.gamma_m);
        }
        const double I_nu_peak = compute_syn_I_peak(B, e.column_den);
        const double Y_c = icy_gamma_spectrum(c.ys, e.gamma_c);
        e.gamma_a = compute_syn_gamma_a_ic(B, I_nu_peak, e.gamma_m, e.gamma_c, rad.p, c.ys, Y_c);
        e.regime = determine_regime(e.gamma_a, e.gamma_c, e.gamma_m);
        gamma_c_prev = e.gamma_c;
        if (k == inj_idx - 1) {
            injc = e.gamma_c;
            injm = e.gamma_m;
            injM = e.gamma_M;
        }
        c.gamma_m = e.gamma_m;
        c.gamma_c = e.gamma_c;
        c.gamma_a = e.gamma_a;
        c.gamma_M = e.gamma_M;
        c.column_den = e.column_den;
        c.Y_c = Y_c;
        c.p = rad.p;
        c.regime = e.regime;
        c.nu_a = compute_syn_freq(e.gamma_a, B);
        c.nu_m = compute_syn_freq(e.gamma_m, B);
        c.nu_M = compute_syn_freq(e.gamma_M, B);
        // ---- generate_syn_photons -------------------------------------------------------------------
        double coef[PH_NCOEF];
        build_photon(e, B, rad.p, coef);
        for (int q = 0; q < PH_NCOEF; ++q) coef_out(k, q, coef[q]);
    }
}

// ---------------------------------------------------------------------------------------------
// Klein-Nishina cross-section ratio and its LUT: inverse-compton.cpp:257-378
// ---------------------------------------------------------------------------------------------
VAG_HD double compton_ratio_from_x(double x) {
    if (x < 1e-2) return 1 - 2 * x;
    if (x > 1e2) return 3. / 8 * (log(2 * x) + 0.5) / x;
    const double l = log1p(2.0 * x);
    const double invx = 1.0 / x;
    const double invx2 = invx * invx;
    const double term1 = 1.0 + 2.0 * x;
    const double invt1 = 1.0 / term1;
    const double invt1_2 = invt1 * invt1;
    const double a = (1.0 + x) * invx2 * invx;
    const double b = 2.0 * x * (1.0 + x) * invt1 - l;
    const double c = 0.5 * l * invx;
    const double d = (1.0 + 3.0 * x) * invt1_2;
    return 0.75 * (a * b + c - d);
}
constexpr int KN_LUT_N = 128;
constexpr double KN_LG2_X_MIN = -6.6438561897747247, KN_LG2_X_MAX = 6.6438561897747247;
struct KnLut {
    double ratio[KN_LUT_N], lg2_ratio[KN_LUT_N];
};
VAG_HD void kn_lut_entry(int i, double& ratio, double& lg2_ratio) {
    const double step = (KN_LG2_X_MAX - KN_LG2_X_MIN) / (double)(KN_LUT_N - 1);
    ratio = compton_ratio_from_x(fast_exp2(KN_LG2_X_MIN + step * (double)i));
    lg2_ratio = fast_log2(ratio);
}
// compton_correction_pair: inverse-compton.cpp:353-378
VAG_HD void compton_correction_pair(const KnLut& lut, double nu, double& corr, double& lg2_corr) {
    constexpr double inv_ln2 = 1.4426950408889634;
    const double x = con::h / (con::me * con::c2) * nu;
    if (!(x > 0)) {
        corr = 0;
        lg2_corr = -kInf;
        return;
    }
    if (x <= 1e-2) {
        corr = 1 - 2 * x;
        lg2_corr = -(2 * x + 2 * x * x) * inv_ln2;
        return;
    }
    if (x >= 1e2) {
        corr = compton_ratio_from_x(x);
        lg2_corr = fast_log2(corr);
        return;
    }
    const double step = (KN_LG2_X_MAX - KN_LG2_X_MIN) / (double)(KN_LUT_N - 1);
    const double inv_step = 1.0 / step;
    const double pos = (fast_log2(x) - KN_LG2_X_MIN) * inv_step;
    if (pos <= 0) {
        corr = lut.ratio[0];
        lg2_corr = lut.lg2_ratio[0];
        return;
    }
    if (pos >= (double)(KN_LUT_N - 1)) {
        corr = lut.ratio[KN_LUT_N - 1];
        lg2_corr = lut.lg2_ratio[KN_LUT_N - 1];
        return;
    }
    const int idx = (int)pos;
    const double frac = pos - (double)idx;
    corr = lut.ratio[idx] + (lut.ratio[idx + 1] - lut.ratio[idx]) * frac;
    lg2_corr = lut.lg2_ratio[idx] + (lut.lg2_ratio[idx + 1] - lut.lg2_ratio[idx]) * frac;
}

// ---------------------------------------------------------------------------------------------
// ICPhoton: per-cell SSC spectrum (inverse-compton.h:270-607)
// ---------------------------------------------------------------------------------------------
constexpr int IC_CAP_SEED = 192;   // seed-frequency lattice nodes
constexpr int IC_CAP_GAMMA = 96;   // electron lattice nodes
constexpr int IC_CAP_OUT = 160;    // output lattice nodes (clamped to the observer band)
constexpr int IC_CAP_LAT = 2 * IC_CAP_SEED + 2 * IC_CAP_GAMMA;
constexpr double IC_x0 = 0.47140452079103166;
constexpr double IC_Q = 3.321928094887362 / 8;  // lattice quantum log2(10)/8

struct IcTable {  // per cell, per shock: header + log2_I_nu_IC[IC_CAP_OUT]
    double phase;            // log2_nu_IC(k) = phase + IC_Q * (idx0 + 2 k)
    double lg2_theory_min, lg2_theory_max;
    int idx0, n;
};

// warp scratch (doubles): nu_seed, lg2_nu_seed, dnu_seed, fv_th, lg2fv_th, lg2r, inv_lg2r, cdf_th,
// ratio_th, fv_buf, cdf_buf, ratio_buf [12 x IC_CAP_SEED], gamma, dN_e [2 x IC_CAP_GAMMA],
// corr_lat, lg2corr_lat [2 x IC_CAP_LAT], I_buf [IC_CAP_OUT]
constexpr int IC_SCRATCH_DOUBLES = 12 * IC_CAP_SEED + 2 * IC_CAP_GAMMA + 2 * IC_CAP_LAT + IC_CAP_OUT;

// power_law_bin_integral: inverse-compton.h:406-419
VAG_HD double power_law_bin_integral(double f_lo, double f_hi, double nu_lo, double nu_hi, double lg2f_lo, double lg2f_hi,
                                     double lg2r, double inv_lg2r, double trap) {
    if (!(f_lo > 0) || !(f_hi > 0)) return trap;
    const double s1 = 1 + (lg2f_hi - lg2f_lo) * inv_lg2r;
    if (fabs(s1) > 1e-3) return (f_hi * nu_hi - f_lo * nu_lo) / s1;
    return f_lo * nu_lo * lg2r * con::ln2;
}

// Generates the table of one cell.  Returns 0, or a VAG_ST_* bit when a lattice exceeds its capacity.
// `seed_log2_I(log2_nu)` evaluates the cell's (IC-corrected) synchrotron spectrum.
template <class Par, class Seed>
VAG_HD int ic_generate(const Par& par, const IcCell& c, const Seed& seed_log2_I, bool KN, const KnLut& lut,
                       double nu_eval_min, double nu_eval_max, double* S, IcTable& hdr, double* tab) {
    hdr.n = 0;
    hdr.idx0 = 0;
    hdr.phase = 0;
    // compute_grid_params: inverse-compton.h:291-331
    const double tail_factor = vmax(-log(1e-2), 5.0);
    const double gamma_min = vmin(c.gamma_m, c.gamma_c) / 30;
    const double gamma_max = vmax(c.gamma_M * tail_factor, gamma_min);
    const double nu_min = vmin(c.nu_a, c.nu_m) / 10;
    const double nu_max = vmax(c.nu_M * tail_factor, nu_min);
    double nu_IC_min = 4 * IC_x0 * nu_min * gamma_min * gamma_min;
    const double nu_ic_base = 4 * IC_x0 * c.nu_M * c.gamma_M * c.gamma_M;
    const double nu_ic_single_cut = vmax(nu_ic_base * tail_factor * tail_factor, nu_ic_base * tail_factor);
    double nu_IC_max = nu_ic_single_cut * 2.0;
    hdr.lg2_theory_max = fast_log2(nu_IC_max);
    hdr.lg2_theory_min = fast_log2(nu_IC_min);
    nu_IC_min = vmax(nu_IC_min, vmin(nu_eval_min / 4.0, nu_IC_max / 16.0));
    nu_IC_max = vmin(nu_IC_max, vmax(nu_eval_max * 4.0, nu_IC_min * 16.0));
    auto posfin = [](double x) { return isfinite(x) && x > 0; };
    if (!(posfin(gamma_min) && posfin(gamma_max) && posfin(nu_min) && posfin(nu_max) && posfin(nu_IC_min) &&
          posfin(nu_IC_max)))
        return 0;  // generated = true with an empty grid: every query returns -inf

    double* nu_seed = S;
    double* lg2_nu_seed = nu_seed + IC_CAP_SEED;
    double* dnu_seed = lg2_nu_seed + IC_CAP_SEED;
    double* fv_th = dnu_seed + IC_CAP_SEED;
    double* lg2fv_th = fv_th + IC_CAP_SEED;
    double* lg2r = lg2fv_th + IC_CAP_SEED;
    double* inv_lg2r = lg2r + IC_CAP_SEED;
    double* cdf_th = inv_lg2r + IC_CAP_SEED;
    double* ratio_th = cdf_th + IC_CAP_SEED;
    double* fv_buf = ratio_th + IC_CAP_SEED;
    double* cdf_buf = fv_buf + IC_CAP_SEED;
    double* ratio_buf = cdf_buf + IC_CAP_SEED;
    double* gamma = ratio_buf + IC_CAP_SEED;
    double* dN_e = gamma + IC_CAP_GAMMA;
    double* corr_lat = dN_e + IC_CAP_GAMMA;
    double* lg2corr_lat = corr_lat + IC_CAP_LAT;
    double* I_buf = lg2corr_lat + IC_CAP_LAT;

    // initialize_grids: inverse-compton.h:333-369
    const double step = 2 * IC_Q;  // nu_mult = gamma_mult = ic_mult = 2
    const double lg2_nu_lo = fast_log2(nu_min), lg2_nu_hi = fast_log2(nu_max);
    const double lg2_g_lo = fast_log2(gamma_min), lg2_g_hi = fast_log2(gamma_max);
    long long nn = (long long)ceil((lg2_nu_hi - lg2_nu_lo) / step) + 1;
    const int n_seed = (int)(nn < 2 ? 2 : nn);
    nn = (long long)ceil((lg2_g_hi - lg2_g_lo) / step) + 1;
    const int n_gam = (int)(nn < 2 ? 2 : nn);
    if (n_seed > IC_CAP_SEED || n_gam > IC_CAP_GAMMA) return VAG_ST_CAPACITY;
    const double phase = lg2_nu_lo + 2 * lg2_g_lo + fast_log2(4 * IC_x0);
    const long long n_lo = (long long)floor((fast_log2(nu_IC_min) - phase) / (IC_Q * 2));
    const long long n_hi = (long long)ceil((fast_log2(nu_IC_max) - phase) / (IC_Q * 2));
    const long long n_ic_ll = (n_hi - n_lo > 1 ? n_hi - n_lo : 1) + 1;
    if (n_ic_ll > IC_CAP_OUT) return VAG_ST_CAPACITY;
    const int n_ic = (int)n_ic_ll;
    const long long ic_idx0 = n_lo * 2;
    hdr.phase = phase;
    hdr.idx0 = (int)ic_idx0;

    par.for_each(n_seed, [&](int j) {
        lg2_nu_seed[j] = lg2_nu_lo + step * (double)j;
        nu_seed[j] = fast_exp2(lg2_nu_seed[j]);
    });
    par.for_each(n_gam, [&](int i) { gamma[i] = fast_exp2(lg2_g_lo + step * (double)i); });
    // sample_distributions (inverse-compton.h:371-395) + seed tables (compute_IC_spectrum :546-556)
    par.for_each(n_gam, [&](int i) {
        const double gi = gamma[i];
        const double dg = 0.5 * ((i + 1 < n_gam ? gamma[i + 1] : gamma[i]) - (i > 0 ? gamma[i - 1] : gamma[i]));
        dN_e[i] = electron_column_den(c, gi) / (gi * gi) * dg;
    });
    par.for_each(n_seed, [&](int j) {
        const double I_seed = fast_exp2(seed_log2_I(lg2_nu_seed[j]));
        const double f = I_seed / (nu_seed[j] * nu_seed[j]);
        fv_th[j] = f;
        lg2fv_th[j] = (f > 0) ? fast_log2(f) : -kInf;
        if (j + 1 < n_seed) {
            dnu_seed[j] = nu_seed[j + 1] - nu_seed[j];
            lg2r[j] = lg2_nu_seed[j + 1] - lg2_nu_seed[j];
            inv_lg2r[j] = (lg2r[j] != 0) ? 1 / lg2r[j] : 0;
        }
    });
    par.for_each(n_ic, [&](int k) { I_buf[k] = 0; });
    const int nu_last = n_seed - 1;

    // build_cdf_thomson: inverse-compton.h:421-436 (into cdf_th/ratio_th for KN, cdf_buf/ratio_buf otherwise)
    double* cdf_t = KN ? cdf_th : cdf_buf;
    double* ratio_t = KN ? ratio_th : ratio_buf;
    par.for_each(nu_last, [&](int j) {
        const double trap = 0.5 * (fv_th[j] + fv_th[j + 1]) * dnu_seed[j];
        const double exact = power_law_bin_integral(fv_th[j], fv_th[j + 1], nu_seed[j], nu_seed[j + 1], lg2fv_th[j],
                                                    lg2fv_th[j + 1], lg2r[j], inv_lg2r[j], trap);
        fv_buf[j] = exact;  // staged; the running sum below keeps the reference's summation order
        ratio_t[j] = (trap > 0) ? exact / trap : 1;
    });
    par.for_each(1, [&](int) {
        double run = 0;  // cdf_t[j] = cdf_t[j + 1] + fv_buf[j], addends fetched eight ahead
        cdf_t[nu_last] = 0;
        int j = nu_last - 1;
        for (; j - 7 >= 0; j -= 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fv_buf[j - q];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                run = run + v[q];
                cdf_t[j - q] = run;
            }
        }
        for (; j >= 0; --j) {
            run = run + fv_buf[j];
            cdf_t[j] = run;
        }
    });

    // accumulate_IC (inverse-compton.h:486-527): lane <-> output node k; the gamma loop stays outside
    const double expq1 = fast_exp2(IC_Q * 1.0);
    const long long ns_top = (long long)(n_seed - 1) * 2;
    auto accumulate = [&](int i, const double* fv, const double* cdf, const double* ratio) {
        if (cdf[0] <= 0) return;
        const double dNb = dN_e[i];
        const long long n_off = ic_idx0 - 4 * (long long)i;
        par.for_each(n_ic, [&](int k) {
            const long long nq = n_off + 2 * (long long)k;
            if (nq < 0) {
                I_buf[k] += dNb * cdf[0];
            } else if (nq < ns_top) {
                const int j = (int)(nq / 2);
                const int fi = (int)(nq % 2);
                const double nu_lo = nu_seed[j];
                const double dnu = dnu_seed[j];
                const double f_lo = fv[j], f_hi = fv[j + 1];
                const double frac = (nu_lo * (fi ? expq1 : 1.0) - nu_lo) / dnu;
                const double rem = 1.0 - frac;
                const double f_seed = f_lo * rem + f_hi * frac;
                I_buf[k] += dNb * (cdf[j + 1] + 0.5 * (f_seed + f_hi) * rem * dnu * ratio[j]);
            }
        });
    };

    if (KN) {
        const int n_lat = 2 * (n_gam - 1) + 2 * (n_seed - 1) + 1;
        const double lg2_base = fast_log2(gamma[0]) + lg2_nu_seed[0];
        par.for_each(n_lat, [&](int k) {
            compton_correction_pair(lut, fast_exp2(lg2_base + IC_Q * (double)k), corr_lat[k], lg2corr_lat[k]);
        });
        for (int i = 0; i < n_gam; ++i) {
            if (dN_e[i] <= 0) continue;
            // build_cdf_KN: inverse-compton.h:438-484
            const double gamma_i = gamma[i];
            const int ig = 2 * i;
            const double nu_split = 1e-4 * (con::me * con::c2 / con::h) / gamma_i;
            int j_split = 0;
            while (j_split < nu_last && nu_seed[j_split] < nu_split) ++j_split;
            par.for_each(n_seed - j_split, [&](int jj) {
                const int j = j_split + jj;
                fv_buf[j] = fv_th[j] * corr_lat[ig + 2 * j];
            });
            par.for_each(nu_last - j_split, [&](int jj) {
                const int j = j_split + jj;
                const double lg2f_lo = lg2fv_th[j] + lg2corr_lat[ig + 2 * j];
                const double lg2f_hi = lg2fv_th[j + 1] + lg2corr_lat[ig + 2 * (j + 1)];
                const double trap = 0.5 * (fv_buf[j] + fv_buf[j + 1]) * dnu_seed[j];
                const double exact = power_law_bin_integral(fv_buf[j], fv_buf[j + 1], nu_seed[j], nu_seed[j + 1], lg2f_lo,
                                                            lg2f_hi, lg2r[j], inv_lg2r[j], trap);
                cdf_buf[j] = exact;  // staged
                ratio_buf[j] = (trap > 0) ? exact / trap : 1;
            });
            par.for_each(1, [&](int) {
                // the reference's running sum, in its order; eight addends are fetched ahead of the dependent chain of
                // additions (one lane, scratch in global memory: the load latency used to sit inside the chain)
                double run = 0;
                cdf_buf[nu_last] = 0;
                int j = nu_last - 1;
                for (; j - 7 >= j_split; j -= 8) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = cdf_buf[j - q];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        run = run + v[q];
                        cdf_buf[j - q] = run;
                    }
                }
                for (; j >= j_split; --j) {
                    run = run + cdf_buf[j];
                    cdf_buf[j] = run;
                }
            });
            if (j_split > 0) {
                const double delta = cdf_buf[j_split] - cdf_th[j_split];
                par.for_each(j_split, [&](int j) {
                    fv_buf[j] = fv_th[j];
                    ratio_buf[j] = ratio_th[j];
                    cdf_buf[j] = cdf_th[j] + delta;
                });
            }
            accumulate(i, fv_buf, cdf_buf, ratio_buf);
        }
    } else {
        for (int i = 0; i < n_gam; ++i) {
            if (dN_e[i] <= 0) continue;
            accumulate(i, fv_th, cdf_buf, ratio_buf);
        }
    }
    const double log2_scale = fast_log2(0.25 * con::sigmaT);
    par.for_each(n_ic, [&](int k) {
        const double lg2nu = phase + IC_Q * (double)(ic_idx0 + 2 * (long long)k);
        tab[k] = fast_log2(I_buf[k]) + lg2nu + log2_scale;
    });
    hdr.n = n_ic;
    return 0;
}

// ICPhoton::compute_log2_I_nu on a generated table (inverse-compton.h:614-647).  `breach` is set
// when the query falls outside the clamped band but inside the theoretical range (the reference
// would rebuild the cell's full-range spectrum there; the band is derived from the same
// frequency/Doppler set, so this does not happen -- it is reported, not silently extrapolated).
VAG_HD double ic_table_log2_I_nu(const IcTable& h, const double* tab, double log2_nu, bool& breach) {
    const int n = h.n;
    auto node = [&](int k) { return h.phase + IC_Q * (double)((long long)h.idx0 + 2 * (long long)k); };
    if (n >= 2 && ((log2_nu > node(n - 1) && log2_nu < h.lg2_theory_max) ||
                   (log2_nu < node(0) && log2_nu > h.lg2_theory_min)))
        breach = true;
    if (n < 2 || log2_nu > node(n - 1)) return -kInf;
    // idx = last index with node(idx) <= log2_nu, clamped to [0, n-2]
    int idx = (int)floor((log2_nu - node(0)) / (2 * IC_Q));
    if (idx < 0) idx = 0;
    if (idx > n - 2) idx = n - 2;
    while (idx + 2 < n && node(idx + 1) <= log2_nu) ++idx;
    while (idx > 0 && node(idx) > log2_nu) --idx;
    const double dl = node(idx + 1) - node(idx);
    const double slope = (dl != 0) ? (tab[idx + 1] - tab[idx]) / dl : 0;
    return tab[idx] + (log2_nu - node(idx)) * slope;
}

}  // namespace vag
