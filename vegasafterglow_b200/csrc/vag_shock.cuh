// vag_shock.cuh -- K1: blast-wave dynamics of one (phi,theta) row.
//
// Restates, for non-spreading typed jets without energy/mass injection:
//   jump conditions / helpers   src/dynamics/shock-physics.h:40-371, src/dynamics/shock.cpp:90-138
//   forward shock ODE           src/dynamics/forward-shock.tpp:27-208
//   forward+reverse shock ODE   src/dynamics/reverse-shock.tpp:17-591
// State vectors drop the constant theta component (d theta/dt = 0 without spreading: its error
// term is identically zero, so the step controller is unaffected).
#pragma once

#include "vag_dopri5.cuh"
#include "vag_grid.cuh"
#include "vag_math.cuh"
#include "vag_model.cuh"

namespace vag {

// Per-row view of the SoA shock table: 6 arrays of n_t doubles (theta is the row's constant).
struct ShockRow {
    double* t_comv;
    double* r;
    double* Gamma;
    double* Gamma_th;
    double* B;
    double* N_p;
};

// ---------------------------------------------------------------------------------------------
// shock-physics.h helpers
// ---------------------------------------------------------------------------------------------
// compute_downstr_4vel: src/dynamics/shock.cpp:90-138
VAG_HD double compute_downstr_4vel(double gamma_rel, double sigma) {
    const double ad_idx = adiabatic_idx(gamma_rel);
    const double gamma_m_1 = gamma_rel - 1;
    const double ad_idx_m_2 = ad_idx - 2;
    const double ad_idx_m_1 = ad_idx - 1;
    if (sigma <= con::sigma_cut) {
        return sqrt(vmax(gamma_m_1 * ad_idx_m_1 * ad_idx_m_1 / (-ad_idx * ad_idx_m_2 * gamma_m_1 + 2), 0.0));
    }
    const double gamma_sq = gamma_rel * gamma_rel;
    const double gamma_p_1 = gamma_rel + 1;
    const double term1 = -ad_idx * ad_idx_m_2;
    const double term2 = gamma_sq - 1;
    const double A = term1 * gamma_m_1 + 2;
    const double B = -gamma_p_1 * (-ad_idx_m_2 * (ad_idx * gamma_sq + 1) + ad_idx * ad_idx_m_1 * gamma_rel) * sigma -
                     gamma_m_1 * (term1 * (gamma_sq - 2) + 2 * gamma_rel + 3);
    const double C = gamma_p_1 * (ad_idx * (1 - ad_idx / 4) * term2 + 1) * sigma * sigma +
                     term2 * (2 * gamma_rel + ad_idx_m_2 * (ad_idx * gamma_rel - 1)) * sigma +
                     gamma_p_1 * gamma_m_1 * gamma_m_1 * ad_idx_m_1 * ad_idx_m_1;
    const double D = -gamma_m_1 * gamma_p_1 * gamma_p_1 * ad_idx_m_2 * ad_idx_m_2 * sigma * sigma / 4;
    const double b = B / A;
    const double c = C / A;
    const double d = D / A;
    const double P = c - b * b / 3;
    const double Q = 2 * b * b * b / 27 - b * c / 3 + d;
    const double u = sqrt(vmax(-P, 0.0) / 3);
    const double denom = 2 * P * u;
    const double v = (denom != 0) ? vclamp(3 * Q / denom, -1.0, 1.0) : 0.0;
    const double x_max = 2 * u * cos(acos(v) / 3) - b / 3;
    if (x_max <= 0) return 0;
    const double prod = -d / x_max;
    const double sum = (c - prod) / x_max;
    const double uds = (sum + sqrt(vmax(sum * sum - 4 * prod, 0.0))) / 2;
    return sqrt(vmax(uds, 0.0));
}

// sigma = 0 branch of compute_downstr_4vel with the adiabatic index supplied by the caller
VAG_HD double compute_downstr_4vel_cold(double gamma_rel, double ad_idx) {
    const double gamma_m_1 = gamma_rel - 1;
    const double ad_idx_m_2 = ad_idx - 2;
    const double ad_idx_m_1 = ad_idx - 1;
    // ad_idx in (4/3, 5/3]: the denominator is >= 2
    return vsqrt(vmax(vdiv(gamma_m_1 * ad_idx_m_1 * ad_idx_m_1, -ad_idx * ad_idx_m_2 * gamma_m_1 + 2), 0.0));
}

// shock-physics.h:40-66
VAG_HD double compute_upstr_4vel(double u_down, double gamma_rel) {
    return sqrt((1 + u_down * u_down) * vmax((gamma_rel - 1) * (gamma_rel + 1), 0.0)) + u_down * gamma_rel;
}
VAG_HD double compute_4vel_jump(double gamma_rel, double sigma_upstr) {
    const double u_down_s = compute_downstr_4vel(gamma_rel, sigma_upstr);
    const double u_up_s = compute_upstr_4vel(u_down_s, gamma_rel);
    double ratio_u = u_up_s / u_down_s;
    if (u_down_s == 0.) ratio_u = 4 * gamma_rel;
    return ratio_u;
}
// shock-physics.h:75-78
VAG_HD double compute_sound_speed_ad(double Gamma_rel, double ad_idx) {
    return vsqrt(vmax(vdiv(ad_idx * (ad_idx - 1) * (Gamma_rel - 1), 1 + (Gamma_rel - 1) * ad_idx), 0.0)) * con::c;
}
VAG_HD double compute_sound_speed(double Gamma_rel) { return compute_sound_speed_ad(Gamma_rel, adiabatic_idx(Gamma_rel)); }
// compute_4vel_jump with the adiabatic index of gamma_rel supplied by the caller
VAG_HD double compute_downstr_4vel(double gamma_rel, double sigma);
VAG_HD double compute_4vel_jump_ad(double gamma_rel, double sigma_upstr, double ad_idx) {
    const double u_down_s =
        (sigma_upstr <= con::sigma_cut) ? compute_downstr_4vel_cold(gamma_rel, ad_idx) : compute_downstr_4vel(gamma_rel, sigma_upstr);
    const double u_up_s = vsqrt((1 + u_down_s * u_down_s) * vmax((gamma_rel - 1) * (gamma_rel + 1), 0.0)) + u_down_s * gamma_rel;
    // u_down_s == 0 is overridden below (the quotient is discarded), otherwise it is positive
    double ratio_u = vdiv(u_up_s, u_down_s);
    if (u_down_s == 0.) ratio_u = 4 * gamma_rel;
    return ratio_u;
}
// shock-physics.h:88-102
VAG_HD double compute_effective_Gamma(double adx, double Gamma) { return vdiv(adx * Gamma * Gamma - adx + 1, Gamma); }
VAG_HD double compute_effective_Gamma_dGamma(double adx, double Gamma) {
    const double Gamma2 = Gamma * Gamma;
    return vdiv(adx * Gamma2 + adx - 1, Gamma2);
}
// shock-physics.h:179-181
VAG_HD double compute_upstr_B(double rho_up, double sigma) { return sqrt((4 * con::pi * con::c2) * sigma * rho_up); }
// shock-physics.h:192-204
VAG_HD double compute_rel_Gamma(double gamma1, double gamma2) {
    const double u1u2 = vsqrt(vmax((gamma1 - 1) * (gamma1 + 1) * (gamma2 - 1) * (gamma2 + 1), 0.0));
    const double d = gamma1 - gamma2;
    const double denom = gamma1 * gamma2 - 1 + u1u2;
    if (denom <= 0) return 1;
    return 1 + vdiv(d * d, denom);
}
// shock-physics.h:223-229
VAG_HD double compute_adiabatic_cooling_rate2(double ad_idx, double r, double x, double u, double drdt, double dxdt) {
    double dlnvdt = 2 * drdt / r;
    if (x > 0) dlnvdt += dxdt / x;
    return -(ad_idx - 1) * dlnvdt * u;
}
// shock-physics.h:247-288
VAG_HD double radiative_efficiency(const RadCfg& rad, double t_comv, double Gamma_th, double e_th) {
    if (rad.eps_e_rad == 0) return 0;
    const double gamma_m = rad.gamma_m_coeff * (Gamma_th - 1) + 1;
    const double den = e_th * t_comv;
    const bool ok = den > 1e-100 && den < 1e100;  // the ordinary case: branch-free arithmetic
    const double den_s = ok ? den : 1.0;
    const double gamma_bar_f = vdiv(rad.gamma_c_coeff, den_s);
    const double gamma_c_f = 0.5 * (gamma_bar_f + vsqrt(gamma_bar_f * gamma_bar_f + 4));
    double ratio = vdiv(gamma_m, gamma_c_f);
    if (!ok) {  // e_th = 0 (Gamma clamped to 1), overflow ...: IEEE semantics as in the reference
        const double gamma_bar = rad.gamma_c_coeff / den;
        const double gamma_c = 0.5 * (gamma_bar + sqrt(gamma_bar * gamma_bar + 4));
        ratio = gamma_m / gamma_c;
    }
    // fast_pow(ratio, p - 2) evaluated unconditionally on a safe argument and selected afterwards, so that the
    // chain can overlap with the rest of the right-hand side instead of sitting behind a branch
    const bool slow_cooling = ratio < 1 && rad.p > 2;
    const double ratio_s = (slow_cooling && ratio > 1e-300) ? ratio : 0.5;
    const double pw = dexp2_nc(vmax((rad.p - 2) * dlog2_nc(ratio_s), -1000.0));
    if (slow_cooling) return rad.eps_e_rad * ((ratio > 1e-300) ? pw : 0.0);
    return rad.eps_e_rad;
}
// shock-physics.h:300-312
VAG_HD double compute_Gamma_therm(double U_th, double mass, bool limiter = false) {
    if (mass == 0) return 1;
    const double Gamma_th = U_th / (mass * con::c2) + 1;
    if (limiter && Gamma_th < con::gamma_therm_cut) return 1;
    return Gamma_th;
}
// shock-physics.h:352-371
VAG_HD double compute_compression(double Gamma_upstr, double Gamma_downstr, double sigma_upstr) {
    return compute_4vel_jump(compute_rel_Gamma(Gamma_upstr, Gamma_downstr), sigma_upstr);
}
VAG_HD double compute_downstr_B(double eps_B, double rho_upstr, double B_upstr, double Gamma_th, double comp_ratio) {
    const double rho_downstr = rho_upstr * comp_ratio;
    const double e_th = (Gamma_th - 1) * rho_downstr * con::c2;
    return sqrt(8 * con::pi * eps_B * e_th) + B_upstr * comp_ratio;
}

// simpson_logspace / enclosed_mass / enclosed_thermal_energy: shock-physics.h:401-450
template <class F>
VAG_HD double simpson_logspace(const F& f, double r) {
    constexpr int N = 32;
    const double u_max = log(r);
    const double u_min = u_max - 18;
    const double h = (u_max - u_min) / N;
    double sum = f(u_min) + f(u_max);
    for (int i = 1; i < N; i += 2) sum += 4 * f(u_min + i * h);
    for (int i = 2; i < N; i += 2) sum += 2 * f(u_min + i * h);
    return sum * h / 3;
}
VAG_HD double enclosed_mass_numeric(const ModelCfg& m, double r) {
    return simpson_logspace(
        [&](double u) {
            const double ri = exp(u);
            return medium_rho(m, ri) * ri * ri * ri;
        },
        r);
}
VAG_HD double enclosed_thermal_energy_numeric(const ModelCfg& m, double r, double Gamma, double ad_idx, double eps_e) {
    const double cooling_exp = 3 * (ad_idx - 1);
    return (1 - eps_e) * (Gamma - 1) * con::c2 *
           simpson_logspace(
               [&](double u) {
                   const double ri = exp(u);
                   return medium_rho(m, ri) * ri * ri * ri * pow(ri / r, cooling_exp);
               },
               r);
}
// enclosed_thermal_energy_medium: shock-physics.h:452-469 (ISM closed form, otherwise Simpson)
VAG_HD double enclosed_thermal_energy_medium(const ModelCfg& m, double r, double Gamma, double ad_idx, double eps_e) {
    if (m.medium_type == VAG_MEDIUM_ISM) {
        const double rho = m.rho_ism;
        const double cooling_exp = 3 * (ad_idx - 1);
        const double pow_exp = 3 + cooling_exp;
        const double x0 = exp(-18.0);
        const double attenuation = 1 - pow(x0, pow_exp);
        const double integral = rho * r * r * r * attenuation / pow_exp;
        return (1 - eps_e) * (Gamma - 1) * con::c2 * integral;
    }
    return enclosed_thermal_energy_numeric(m, r, Gamma, ad_idx, eps_e);
}

// Default table values of an untouched row (Shock ctor, src/dynamics/shock.cpp:12-25).
VAG_HD void fill_default_row(const ShockRow& s, int k0, int n_t) {
    for (int k = k0; k < n_t; ++k) {
        s.t_comv[k] = 0;
        s.r[k] = 0;
        s.Gamma[k] = 1;
        s.Gamma_th[k] = 1;
        s.B[k] = 0;
        s.N_p[k] = 0;
    }
}
// set_stopping_shock: shock-physics.h:388-397
VAG_HD void set_stopping_row(const ShockRow& s, int n_t, double t_comv, double r) {
    for (int k = 0; k < n_t; ++k) {
        s.t_comv[k] = t_comv;
        s.r[k] = r;
        s.Gamma[k] = 1;
        s.Gamma_th[k] = 1;
        s.B[k] = 0;
        s.N_p[k] = 0;
    }
}

// save_fwd_shock_state: forward-shock.tpp:151-173 (+ write_shock_state shock-physics.h:329-338)
VAG_HD void save_fwd_state(const ModelCfg& m, double eps_B, const ShockRow& s, int k, double Gamma, double m2,
                           double U2_th, double r, double t_comv) {
    const double comp_ratio = compute_compression(1, Gamma, 0);
    const double rho = medium_rho(m, r);
    const double Gamma_th = compute_Gamma_therm(U2_th, m2);
    const double B = compute_downstr_B(eps_B, rho, 0, Gamma_th, comp_ratio);
    s.t_comv[k] = t_comv;
    s.r[k] = r;
    s.Gamma[k] = Gamma;
    s.Gamma_th[k] = Gamma_th;
    s.B[k] = B;
    s.N_p[k] = m2 / con::mp;
}

// ---------------------------------------------------------------------------------------------
// forward shock: state = [Gamma, m2, U2_th, r, t_comv]
// ---------------------------------------------------------------------------------------------
// INJ: the jet injects energy (magnetar): the reference's ForwardState then carries eps_jet, whose
// derivative deps_dt(t) enters the step-size control (forward-shock.hpp:23-26, forward-shock.tpp:46-48,89-91).
// SPREAD: lateral spreading, theta becomes a state variable (forward-shock.tpp:36-40,57-60,78-84,110-115).
template <bool INJ, bool SPREAD>
struct FwdEqnT {
    const ModelCfg& m;
    double m_jet0, theta0, dOmega0, theta_s;
    int mode;  // 0: rhs_general; 1 / 2: rhs_fast without / with radiative losses (ISM or Wind(k = 2), no injection, no spreading)
    enum { iG = 0, iM2 = 1, iU = 2, iR = 3, iT = 4, iTh = 5, iE = SPREAD ? 6 : 5, N = 5 + (INJ ? 1 : 0) + (SPREAD ? 1 : 0) };

    VAG_HD FwdEqnT(const ModelCfg& m_, double theta, double theta_s_) : m(m_), theta0(theta), theta_s(theta_s_) {
        m_jet0 = jet_eps_k(m, theta) / jet_Gamma0(m, theta) / con::c2;  // forward-shock.tpp:20
        m_jet0 /= 1 + m.sigma0;                                          // :21-23
        dOmega0 = 1 - cos(theta);                                        // :18
        mode = (INJ || SPREAD || m.wind_generic) ? 0 : (m.fwd.eps_e_rad != 0 ? 2 : 1);
    }

    // ForwardShockEqn::operator(): forward-shock.tpp:27-118.  As for the shock pair (FREqn), rows without injection
    // and spreading evaluate the branch-free rhs_fast -- one basic block per stage, so that the independent chains of
    // the right-hand side overlap with one warp per scheduler -- and repeat the evaluation with rhs_general whenever
    // an operand leaves the range that arithmetic is valid for.
    VAG_HD void operator()(const double* x, double* d, double t) const {
        if (!INJ && !SPREAD) {
            if (mode == 2) {
                if (rhs_fast<true>(x, d)) return;
            } else if (mode == 1) {
                if (rhs_fast<false>(x, d)) return;
            }
        }
        rhs_general(x, d, t);
    }

    template <bool RAD>
    VAG_HD bool rhs_fast(const double* x, double* d) const {
        const double Gamma = x[iG], r = x[iR];
        const double u2 = (Gamma - 1) * (Gamma + 1);
        bool ok = u2 >= 0 && r > 1e-150 && r < 1e150;  // Gamma < 1 must give NaN as in the reference: general path
        const double u = vsqrt(u2 >= 0 ? u2 : 1.0);
        const double dr = u * (Gamma + u) * con::c;
        d[iR] = dr;
        d[iT] = Gamma + u;
        const double rho_wind = vdiv(m.wind_A, m.wind_r02 + r * r) + m.rho_ism;  // medium.h:107-109
        const double rho = (m.medium_type == VAG_MEDIUM_ISM) ? m.rho_ism : rho_wind;
        const double dm2 = r * r * rho * dr;
        d[iM2] = dm2;
        double eps_rad = 0;
        if (RAD) {  // radiative_efficiency (shock-physics.h:247-288), ordinary operand range only
            const RadCfg& rad = m.fwd;
            const double e_th = (Gamma - 1) * 4 * Gamma * rho * con::c2;
            const double gamma_m = rad.gamma_m_coeff * (Gamma - 1) + 1;
            const double den = e_th * x[iT];
            const bool den_ok = den > 1e-100 && den < 1e100;
            ok = ok && den_ok;
            const double gamma_bar = vdiv(rad.gamma_c_coeff, den_ok ? den : 1.0);
            const double gamma_c = 0.5 * (gamma_bar + vsqrt(gamma_bar * gamma_bar + 4));
            const double ratio = vdiv(gamma_m, gamma_c);
            const bool slow_cooling = ratio < 1 && rad.p > 2;
            const bool ratio_ok = ratio > 1e-300;
            double pw = dexp2_nc(vmax((rad.p - 2) * dlog2_nc((slow_cooling && ratio_ok) ? ratio : 0.5), -1000.0));
            VAG_KEEP(pw);
            eps_rad = slow_cooling ? rad.eps_e_rad * (ratio_ok ? pw : 0.0) : rad.eps_e_rad;
        }
        const double ad_idx = adiabatic_idx_fast(Gamma);
        const double dlnV_r = vdiv(3 * dr, r);
        double dG;
        {
            const double Gamma2 = Gamma * Gamma;
            const double Gamma_eff = vdiv(ad_idx * (Gamma2 - 1) + 1, Gamma);
            const double dGamma_eff = vdiv(ad_idx * (Gamma2 + 1) - 1, Gamma2);
            const double U = x[iU];
            const double a1 = -(Gamma - 1) * (Gamma_eff + 1) * con::c2 * dm2;
            const double a2 = (ad_idx - 1) * Gamma_eff * U * dlnV_r;
            const double b1 = (m_jet0 + x[iM2]) * con::c2;
            const double b2 = (dGamma_eff + vdiv(Gamma_eff * (ad_idx - 1), Gamma)) * U;
            const double b = b1 + b2;
            const bool b_ok = b > 1e-290 && b < 1e290;
            dG = vdiv(a1 + a2, b_ok ? b : 1.0);
            ok = ok && b_ok && fabs(dG) < kInf;
            d[iG] = dG;
        }
        {
            const double dlnVdt = dlnV_r - vdiv(dG, Gamma);
            d[iU] = (1 - eps_rad) * (Gamma - 1) * con::c2 * dm2 - (ad_idx - 1) * dlnVdt * x[iU];
        }
        return ok;
    }

    VAG_HD void rhs_general(const double* x, double* d, double t) const {
        const double Gamma = x[iG];
        const double deps = INJ ? jet_deps_dt(m, theta0, t) : 0.0;
        if (INJ) d[INJ ? iE : 0] = deps;
        const double u2 = (Gamma - 1) * (Gamma + 1);
        const double u = sqrt(u2);  // IEEE: a stage value of Gamma below 1 must give NaN as in the reference
        d[iR] = u * (Gamma + u) * con::c;
        d[iT] = Gamma + u;
        double dth = 0, sin_theta = 0, cos_theta = 1;
        if (SPREAD) {
            const double th = x[SPREAD ? iTh : 0];
            if (th < 0.5 * con::pi) {  // compute_dtheta_dt: shock-physics.h:154-158
                constexpr double Q = 7;
                const double f = 1 / (1 + u * theta_s * Q);
                dth = d[iR] / (2 * Gamma * x[iR]) * sqrt((2 * u2 + 3) / (4 * u2 + 3)) * f;
            }
            d[SPREAD ? iTh : 0] = dth;
            sin_theta = sin(th);
            cos_theta = cos(th);
        }
        const double rho = medium_rho(m, x[iR]);
        d[iM2] = x[iR] * x[iR] * rho * d[iR];
        const double e_th = (Gamma - 1) * 4 * Gamma * rho * con::c2;
        const double eps_rad = radiative_efficiency(m.fwd, x[iT], Gamma, e_th);
        const double ad_idx = adiabatic_idx_fast(Gamma);
        const double dlnV_r = vdiv(3 * d[iR], x[iR]);  // 3 / r * dr/dt, r > 0
        // lateral-expansion term sin(theta) / (1 - cos(theta)) * dtheta/dt of both rate equations
        const double lat = SPREAD ? sin_theta / (1 - cos_theta) * dth : 0.0;
        // compute_dGamma_dt
        {
            const double Gamma2 = Gamma * Gamma;
            const double Gamma_eff = vdiv(ad_idx * (Gamma2 - 1) + 1, Gamma);
            const double dGamma_eff = vdiv(ad_idx * (Gamma2 + 1) - 1, Gamma2);
            double dlnVdt = dlnV_r;
            double U = x[iU];
            double dm_dt_swept = d[iM2];
            double m_swept = x[iM2];
            if (SPREAD) {
                const double f_spread = (1 - cos_theta) / dOmega0;
                dm_dt_swept = dm_dt_swept * f_spread + m_swept / dOmega0 * sin_theta * dth;
                m_swept *= f_spread;
                dlnVdt += lat;
                U *= f_spread;
            }
            double a1 = -(Gamma - 1) * (Gamma_eff + 1) * con::c2 * dm_dt_swept;
            if (INJ) a1 += deps;
            const double a2 = (ad_idx - 1) * Gamma_eff * U * dlnVdt;
            const double b1 = (m_jet0 + m_swept) * con::c2;
            const double b2 = (dGamma_eff + vdiv(Gamma_eff * (ad_idx - 1), Gamma)) * U;
            d[iG] = (a1 + a2) / (b1 + b2);
        }
        // compute_dU_dt
        {
            double dm_dt_swept = d[iM2];
            double dlnVdt = dlnV_r - vdiv(d[iG], Gamma);
            if (SPREAD) {
                dm_dt_swept = dm_dt_swept + x[iM2] * lat;
                dlnVdt += lat;
                dlnVdt += lat / (ad_idx - 1);
            }
            d[iU] = (1 - eps_rad) * (Gamma - 1) * con::c2 * dm_dt_swept - (ad_idx - 1) * dlnVdt * x[iU];
        }
    }

    // set_init_state: forward-shock.tpp:120-149
    VAG_HD void set_init_state(double* x, double theta, double t0) const {
        const double Gamma4 = jet_Gamma0(m, theta);
        const double beta4 = gamma_to_beta(Gamma4);
        x[iR] = beta4 * con::c * t0 * Gamma4 * Gamma4 * (1 + beta4);
        x[iT] = x[iR] / sqrt((Gamma4 - 1) * (Gamma4 + 1)) / con::c;
        // enclosed_mass_medium (shock-physics.h:442-449): analytic for ISM / Wind, Simpson for a generic Medium
        x[iM2] = m.wind_generic ? enclosed_mass_numeric(m, x[iR]) : medium_mass(m, x[iR]);
        x[iG] = Gamma4;
        if (SPREAD) x[SPREAD ? iTh : 0] = theta;
        if (INJ) x[INJ ? iE : 0] = jet_eps_k(m, theta);
        const double ad_idx = adiabatic_idx(Gamma4);
        x[iU] = enclosed_thermal_energy_medium(m, x[iR], Gamma4, ad_idx, m.fwd.radiative ? m.fwd.eps_e : 0.0);
    }
};
using FwdEqn = FwdEqnT<false, false>;

// Raw dense-output samples of one row.  The ODE kernel stores only the interpolated state vector
// of every lattice node (component c -> plane[c][k]); the derived shock quantities are computed
// afterwards, one thread per cell, by finish_*_cell (same arithmetic as save_*_shock_state, moved
// out of the latency-critical sequential loop).  The planes are the shock-table planes themselves
// (forward table = components 0..5, reverse table = components 6..10), finished in place.
struct RawRow {
    double* c[11];
    double* theta;  // Shock::theta of a spreading row (nullptr otherwise)
};
// per-row record left by the ODE kernel for the finishing pass
struct RowDyn {
    int n_saved;        // nodes [0, n_saved) hold raw states; -1: the tables already hold final values
    int injection_idx;
    double V3_comv_x, rho3_x, B3_ordered_x;  // crossing state (reverse-shock.hpp:66-70)
};

// grid_solve_fwd_shock: forward-shock.tpp:175-208.  Returns VAG_ST_* bits.
// The accepted-step loop of integrate_adaptive/dense output is flattened to one dopri5 ATTEMPT per
// iteration (identical sequence of attempts, step sizes and accepted states): in a warp of 32 rows a
// rejected attempt of one row then costs the other rows nothing.
template <bool INJ, bool SPREAD>
VAG_HD int solve_fwd_row(const ModelCfg& m, double theta, double theta_s, double t_dec, const double* t, int n_t,
                         const ShockRow& s, const RawRow& raw, RowDyn& rd) {
    using Eqn = FwdEqnT<INJ, SPREAD>;
    Eqn eqn(m, theta, theta_s);
    double x[Eqn::N];
    const double t0 = vmin(t[0], vmin(0.1 * unit::sec, 0.1 * t_dec));
    eqn.set_init_state(x, theta, t0);
    rd.injection_idx = n_t;
    rd.V3_comv_x = rd.rho3_x = rd.B3_ordered_x = 0;
    if (x[Eqn::iG] <= con::Gamma_cut) {
        set_stopping_row(s, n_t, x[Eqn::iT], x[Eqn::iR]);
        if (SPREAD)
            for (int k = 0; k < n_t; ++k) raw.theta[k] = theta;  // set_stopping_shock: shock-physics.h:392
        rd.n_saved = -1;
        return 0;
    }
    // five to seven state variables: the register-resident stepper fits without spills (the 11-variable
    // pair system uses the shared-memory one)
    Dopri5<Eqn::N> st;
    st.initialize(x, t0, 0.01 * t0, m.rtol);
    const double t_back = t[n_t - 1];
    int k = 0, status = 0, fails = 0, steps = 0;
    const int max_fails = m.max_ode_fails, max_steps = m.max_ode_steps;  // in registers: the loop stores to global memory
    // the next two lattice nodes ride in registers (+inf behind the last one): the loop test never waits for a load
    double t_next = t[0], t_next2 = 1 < n_t ? t[1] : kInf;
    st.begin_step(eqn);
    while (st.t <= t_back) {
        if (!st.try_step(eqn)) {
            if (++fails >= max_fails) {
                status |= VAG_ST_ODE_FAIL500;
                break;
            }
            continue;
        }
        fails = 0;
        if (++steps > max_steps) {
            status |= VAG_ST_ODE_STEP_CAP;
            break;
        }
        while (st.t > t_next) {  // t_next = t[k]
            st.calc_state(t_next, x);
#pragma unroll
            for (int c = 0; c < 5; ++c) raw.c[c][k] = x[c];
            if (SPREAD) raw.theta[k] = x[SPREAD ? Eqn::iTh : 0];
            ++k;
            t_next = t_next2;
            t_next2 = k + 1 < n_t ? t[k + 1] : kInf;
        }
        st.t_old = st.t;  // dense_output_runge_kutta::do_step: the next step starts here
    }
    if (SPREAD)
        for (int q = k; q < n_t; ++q) raw.theta[q] = 0;  // Shock ctor default (shock.cpp:15)
    rd.n_saved = k;
    return status;
}

// save_fwd_shock_state of node k from its raw state (forward-only rows), or the Shock-ctor defaults
// for nodes the integration never reached (shock.cpp:12-25)
VAG_HD void default_cell(const ShockRow& s, int k) {
    s.t_comv[k] = 0;
    s.r[k] = 0;
    s.Gamma[k] = 1;
    s.Gamma_th[k] = 1;
    s.B[k] = 0;
    s.N_p[k] = 0;
}
VAG_HD void finish_fwd_cell(const ModelCfg& m, const RowDyn& rd, const ShockRow& s, const RawRow& raw, int k) {
    if (rd.n_saved < 0) return;
    if (k >= rd.n_saved) {
        default_cell(s, k);
        return;
    }
    double x[FwdEqn::N];
#pragma unroll
    for (int c = 0; c < FwdEqn::N; ++c) x[c] = raw.c[c][k];
    save_fwd_state(m, m.fwd.eps_B, s, k, x[FwdEqn::iG], x[FwdEqn::iM2], x[FwdEqn::iU], x[FwdEqn::iR], x[FwdEqn::iT]);
}

// ---------------------------------------------------------------------------------------------
// forward + reverse shock pair: reverse-shock.hpp:29-45 without theta
// state = [Gamma, x4, x3, m2, m3, U2_th, U3_th, r, t_comv, eps4, m4]
// ---------------------------------------------------------------------------------------------
VAG_HD double smoothstep(double edge0, double edge1, double x) {  // reverse-shock.tpp:11-20
    double t = vdiv(x - edge0, edge1 - edge0);  // edges are distinct finite constants
    if (t < 0.0)
        t = 0.0;
    else if (t > 1.0)
        t = 1.0;
    return t * t * (3.0 - 2.0 * t);
}

struct FRState {
    double Gamma, x4, x3, m2, m3, U2_th, U3_th, r, t_comv, eps4, m4;
};

struct FREqn {
    const ModelCfg& m;
    double Gamma4, deps0_dt, dm0_dt, u4, theta0;
    double cs4, beta4;  // compute_sound_speed(Gamma4), gamma_to_beta(Gamma4): row constants of the RHS
    int mode;           // right-hand side of this row: 0 rhs_general, 1 / 2 rhs_fast without / with radiative losses
    // crossing state: reverse-shock.hpp:66-70
    double u_x, r_x, B3_ordered_x, V3_comv_x, rho3_x;
    enum { iG = 0, iX4, iX3, iM2, iM3, iU2, iU3, iR, iT, iE4, iM4, N };

    // FRShockEqn ctor: reverse-shock.tpp:22-40
    VAG_HD FREqn(const ModelCfg& m_, double theta) : m(m_), theta0(theta) {
        Gamma4 = jet_Gamma0(m, theta);
        deps0_dt = jet_eps_k(m, theta) / m.T0;
        dm0_dt = deps0_dt / (Gamma4 * con::c2);
        dm0_dt /= 1 + m.sigma0;  // reverse-shock.tpp:36-38
        u4 = sqrt(Gamma4 * Gamma4 - 1) * con::c;
        cs4 = compute_sound_speed(Gamma4);
        beta4 = gamma_to_beta(Gamma4);
        u_x = r_x = B3_ordered_x = V3_comv_x = rho3_x = 0;
        // cold ejecta without energy injection in an ISM / Wind(k = 2) medium: the branch-free right-hand side
        mode = (m.has_magnetar || m.wind_generic || m.sigma0 > 0) ? 0 : (m.fwd.eps_e_rad != 0 ? 2 : 1);
    }

    VAG_HD double injection_efficiency(double dm4) const {  // reverse-shock.tpp:42-47
        if (dm0_dt > 0 && dm4 > 0) return vmin(vdiv(dm4, dm0_dt), 1.0);
        return 0.0;
    }
    VAG_HD double shell_sigma(double eps4, double m4) const {  // reverse-shock.tpp:359-363
        const double sigma = eps4 / (Gamma4 * m4 * con::c2) - 1;
        return (sigma > con::sigma_cut) ? sigma : 0;
    }
    VAG_HD bool crossing_complete_m(double m3, double m4, double t) const {  // reverse-shock.tpp:49-60
        if (m3 < 0.999 * m4) return false;
        if (smoothstep(m.T0 * 1.5, m.T0 * 0.5, t) > 1e-6) return false;
        return true;
    }
    VAG_HD bool crossing_complete(const double* x, double t) const { return crossing_complete_m(x[iM3], x[iM4], t); }

    // FRShockEqn::operator(): reverse-shock.tpp:252-294 (+ the rate terms :62-250)
    // Rows of unmagnetised ejecta without energy injection (mode 1 / 2) evaluate rhs_fast, ONE basic block of
    // straight-line code; whenever an operand leaves the range its branch-free arithmetic is valid for (or the shell
    // turns out magnetised) it reports false and the evaluation is repeated by rhs_general, which keeps the IEEE
    // operators and the branches of the reference.
    VAG_HD void operator()(const double* xr, double* d, double t) const {
        if (mode == 2) {
            if (rhs_fast<true>(xr, d, t)) return;
        } else if (mode == 1) {
            if (rhs_fast<false>(xr, d, t)) return;
        }
        rhs_general(xr, d, t);
    }

    // The same right-hand side as rhs_general for sigma = 0, no injection, ISM / Wind(k = 2): every condition is a
    // select, every division / square root the branch-free faithful form on an operand checked to be positive, normal
    // and finite (`ok`).  Measured on the B200 (profiles/r02c_*): blocks of > 100 instructions retire at ~80 stall
    // samples per instruction, the 20-60 instruction blocks the branches of rhs_general leave at 200-400 -- with one
    // warp per scheduler only a long block lets the independent chains of this function overlap.
    template <bool RAD>
    VAG_HD bool rhs_fast(const double* xr, double* d, double t) const {
        const double Gamma = vclamp(xr[iG], 1.0, Gamma4);
        const double m4 = xr[iM4];
        const double m3 = vclamp(xr[iM3], 0.0, vmax(m4, 0.0));
        const double x3 = vmax(xr[iX3], 0.0);
        const double U3 = vmax(xr[iU3], 0.0);
        const double x4 = xr[iX4], m2 = xr[iM2], U2 = xr[iU2], r = xr[iR], t_comv = xr[iT], eps4 = xr[iE4];
        bool ok = r > 1e-150 && r < 1e150;

        const double u3 = vsqrt((Gamma - 1) * (Gamma + 1));
        const double dr = u3 * (Gamma + u3) * con::c;
        const double dtc = Gamma + u3;
        d[iR] = dr;
        d[iT] = dtc;
        // medium_rho: ISM medium.h:58, Wind(k = 2) medium.h:107-109
        const double rho_wind = vdiv(m.wind_A, m.wind_r02 + r * r) + m.rho_ism;
        const double rho = (m.medium_type == VAG_MEDIUM_ISM) ? m.rho_ism : rho_wind;
        const double dm2 = r * r * rho * dr;
        d[iM2] = dm2;

        const double inject_w = smoothstep(m.T0 * 1.5, m.T0 * 0.5, t);
        const bool inj_on = inject_w > 1e-6;
        const double deps4 = inj_on ? inject_w * deps0_dt : 0.0;
        const double dm4 = inj_on ? inject_w * dm0_dt : 0.0;
        d[iE4] = deps4;
        d[iM4] = dm4;

        // compute_rel_Gamma (shock-physics.h:192-204)
        double Gamma34;
        {
            const double u1u2 = vsqrt(vmax((Gamma4 - 1) * (Gamma4 + 1) * (Gamma - 1) * (Gamma + 1), 0.0));
            const double dg = Gamma4 - Gamma;
            const double denom = Gamma4 * Gamma - 1 + u1u2;
            const bool degenerate = denom <= 0;
            const double g = 1 + vdiv(dg * dg, degenerate ? 1.0 : denom);
            Gamma34 = degenerate ? 1.0 : g;
        }
        const double ad2 = adiabatic_idx_fast(Gamma);
        const double ad34 = adiabatic_idx_fast(Gamma34);
        const double cs34 = compute_sound_speed_ad(Gamma34, ad34);
        const double cs34_dtc = cs34 * dtc;
        const double dlnv_r = vdiv(2 * dr, r);
        // shell_sigma: the general path when the shell is (or becomes) magnetised
        {
            const double den = Gamma4 * m4 * con::c2;
            const bool den_ok = den > 1e-290 && den < 1e290;
            const double sigma = vdiv(eps4, den_ok ? den : 1.0) - 1;
            ok = ok && den_ok && !(sigma > con::sigma_cut);
        }
        const double comp_ratio = compute_4vel_jump_ad(Gamma34, 0.0, ad34);

        // injection_efficiency
        const bool inj_eff = dm0_dt > 0 && dm4 > 0;
        double f_inj = vmin(vdiv(dm4, inj_eff ? dm0_dt : 1.0), 1.0);
        VAG_KEEP(f_inj);
        const double f = inj_eff ? f_inj : 0.0;
        const double sound4 = cs4 * dtc;
        const double dx4 = (f > 1e-6) ? f * u4 + (1 - f) * sound4 : sound4;
        d[iX4] = dx4;

        double eps_rad = 0;
        if (RAD) {  // radiative_efficiency (shock-physics.h:247-288), ordinary operand range only
            const RadCfg& rad = m.fwd;
            const double e_th = (Gamma - 1) * 4 * Gamma * rho * con::c2;
            const double gamma_m = rad.gamma_m_coeff * (Gamma - 1) + 1;
            const double den = e_th * t_comv;
            const bool den_ok = den > 1e-100 && den < 1e100;
            ok = ok && den_ok;
            const double gamma_bar = vdiv(rad.gamma_c_coeff, den_ok ? den : 1.0);
            const double gamma_c = 0.5 * (gamma_bar + vsqrt(gamma_bar * gamma_bar + 4));
            const double ratio = vdiv(gamma_m, gamma_c);
            const bool slow_cooling = ratio < 1 && rad.p > 2;
            const bool ratio_ok = ratio > 1e-300;
            double pw = dexp2_nc(vmax((rad.p - 2) * dlog2_nc((slow_cooling && ratio_ok) ? ratio : 0.5), -1000.0));
            VAG_KEEP(pw);
            eps_rad = slow_cooling ? rad.eps_e_rad * (ratio_ok ? pw : 0.0) : rad.eps_e_rad;
        }

        // compute_dx3_dt
        const double remaining = vmax(m4 - m3, 0.0);
        const bool has_shell = !(m4 <= 0);
        const double m4_s = has_shell ? m4 : 1.0;
        const double crossing_w = f + vdiv((1.0 - f) * remaining, m4_s);
        const double penetration = vdiv(Gamma * comp_ratio, Gamma4) - 1;
        const bool crossing_on = has_shell && !(crossing_w < 1e-6) && !(penetration <= 0);
        double dx3;
        {
            const double pen_s = crossing_on ? penetration : 1.0;
            const double beta3 = vdiv(u3, Gamma);
            const double dx3dt = vdiv((Gamma4 - Gamma) * (Gamma4 + Gamma) * (1 + beta3) * con::c,
                                      Gamma4 * Gamma4 * (beta3 + beta4) * pen_s);
            double crossing = fabs(dx3dt * Gamma);
            // v_ms of rhs_general with sigma = 0: va2 = 0, sqrt(0 + cs2 (1 - 0)) c
            const double cs2 = cs34 * cs34 / (con::c * con::c);
            const double v_ms = vsqrt(cs2) * con::c;
            crossing = (penetration < 1) ? vmin(crossing, v_ms * dtc) : crossing;
            dx3 = crossing_on ? crossing_w * crossing + (1.0 - crossing_w) * cs34_dtc : cs34_dtc;
        }
        d[iX3] = dx3;

        // compute_dm3_dt
        double dm3;
        {
            const bool feeding = has_shell && !(remaining <= 0 && f < 1e-6);
            const double eff_mass = f * m4 + (1.0 - f) * remaining;
            const bool x4_ok = x4 > 1e-290 && x4 < 1e290;
            ok = ok && x4_ok;
            const double column_den3 = vdiv(eff_mass * comp_ratio, x4_ok ? x4 : 1.0);
            const double dm3dt = column_den3 * dx3;
            const double ratio = vdiv(m3, m4_s);
            const double cap_w = smoothstep(0, 1.0, ratio);
            const double capped_rate = vmin(dm3dt, dm4);
            const double injected = (1.0 - cap_w) * dm3dt + cap_w * capped_rate;
            dm3 = feeding ? ((f > 1e-6) ? injected : dm3dt) : 0.;
        }
        d[iM3] = dm3;

        // compute_dU2_dt
        double dU2;
        {
            const double shock_heating = dm2 * (Gamma - 1) * con::c2;
            const double x4_s = (x4 > 0) ? x4 : 1.0;
            const double dlnvdt = dlnv_r + ((x4 > 0) ? vdiv(dx4, x4_s) : 0.0);
            const double adiabatic_cooling = -(ad2 - 1) * dlnvdt * U2;
            dU2 = (1 - eps_rad) * shock_heating + adiabatic_cooling;
        }
        d[iU2] = dU2;
        // compute_dU3_dt
        double dU3;
        {
            const double x3_s = (x3 > 0) ? x3 : 1.0;
            const double dlnvdt = dlnv_r + ((x3 > 0) ? vdiv(dx3, x3_s) : 0.0);
            const double adiabatic_cooling = -(ad34 - 1) * dlnvdt * U3;
            const double shock_heating = dm3 * (Gamma34 - 1) * con::c2;
            dU3 = shock_heating + adiabatic_cooling;
        }
        d[iU3] = dU3;

        // compute_dGamma_dt
        {
            const double Gamma_eff2 = compute_effective_Gamma(ad2, Gamma);
            const double Gamma_eff3 = compute_effective_Gamma(ad34, Gamma);
            const double dGamma_eff2 = compute_effective_Gamma_dGamma(ad2, Gamma);
            const double dGamma_eff3 = compute_effective_Gamma_dGamma(ad34, Gamma);
            const double a = (Gamma - 1) * con::c2 * dm2 + (Gamma - Gamma4) * con::c2 * dm3 + Gamma_eff2 * dU2 +
                             Gamma_eff3 * dU3;
            const double b = (m2 + m3) * con::c2 + dGamma_eff2 * U2 + dGamma_eff3 * U3;
            const bool b_ok = b > 1e-290 && b < 1e290;
            const double q = vdiv(-a, b_ok ? b : 1.0);
            ok = ok && b_ok && fabs(q) < kInf;
            d[iG] = q;
        }
        return ok;
    }

    VAG_HD void rhs_general(const double* xr, double* d, double t) const {
        // projection onto the physical domain
        const double Gamma = vclamp(xr[iG], 1.0, Gamma4);
        const double m4 = xr[iM4];
        const double m3 = vclamp(xr[iM3], 0.0, vmax(m4, 0.0));
        const double x3 = vmax(xr[iX3], 0.0);
        const double U3 = vmax(xr[iU3], 0.0);
        const double x4 = xr[iX4], m2 = xr[iM2], U2 = xr[iU2], r = xr[iR], t_comv = xr[iT], eps4 = xr[iE4];

        const double u3 = vsqrt((Gamma - 1) * (Gamma + 1));  // Gamma clamped to [1, Gamma4]
        const double dr = u3 * (Gamma + u3) * con::c;
        const double dtc = Gamma + u3;
        d[iR] = dr;
        d[iT] = dtc;
        const double rho = medium_rho(m, r);
        const double dm2 = r * r * rho * dr;  // compute_dm2_dt :215-218
        d[iM2] = dm2;

        // compute_deps4_dt / compute_dm4_dt :220-250
        const double inject_w = smoothstep(m.T0 * 1.5, m.T0 * 0.5, t);
        double deps4 = 0, dm4 = 0;
        if (inject_w > 1e-6) {
            deps4 = inject_w * deps0_dt;
            dm4 = inject_w * dm0_dt;
        }
        const double deps_inj = jet_deps_dt(m, theta0, t);  // 0 without a magnetar
        deps4 += deps_inj;                                   // compute_deps4_dt :231-233
        d[iE4] = deps4;
        d[iM4] = dm4;

        // quantities several rate terms of the reference recompute: evaluated once here (same
        // expressions, so the values are identical)
        const double Gamma34 = compute_rel_Gamma(Gamma4, Gamma);
        const double ad2 = adiabatic_idx_fast(Gamma);
        const double ad34 = adiabatic_idx_fast(Gamma34);
        const double cs34 = compute_sound_speed_ad(Gamma34, ad34);
        const double cs34_dtc = cs34 * dtc;
        const double dlnv_r = vdiv(2 * dr, r);  // compute_adiabatic_cooling_rate2: 2 drdt / r  (r > 0)
        const double sigma = shell_sigma(eps4, m4);
        const double comp_ratio = compute_4vel_jump_ad(Gamma34, sigma, ad34);

        const double f = injection_efficiency(dm4);
        // compute_dx4_dt :205-213
        double dx4;
        {
            const double sound_expansion = cs4 * dtc;
            dx4 = (f > 1e-6) ? f * u4 + (1 - f) * sound_expansion : sound_expansion;
        }
        d[iX4] = dx4;

        // The nested conditions of the reference's rate terms are evaluated as straight-line code and selected
        // afterwards (same arithmetic on the taken side): with one warp per scheduler a branch ends the window
        // in which the independent chains of this right-hand side can overlap.
        // compute_dU2_dt :96-110 -- the radiative efficiency only needs (Gamma, rho, t_comv): start it early
        const double e_th = (Gamma - 1) * 4 * Gamma * rho * con::c2;
        const double eps_rad = radiative_efficiency(m.fwd, t_comv, Gamma, e_th);

        // compute_dx3_dt :124-179
        const double remaining = vmax(m4 - m3, 0.0);
        const bool has_shell = !(m4 <= 0);
        const double m4_s = has_shell ? m4 : 1.0;
        const double crossing_w = f + vdiv((1.0 - f) * remaining, m4_s);
        const double penetration = vdiv(Gamma * comp_ratio, Gamma4) - 1;
        const bool crossing_on = has_shell && !(crossing_w < 1e-6) && !(penetration <= 0);
        double dx3;
        {
            const double sound_expansion = cs34_dtc;
            const double pen_s = crossing_on ? penetration : 1.0;
            const double beta3 = vdiv(u3, Gamma);  // gamma_to_beta(Gamma)
            const double dx3dt = vdiv((Gamma4 - Gamma) * (Gamma4 + Gamma) * (1 + beta3) * con::c,
                                      Gamma4 * Gamma4 * (beta3 + beta4) * pen_s);
            double crossing = fabs(dx3dt * Gamma);
            const double va2 = vdiv(sigma, 1 + sigma);  // sigma >= 0
            const double cs2 = cs34 * cs34 / (con::c * con::c);
            const double v_ms = vsqrt(va2 + cs2 * (1 - va2)) * con::c;
            crossing = (penetration < 1) ? vmin(crossing, v_ms * dtc) : crossing;
            dx3 = crossing_on ? crossing_w * crossing + (1.0 - crossing_w) * sound_expansion : sound_expansion;
        }
        d[iX3] = dx3;

        // compute_dm3_dt :181-203
        double dm3;
        {
            const bool feeding = has_shell && !(remaining <= 0 && f < 1e-6);
            const double eff_mass = f * m4 + (1.0 - f) * remaining;
            const double column_den3 = eff_mass * comp_ratio / x4;  // IEEE: x4 is an unclamped state
            const double dm3dt = column_den3 * dx3;
            const double ratio = vdiv(m3, m4_s);
            const double cap_w = smoothstep(0, 1.0, ratio);
            const double capped_rate = vmin(dm3dt, dm4);
            const double injected = (1.0 - cap_w) * dm3dt + cap_w * capped_rate;
            dm3 = feeding ? ((f > 1e-6) ? injected : dm3dt) : 0.;
        }
        d[iM3] = dm3;

        // compute_dU2_dt :96-110
        double dU2;
        {
            const double shock_heating = dm2 * (Gamma - 1) * con::c2;
            const double x4_s = (x4 > 0) ? x4 : 1.0;
            const double dlnvdt = dlnv_r + ((x4 > 0) ? vdiv(dx4, x4_s) : 0.0);
            const double adiabatic_cooling = -(ad2 - 1) * dlnvdt * U2;
            dU2 = (1 - eps_rad) * shock_heating + adiabatic_cooling;
        }
        d[iU2] = dU2;
        // compute_dU3_dt :112-122
        double dU3;
        {
            const double x3_s = (x3 > 0) ? x3 : 1.0;
            const double dlnvdt = dlnv_r + ((x3 > 0) ? vdiv(dx3, x3_s) : 0.0);
            const double adiabatic_cooling = -(ad34 - 1) * dlnvdt * U3;
            const double shock_heating = dm3 * (Gamma34 - 1) * con::c2;
            dU3 = shock_heating + adiabatic_cooling;
        }
        d[iU3] = dU3;

        // compute_dGamma_dt :62-94
        {
            const double Gamma_eff2 = compute_effective_Gamma(ad2, Gamma);
            const double Gamma_eff3 = compute_effective_Gamma(ad34, Gamma);
            const double dGamma_eff2 = compute_effective_Gamma_dGamma(ad2, Gamma);
            const double dGamma_eff3 = compute_effective_Gamma_dGamma(ad34, Gamma);
            const double deps_dt = deps_inj;  // :75-77
            const double a = (Gamma - 1) * con::c2 * dm2 + (Gamma - Gamma4) * con::c2 * dm3 + Gamma_eff2 * dU2 +
                             Gamma_eff3 * dU3 - deps_dt;
            const double b = (m2 + m3) * con::c2 + dGamma_eff2 * U2 + dGamma_eff3 * U3;
            const double q = -a / b;
            d[iG] = (b == 0 || isnan(q) || isinf(q)) ? 0 : q;
        }
    }

    // set_init_state: reverse-shock.tpp:314-357 (+ compute_init_comv_shell_width :381-390)
    VAG_HD void set_init_state(double* x, double t0) const {
        const double beta4 = gamma_to_beta(Gamma4);
        x[iR] = beta4 * con::c * t0 * Gamma4 * Gamma4 * (1 + beta4);
        x[iT] = x[iR] / sqrt((Gamma4 - 1) * (Gamma4 + 1)) / con::c;
        const double dt = vmin(t0, m.T0);
        x[iE4] = deps0_dt * dt;
        x[iM4] = dm0_dt * dt;
        if (t0 < m.T0) {
            x[iX4] = Gamma4 * t0 * beta4 * con::c;
        } else {
            const double cs = compute_sound_speed(Gamma4);
            x[iX4] = Gamma4 * m.T0 * beta4 * con::c + cs * (t0 - m.T0) * Gamma4;
        }
        x[iM2] = enclosed_mass_numeric(m, x[iR]);
        const double m_jet_total = dm0_dt * m.T0;
        if (m_jet_total > 0 && x[iM2] > 0) {
            x[iG] = Gamma4 / (1 + x[iM2] / m_jet_total);
        } else {
            x[iG] = Gamma4;
        }
        const double ad_idx = adiabatic_idx(x[iG]);
        x[iU2] = enclosed_thermal_energy_numeric(m, x[iR], x[iG], ad_idx, m.fwd.radiative ? m.fwd.eps_e : 0.0);
        const double Gamma34 = compute_rel_Gamma(Gamma4, x[iG]);
        if (Gamma34 > 1 && x[iM4] > 0 && x[iX4] > 0) {
            constexpr double seed_frac = 1e-8;
            const double sigma = shell_sigma(x[iE4], x[iM4]);
            const double comp_ratio = compute_4vel_jump(Gamma34, sigma);
            x[iX3] = x[iX4] * seed_frac;
            x[iM3] = x[iM4] * comp_ratio * x[iX3] / x[iX4];
            x[iU3] = (Gamma34 - 1) * x[iM3] * con::c2;
        } else {
            x[iM3] = 0;
            x[iU3] = 0;
            x[iX3] = 0;
        }
    }

    // save_cross_state: reverse-shock.tpp:298-312
    VAG_HD void save_cross_state(const double* x) {
        r_x = x[iR];
        u_x = sqrt((x[iG] - 1) * (x[iG] + 1));
        V3_comv_x = r_x * r_x * x[iX3];
        const double sigma4 = shell_sigma(x[iE4], x[iM4]);
        const double comp_ratio34 = compute_compression(Gamma4, x[iG], sigma4);
        const double rho4 = x[iM4] / (x[iR] * x[iR] * x[iX4]);
        rho3_x = rho4 * comp_ratio34;
        const double B4 = compute_upstr_B(rho4, sigma4);
        B3_ordered_x = B4 * comp_ratio34;
    }

    // save_rvs_shock_state: reverse-shock.tpp:403-426
    VAG_HD void save_rvs_state(double eps_B, const ShockRow& s, int k, int injection_idx, const double* x) const {
        double Gamma3_th, B3;
        if (k <= injection_idx) {
            const double sigma4 = shell_sigma(x[iE4], x[iM4]);
            const double comp_ratio34 = compute_compression(Gamma4, x[iG], sigma4);
            const double rho4 = x[iM4] / (x[iR] * x[iR] * x[iX4]);
            Gamma3_th = compute_Gamma_therm(x[iU3], x[iM3], true);
            const double B4 = compute_upstr_B(rho4, sigma4);
            B3 = compute_downstr_B(eps_B, rho4, B4, Gamma3_th, comp_ratio34);
        } else {
            const double V3_comv = x[iR] * x[iR] * x[iX3];
            const double comp_ratio = V3_comv_x / V3_comv;
            Gamma3_th = compute_Gamma_therm(x[iU3], x[iM3]);
            B3 = compute_downstr_B(eps_B, rho3_x, B3_ordered_x, Gamma3_th, comp_ratio);
        }
        s.t_comv[k] = x[iT];
        s.r[k] = x[iR];
        s.Gamma[k] = x[iG];
        s.Gamma_th[k] = Gamma3_th;
        s.B[k] = B3;
        s.N_p[k] = x[iM3] / con::mp;
    }
};

// reverse_shock_early_extrap: reverse-shock.tpp:428-467, split into the scan for the first thermally
// resolved node (extrap_scan: strided over `nthr` cooperating threads, combine the results with min)
// and the per-node rewrite (extrap_cell: independent for every k < idx_cut)
VAG_HD int extrap_scan(const ShockRow& s, int n_t, int tid, int nthr) {
    for (int k = tid; k < n_t; k += nthr)
        if (s.Gamma_th[k] > con::gamma_therm_cut) return k;
    return n_t;
}
VAG_HD bool extrap_applies(int idx_cut, int n_t, int injection_idx) {
    constexpr int offset = 2;
    return !(idx_cut == 0 || idx_cut >= n_t - offset || idx_cut >= injection_idx);
}
VAG_HD void extrap_cell(const ShockRow& s, int idx_cut, int k) {
    constexpr int offset = 2;
    const double log2_r = fast_log2(s.r[idx_cut]);
    const double log2_Gamma_th = fast_log2(s.Gamma_th[idx_cut] - 1);
    const double log2_B = fast_log2(s.B[idx_cut]);
    const double log2_N_p = fast_log2(s.N_p[idx_cut]);
    const double dl = fast_log2(s.r[idx_cut + offset]) - log2_r;
    const double gamma_slope = (fast_log2(s.Gamma_th[idx_cut + offset] - 1) - log2_Gamma_th) / dl;
    const double B_slope = (fast_log2(s.B[idx_cut + offset]) - log2_B) / dl;
    const double N_p_slope = (fast_log2(s.N_p[idx_cut + offset]) - log2_N_p) / dl;
    const double dlog2_r = fast_log2(s.r[k]) - log2_r;
    s.Gamma_th[k] = 1 + fast_exp2(log2_Gamma_th + gamma_slope * dlog2_r);
    s.B[k] = fast_exp2(log2_B + B_slope * dlog2_r);
    s.N_p[k] = fast_exp2(log2_N_p + N_p_slope * dlog2_r);
}

// grid_solve_shock_pair: reverse-shock.tpp:511-591, one dopri5 attempt per loop iteration (see
// solve_fwd_row).  Leaves raw node states + the RowDyn record; finish_pair_cell completes the tables.
VAG_HD int solve_pair_row(const ModelCfg& m, double theta, double t_dec, const double* t, int n_t, const ShockRow& sf,
                          const ShockRow& sr, const RawRow& raw, RowDyn& rd, double* col, int col_stride) {
    FREqn eqn(m, theta);
    double x[FREqn::N];
    const double t0 = vmin(t[0], vmin(0.01 * unit::sec, 0.1 * t_dec));
    eqn.set_init_state(x, t0);
    int injection_idx = n_t;
    rd.injection_idx = n_t;
    rd.V3_comv_x = rd.rho3_x = rd.B3_ordered_x = 0;

    constexpr double RS_Gamma_limit = 1.03;
    if (x[FREqn::iG] <= RS_Gamma_limit) {
        set_stopping_row(sf, n_t, x[FREqn::iT], x[FREqn::iR]);
        set_stopping_row(sr, n_t, x[FREqn::iT], x[FREqn::iR]);
        rd.n_saved = -1;
        return 0;
    }
    double rtol = m.rtol;
    if (eqn.shell_sigma(x[FREqn::iE4], x[FREqn::iM4]) > 0) rtol *= dflt::magnetized_rtol_factor;

    Dopri5S<FREqn::N> st;
    st.initialize(col, col_stride, x, t0, 1e-9 * t0, rtol);

    int k = 0;
    for (; k < n_t && t[k] < t0; k++) {
        eqn.set_init_state(x, t[k]);
#pragma unroll
        for (int c = 0; c < FREqn::N; ++c) raw.c[c][k] = x[c];
    }

    bool reverse_shock_crossing = true;
    bool injection_idx_pending = false;
    double t_cross = 0;
    double t_step_start = t0;
    const double t_back = t[n_t - 1];
    int status = 0, fails = 0, steps = 0;
    const int max_fails = m.max_ode_fails, max_steps = m.max_ode_steps;  // in registers: the loop stores to global memory
    // the next two lattice nodes ride in registers (+inf behind the last one): the loop test never waits for a load
    double t_next = k < n_t ? t[k] : kInf, t_next2 = k + 1 < n_t ? t[k + 1] : kInf;
    st.begin(eqn);
    while (st.t <= t_back) {
        if (!st.try_step(eqn)) {
            if (++fails >= max_fails) {
                status |= VAG_ST_ODE_FAIL500;
                break;
            }
            continue;
        }
        fails = 0;
        if (++steps > max_steps) {
            status |= VAG_ST_ODE_STEP_CAP;
            break;
        }
        if (st.t + st.dt == st.t) {
            status |= VAG_ST_ODE_STALLED;
            break;
        }
        if (reverse_shock_crossing && eqn.crossing_complete(st.x, st.t)) {
            // locate_crossing_time: reverse-shock.tpp:482-495.  crossing_complete reads only m3 and m4
            // of the dense output, so the bisection interpolates just those two components.
            double t_lo = t_step_start, t_hi = st.t;
            double p3[6], p4[6];
            st.dense_poly(FREqn::iM3, p3);
            st.dense_poly(FREqn::iM4, p4);
            for (int iter = 0; iter < 100 && (t_hi - t_lo) > 1e-12 * t_hi; ++iter) {
                const double t_mid = 0.5 * (t_lo + t_hi);
                const double th = st.dense_theta(t_mid);
                if (eqn.crossing_complete_m(st.dense_poly_eval(p3, th), st.dense_poly_eval(p4, th), t_mid)) {
                    t_hi = t_mid;
                } else {
                    t_lo = t_mid;
                }
            }
            st.calc_state(t_hi, x);
            t_cross = t_hi;
            eqn.save_cross_state(x);
            rd.V3_comv_x = eqn.V3_comv_x;
            rd.rho3_x = eqn.rho3_x;
            rd.B3_ordered_x = eqn.B3_ordered_x;
            reverse_shock_crossing = false;
            injection_idx_pending = true;
        }
        t_step_start = st.t;
        while (st.t > t_next) {  // t_next = t[k]
            st.calc_state(t_next, x);
            if (injection_idx_pending && t_next >= t_cross) {
                injection_idx = k > 0 ? k : 1;
                injection_idx_pending = false;
            }
#pragma unroll
            for (int c = 0; c < FREqn::N; ++c) raw.c[c][k] = x[c];
            ++k;
            t_next = t_next2;
            t_next2 = k + 1 < n_t ? t[k + 1] : kInf;
        }
        st.advance();
    }
    rd.n_saved = k;
    rd.injection_idx = injection_idx;
    return status;
}

// save_fwd_shock_state + save_rvs_shock_state of node k from its raw state (pair rows)
VAG_HD void finish_pair_cell(const ModelCfg& m, double theta, const RowDyn& rd, const ShockRow& sf, const ShockRow& sr,
                             const RawRow& raw, int k) {
    if (rd.n_saved < 0) return;
    if (k >= rd.n_saved) {
        default_cell(sf, k);
        default_cell(sr, k);
        return;
    }
    double x[FREqn::N];
#pragma unroll
    for (int c = 0; c < FREqn::N; ++c) x[c] = raw.c[c][k];
    FREqn eqn(m, theta);
    eqn.V3_comv_x = rd.V3_comv_x;
    eqn.rho3_x = rd.rho3_x;
    eqn.B3_ordered_x = rd.B3_ordered_x;
    save_fwd_state(m, m.fwd.eps_B, sf, k, x[FREqn::iG], x[FREqn::iM2], x[FREqn::iU2], x[FREqn::iR], x[FREqn::iT]);
    eqn.save_rvs_state(m.rvs.eps_B, sr, k, rd.injection_idx, x);
}

}  // namespace vag
