// vag_model.cuh -- one parameter set in code units + the typed jet / medium profiles.
//
// Restates (from scratch) the unit conversions of the reference's Python-facing factories
// (pybind/pymodel.cpp:47-186, pybind/pymodel.h:190-204,613-649) and the inline profile classes
// TophatJet / GaussianJet / PowerLawJet (src/environment/jet.h:84-257) and ISM / Wind
// (src/environment/medium.h:50-140) as one POD evaluated by enumerated type on the device.
#pragma once

#include "../../include/vag.h"
#include "vag_common.cuh"
#include "vag_libm.cuh"

namespace vag {

struct RadCfg {
    double eps_e, eps_B, p, xi_e;
    int ssc, kn;
    int radiative;
    // RadiativeEfficiency precomputed coefficients (src/dynamics/shock-physics.h:251-261)
    double gamma_m_coeff, gamma_c_coeff, eps_e_rad;
};

struct ModelCfg {
    // jet
    int jet_type;
    double theta_c, eps_k0, Gamma0, k_e, k_g, gauss_norm, T0;
    int ejecta;           // the reference builds this jet as an `Ejecta` of std::function profiles (pymodel.cpp:47-146)
    double gauss_spread;  // math::gaussian: -2 theta_c^2 (jet.h:376-379)
    double theta_w, E_iso, E_iso_w, Gamma0_w, sigma0;  // Ejecta-family profiles work on E_iso [erg] heights
    int has_magnetar;
    double mag_L, mag_t0, mag_q;
    int spreading;   // lateral spreading of a forward-shock-only model (the pair solver has none, reverse-shock.tpp)
    int axisymmetric;  // Model(axisymmetric=): False = full unmirrored phi grid, every phi row observed (jet_3d = 1)
    int structured;  // jet.spreading: Symmetry::structured, every theta row solved on its own lattice (mesh.h:125-130)  // L0 [code units / 4 pi], t0 [s], q
    // medium
    int medium_type;
    double rho_ism, wind_A, wind_r02;
    // Wind(k_m != 2): generic-Medium profile in CGS, rho = A_cgs / (r0k_cgs + r_cgs^k) + rho_ism_cgs
    int wind_generic;
    double wind_k, wind_A_cgs, wind_r0k_cgs, wind_rho_ism_cgs;
    // observer
    double lumi_dist, z, theta_v;
    // radiation
    RadCfg fwd, rvs;
    int has_rvs;
    // numerics
    double phi_resol, theta_resol, t_resol, rtol;
    int max_ode_steps, max_ode_fails;  // defaults::solver::max_ode_steps, Boost's 500 (vag_debug_set_ode_limits lowers them in tests)
};

VAG_HD RadCfg make_rad(const vag_radiation& r, int radiative) {
    RadCfg c;
    c.eps_e = r.eps_e;
    c.eps_B = r.eps_B;
    c.p = r.p;
    c.xi_e = r.xi_e;
    c.ssc = r.ssc;
    c.kn = r.kn;
    c.radiative = radiative;
    c.gamma_m_coeff = (r.p - 2) / (r.p - 1) * r.eps_e * con::mp / con::me / r.xi_e;
    c.gamma_c_coeff = 6 * con::pi * con::me * con::c / con::sigmaT / (8 * con::pi * r.eps_B);
    c.eps_e_rad = radiative ? r.eps_e : 0;
    return c;
}

VAG_HD ModelCfg make_cfg(const vag_params& p) {
    ModelCfg m;
    m.jet_type = p.jet_type;
    m.theta_c = p.theta_c;
    m.eps_k0 = (p.E_iso * unit::erg) / (4 * con::pi);  // jet.h:97,142,205
    m.Gamma0 = p.Gamma0;
    m.k_e = p.k_e;
    m.k_g = p.k_g;
    m.gauss_norm = -1 / (2 * p.theta_c * p.theta_c);  // jet.h:141
    m.ejecta = (p.jet_type >= VAG_JET_TWO_COMPONENT || p.sigma0 > 0 || p.has_magnetar) ? 1 : 0;
    m.gauss_spread = -2 * p.theta_c * p.theta_c;
    m.T0 = p.duration * unit::sec;
    m.theta_w = p.theta_w;
    m.E_iso = p.E_iso;
    m.E_iso_w = p.E_iso_w;
    m.Gamma0_w = p.Gamma0_w;
    m.sigma0 = p.sigma0;
    m.spreading = (p.spreading && !p.has_rvs) ? 1 : 0;
    m.structured = p.spreading ? 1 : 0;
    m.axisymmetric = p.axisymmetric ? 1 : 0;
    m.has_magnetar = p.has_magnetar;
    // convert_unit_jet (pybind/pymodel.cpp:196-199): deps_dt_cgs(t / unit::sec) * (unit::erg / (4 pi unit::sec))
    m.mag_L = p.magnetar_L0;
    m.mag_t0 = p.magnetar_t0;
    m.mag_q = p.magnetar_q;
    m.medium_type = p.medium_type;
    const double n_ism = p.n_ism / unit::cm3;
    m.rho_ism = n_ism * con::mp;  // medium.h:52,98
    m.wind_A = 0;
    m.wind_r02 = 0;
    if (p.medium_type == VAG_MEDIUM_WIND) {
        m.wind_A = p.A_star * 5e11 * unit::g / unit::cm;           // medium.h:98
        m.wind_r02 = m.wind_A / ((p.n0 / unit::cm3) * 1.3 * con::mp);  // 0 when n0 = inf
    }
    m.wind_generic = 0;
    m.wind_k = 2;
    m.wind_A_cgs = m.wind_r0k_cgs = m.wind_rho_ism_cgs = 0;
    if (p.medium_type == VAG_MEDIUM_WIND && p.wind_k_m > 0 && p.wind_k_m != 2) {  // pybind/pymodel.cpp:169-185
        constexpr double r0_cgs = 1e17;
        const double mp_cgs = con::mp / unit::g;
        m.wind_generic = 1;
        m.wind_k = p.wind_k_m;
        m.wind_A_cgs = p.A_star * 5e11 * pow(r0_cgs, p.wind_k_m - 2);
        m.wind_rho_ism_cgs = p.n_ism * mp_cgs;
        m.wind_r0k_cgs = m.wind_A_cgs / (p.n0 * 1.3 * mp_cgs);
    }
    m.lumi_dist = p.lumi_dist * unit::cm;
    m.z = p.z;
    m.theta_v = p.theta_obs;
    m.fwd = make_rad(p.fwd, p.radiative_fireball);
    m.rvs = make_rad(p.rvs, p.radiative_fireball);
    m.has_rvs = p.has_rvs;
    const bool r = p.has_rvs != 0;
    m.phi_resol = p.phi_resol > 0 ? p.phi_resol : dflt::phi_resolution;
    m.theta_resol = p.theta_resol > 0 ? p.theta_resol : (r ? dflt::rvs_theta_resolution : dflt::theta_resolution);
    m.t_resol = p.t_resol > 0 ? p.t_resol : (r ? dflt::rvs_time_resolution : dflt::time_resolution);
    m.rtol = p.rtol > 0 ? p.rtol : dflt::dynamics_rtol;
    m.max_ode_steps = dflt::max_ode_steps;
    m.max_ode_fails = 500;
    return m;
}

// ---- jet profiles (phi-independent for every typed variant) --------------------------------
// Spelled out in contraction-proof operations (vag_libm.cuh) with the host libm's exp / exp2 / log2: the grid
// builder's CDF quadrature needs the profile bit-identical to the reference build's (vag_grid.cuh), and every
// other caller is served by the same values.  Operation order = the reference build's instruction sequence
// (jet.h:116,175-177,243-245 as compiled: Gaussian Gamma0 is one fma, the power law keeps its true divisions).
// RECIP: where g++ inlined PowerLawJet::Gamma0 several times into one function (the theta-grid sampler's seven stage
// evaluations), -freciprocal-math turned theta / theta_c into theta * (1 / theta_c); single call sites divide.
template <bool RECIP = false>
VAG_HD double jet_powlaw(const ModelCfg& m, double theta, double k) {  // fast_pow(theta / theta_c, k) (fast-math.h:147-149)
    const double ratio = RECIP ? gl::mul(theta, gl::div(1.0, m.theta_c)) : gl::div(theta, m.theta_c);
    return gl::exp2(gl::mul(gl::log2(ratio), k));
}
template <bool RECIP = false>
VAG_HD double jet_Gamma0(const ModelCfg& m, double theta) {
    switch (m.jet_type) {
        // the Ejecta forms (a magnetar or sigma0 > 0 turns a typed jet into math::*_plus_one profiles, pymodel.cpp:47-95):
        // (Gamma0 - 1) + 1 is not always Gamma0, and math::gaussian divides by -2 theta_c^2 where GaussianJet multiplies
        case VAG_JET_TOPHAT:
            if (m.ejecta) return gl::add(theta < m.theta_c ? gl::sub(m.Gamma0, 1.0) : 0.0, 1.0);  // jet.h:360-367
            return theta < m.theta_c ? m.Gamma0 : 1;  // jet.h:116
        case VAG_JET_GAUSSIAN:
            if (m.ejecta)  // jet.h:376-384
                return gl::fma(gl::exp(gl::div(gl::mul(theta, theta), m.gauss_spread)), gl::sub(m.Gamma0, 1.0), 1.0);
            return gl::fma(gl::sub(m.Gamma0, 1.0), gl::exp(gl::mul(gl::mul(theta, theta), m.gauss_norm)), 1.0);  // jet.h:175-177
        case VAG_JET_POWERLAW:  // jet.h:243-245, 395-401
            if (m.ejecta) return gl::add(gl::div(gl::sub(m.Gamma0, 1.0), gl::add(jet_powlaw<false>(m, theta, m.k_g), 1.0)), 1.0);
            return gl::add(gl::div(gl::sub(m.Gamma0, 1.0), gl::add(jet_powlaw<RECIP>(m, theta, m.k_g), 1.0)), 1.0);
        // Ejecta family: Gamma0 = profile(Gamma0 - 1, ...) + 1 (math::*_plus_one, jet.h:403-470)
        case VAG_JET_TWO_COMPONENT:
            return gl::add((theta <= m.theta_c) ? gl::sub(m.Gamma0, 1.0) : (theta <= m.theta_w) ? gl::sub(m.Gamma0_w, 1.0) : 0., 1.0);
        case VAG_JET_STEP_POWERLAW:
            return gl::add((theta <= m.theta_c) ? gl::sub(m.Gamma0, 1.0)
                                                : gl::mul(gl::sub(m.Gamma0_w, 1.0), jet_powlaw(m, theta, -m.k_g)), 1.0);
        default:  // VAG_JET_POWERLAW_WING
            return gl::add((theta <= m.theta_c) ? 0. : gl::mul(gl::sub(m.Gamma0_w, 1.0), jet_powlaw(m, theta, -m.k_g)), 1.0);
    }
}

VAG_HD double jet_eps_k(const ModelCfg& m, double theta) {
    if (m.ejecta) {
        // Ejecta: E_iso(theta) [erg] (math::* profile, jet.h:360-470) * (unit::erg / 4 pi)  (convert_unit_jet, pymodel.cpp:188-211)
        double E;
        switch (m.jet_type) {
            case VAG_JET_TOPHAT:
                E = theta < m.theta_c ? m.E_iso : 0.;
                break;
            case VAG_JET_GAUSSIAN:
                E = gl::mul(m.E_iso, gl::exp(gl::div(gl::mul(theta, theta), m.gauss_spread)));
                break;
            case VAG_JET_POWERLAW:
                E = gl::div(m.E_iso, gl::add(jet_powlaw(m, theta, m.k_e), 1.0));
                break;
            case VAG_JET_TWO_COMPONENT:
                E = (theta <= m.theta_c) ? m.E_iso : (theta <= m.theta_w) ? m.E_iso_w : 0.;
                break;
            case VAG_JET_STEP_POWERLAW:
                E = (theta <= m.theta_c) ? m.E_iso : gl::mul(m.E_iso_w, jet_powlaw(m, theta, -m.k_e));
                break;
            default:
                E = (theta <= m.theta_c) ? 0. : gl::mul(m.E_iso_w, jet_powlaw(m, theta, -m.k_e));
                break;
        }
        return gl::mul(E, unit::erg / (4 * con::pi));
    }
    switch (m.jet_type) {
        case VAG_JET_TOPHAT:
            return theta < m.theta_c ? m.eps_k0 : 0;  // jet.h:107
        case VAG_JET_GAUSSIAN:
            return gl::mul(m.eps_k0, gl::exp(gl::mul(gl::mul(theta, theta), m.gauss_norm)));  // jet.h:164-166
        default:  // VAG_JET_POWERLAW
            return gl::div(m.eps_k0, gl::add(jet_powlaw(m, theta, m.k_e), 1.0));  // jet.h:232-234
    }
}

// energy injection rate per solid angle: math::magnetar_injection (jet.h:518-528) behind the unit wrapper
// of convert_unit_jet (pybind/pymodel.cpp:196-199); t in code units
VAG_HD double jet_deps_dt(const ModelCfg& m, double theta, double t) {
    if (!m.has_magnetar) return 0.;
    double cgs = 0.;
    if (theta <= m.theta_c) {
        const double tt = 1 + (t / unit::sec) / m.mag_t0;
        cgs = m.mag_L * fast_pow(tt, -m.mag_q);
    }
    return cgs * (unit::erg / (4 * con::pi * unit::sec));
}

// ---- medium profiles (isotropic) ------------------------------------------------------------
VAG_HD double medium_rho(const ModelCfg& m, double r) {
    if (m.medium_type == VAG_MEDIUM_ISM) return m.rho_ism;  // medium.h:58
    if (m.wind_generic)                                     // PyWind's Medium behind convert_unit_medium
        return (m.wind_A_cgs / (m.wind_r0k_cgs + pow(r / unit::cm, m.wind_k)) + m.wind_rho_ism_cgs) * (unit::g / unit::cm3);
    return m.wind_A / (m.wind_r02 + r * r) + m.rho_ism;     // medium.h:107-109
}

// analytic enclosed mass per solid angle: ISM medium.h:61, Wind medium.h:115-127
VAG_HD double medium_mass(const ModelCfg& m, double r) {
    if (m.medium_type == VAG_MEDIUM_ISM) return m.rho_ism * r * r * r / 3.0;
    double mass = m.rho_ism * r * r * r / 3.0;
    if (m.wind_A != 0) {
        if (m.wind_r02 > 0) {
            const double a = sqrt(m.wind_r02);
            mass += m.wind_A * (r - a * atan(r / a));
        } else {
            mass += m.wind_A * r;
        }
    }
    return mass;
}

}  // namespace vag
