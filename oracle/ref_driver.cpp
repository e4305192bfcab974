// TEST INFRASTRUCTURE ONLY (oracle/): a thin C-ABI driver around the UNMODIFIED reference
// (VegasAfterglow C++ sources under /root/reference, compiled by oracle/Makefile into
// oracle/_ref/libvagref.so).  All numerics executed here are the reference's own functions;
// this file only mirrors the orchestration of PyModel::compute_emission /
// single_shock_emission (pybind/pymodel.h:873-961) without the pybind11 array types, so that
//   (a) tests can read every intermediate table (Coord, Shock, injection_idx, observer grids),
//   (b) bench.py can time the reference on all host cores with std::thread instead of through
//       the GIL (the reference's best case: SURVEY.md section 8d "pure-C++ std::thread driver").
// Nothing in vegasafterglow_b200/ links to, imports or calls this library.
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <optional>
#include <thread>
#include <vector>

#include "../include/vag.h"
#include "shock_dispatch.h"  // reference: pybind/shock_dispatch.h (pulls include/afterglow.h)

namespace {

struct RefModel {
    JetVariant jet;
    MediumVariant medium;
    Real lumi_dist, z, theta_obs;
    RadParams fwd_rad, rvs_rad;
    bool fwd_ssc, fwd_kn, rvs_ssc, rvs_kn, has_rvs;
    Real phi_resol, theta_resol, t_resol, rtol;
    bool axisymmetric;
    Real theta_w{con::pi / 2};
};

// Ejecta-family jets: the named factories of pybind/pymodel.cpp:97-146 followed by convert_unit_jet
// (pymodel.cpp:188-211).  A constant magnetisation sigma0 > 0 is expressed the only way the reference
// can express it: an Ejecta with a sigma0 profile (tests/python/golden/regenerate.py:140-148).
JetVariant make_ejecta(const vag_params& p) {
    Ejecta jet;
    const Real tc = p.theta_c;
    switch (p.jet_type) {
        case VAG_JET_TOPHAT:
            jet.eps_k = math::tophat(tc, p.E_iso);
            jet.Gamma0 = math::tophat_plus_one(tc, p.Gamma0 - 1);
            break;
        case VAG_JET_GAUSSIAN:
            jet.eps_k = math::gaussian(tc, p.E_iso);
            jet.Gamma0 = math::gaussian_plus_one(tc, p.Gamma0 - 1);
            break;
        case VAG_JET_POWERLAW:
            jet.eps_k = math::powerlaw(tc, p.E_iso, p.k_e);
            jet.Gamma0 = math::powerlaw_plus_one(tc, p.Gamma0 - 1, p.k_g);
            break;
        case VAG_JET_TWO_COMPONENT:
            jet.eps_k = math::two_component(tc, p.theta_w, p.E_iso, p.E_iso_w);
            jet.Gamma0 = math::two_component_plus_one(tc, p.theta_w, p.Gamma0 - 1, p.Gamma0_w - 1);
            break;
        case VAG_JET_STEP_POWERLAW:
            jet.eps_k = math::step_powerlaw(tc, p.E_iso, p.E_iso_w, p.k_e);
            jet.Gamma0 = math::step_powerlaw_plus_one(tc, p.Gamma0 - 1, p.Gamma0_w - 1, p.k_g);
            break;
        default:
            jet.eps_k = math::powerlaw_wing(tc, p.E_iso_w, p.k_e);
            jet.Gamma0 = math::powerlaw_wing_plus_one(tc, p.Gamma0_w - 1, p.k_g);
            break;
    }
    if (p.sigma0 > 0) jet.sigma0 = math::isotropic(p.sigma0);
    jet.spreading = p.spreading != 0;
    jet.T0 = p.duration;
    // initialize_ejecta (pybind/pymodel.cpp:38-45)
    if (p.has_magnetar) jet.deps_dt = math::magnetar_injection(p.magnetar_t0, p.magnetar_q, p.magnetar_L0, tc);
    // convert_unit_jet
    const auto eps_k_cgs = jet.eps_k;
    jet.eps_k = [=](Real phi, Real theta) { return eps_k_cgs(phi, theta) * (unit::erg / (4 * con::pi)); };
    const auto deps_dt_cgs = jet.deps_dt;
    jet.deps_dt = [=](Real phi, Real theta, Real t) {
        return deps_dt_cgs(phi, theta, t / unit::sec) * (unit::erg / (4 * con::pi * unit::sec));
    };
    jet.T0 *= unit::sec;
    return jet;
}

// unit conversions exactly as the Py* factories do (pybind/pymodel.cpp:47-186, pymodel.h:190-204)
JetVariant make_jet(const vag_params& p) {
    if (p.jet_type >= VAG_JET_TWO_COMPONENT || p.sigma0 > 0 || p.has_magnetar) return make_ejecta(p);
    const Real T0 = p.duration * unit::sec;
    switch (p.jet_type) {
        case VAG_JET_TOPHAT:
            return TophatJet(p.theta_c, p.E_iso * unit::erg, p.Gamma0, p.spreading != 0, T0);
        case VAG_JET_GAUSSIAN:
            return GaussianJet(p.theta_c, p.E_iso * unit::erg, p.Gamma0, p.spreading != 0, T0);
        default:
            return PowerLawJet(p.theta_c, p.E_iso * unit::erg, p.Gamma0, p.k_e, p.k_g, p.spreading != 0, T0);
    }
}

MediumVariant make_medium(const vag_params& p) {
    if (p.medium_type == VAG_MEDIUM_ISM) {
        return ISM(p.n_ism / unit::cm3);
    }
    if (p.wind_k_m > 0 && p.wind_k_m != 2) {
        // PyWind, general k_m (pybind/pymodel.cpp:169-185) followed by convert_unit_medium (:211-222)
        const Real k_m = p.wind_k_m;
        constexpr Real r0_cgs = 1e17;
        const Real mp_cgs = con::mp / unit::g;
        const Real A_cgs = p.A_star * 5e11 * std::pow(r0_cgs, k_m - 2);
        const Real rho_ism_cgs = p.n_ism * mp_cgs;
        const Real r0k_cgs = A_cgs / (p.n0 * 1.3 * mp_cgs);
        Medium medium;
        const auto rho_cgs = [=](Real /*phi*/, Real /*theta*/, Real r) noexcept {
            return A_cgs / (r0k_cgs + std::pow(r, k_m)) + rho_ism_cgs;
        };
        medium.rho = [=](Real phi, Real theta, Real r) { return rho_cgs(phi, theta, r / unit::cm) * (unit::g / unit::cm3); };
        medium.isotropic = true;
        return medium;
    }
    return Wind(p.A_star, p.n_ism / unit::cm3, p.n0 / unit::cm3);
}

RefModel build(const vag_params& p) {
    RefModel m{make_jet(p), make_medium(p)};
    m.lumi_dist = p.lumi_dist * unit::cm;
    m.z = p.z;
    m.theta_obs = p.theta_obs;
    m.fwd_rad = RadParams{p.fwd.eps_e, p.fwd.eps_B, p.fwd.p, p.fwd.xi_e};
    m.rvs_rad = RadParams{p.rvs.eps_e, p.rvs.eps_B, p.rvs.p, p.rvs.xi_e};
    m.fwd_rad.radiative = p.radiative_fireball != 0;
    m.rvs_rad.radiative = p.radiative_fireball != 0;
    m.fwd_ssc = p.fwd.ssc;
    m.fwd_kn = p.fwd.kn;
    m.rvs_ssc = p.rvs.ssc;
    m.rvs_kn = p.rvs.kn;
    m.has_rvs = p.has_rvs != 0;
    const bool r = m.has_rvs;
    m.phi_resol = p.phi_resol > 0 ? p.phi_resol : defaults::grid::phi_resolution;
    m.theta_resol =
        p.theta_resol > 0 ? p.theta_resol : (r ? defaults::grid::rvs_theta_resolution : defaults::grid::theta_resolution);
    m.t_resol = p.t_resol > 0 ? p.t_resol : (r ? defaults::grid::rvs_time_resolution : defaults::grid::time_resolution);
    m.rtol = p.rtol > 0 ? p.rtol : defaults::solver::dynamics_rtol;
    m.axisymmetric = p.axisymmetric != 0;
    return m;
}

using XT = xt::xarray<Real>;

// mirror of single_shock_emission (pybind/pymodel.h:873-920)
template <typename Func>
void shock_emission(Shock const& shock, Coord const& coord, Array const& t_obs, Array const& nu_obs, Observer& obs,
                    bool ssc, bool kn, XT& out_sync, XT& out_ssc, bool& has_ssc, Func&& flux_func) {
    auto syn_e = generate_syn_electrons(shock, coord);
    auto syn_ph = generate_syn_photons(shock, syn_e, coord);
    if (ssc) {
        if (kn) {
            KN_cooling(syn_e, syn_ph, shock, coord);
        } else {
            Thomson_cooling(syn_e, syn_ph, shock, coord);
        }
    }
    out_sync = flux_func(obs, t_obs, nu_obs, syn_ph);
    has_ssc = false;
    if (ssc) {
        const Real lg2_1pz = fast_log2(obs.one_plus_z);
        const Real lg2_nu_lo = fast_log2(xt::amin(nu_obs)()) + lg2_1pz;
        const Real lg2_nu_hi = fast_log2(xt::amax(nu_obs)()) + lg2_1pz;
        const Array lg2_dop_min_k = xt::amin(obs.lg2_doppler, {0, 1});
        const Array lg2_dop_max_k = xt::amax(obs.lg2_doppler, {0, 1});
        const Array nu_eval_min_k = xt::exp2(lg2_nu_lo - lg2_dop_max_k);
        const Array nu_eval_max_k = xt::exp2(lg2_nu_hi - lg2_dop_min_k);
        auto IC_ph = generate_IC_photons(syn_e, syn_ph, kn, coord, nu_eval_min_k, nu_eval_max_k);
        out_ssc = flux_func(obs, t_obs, nu_obs, IC_ph);
        has_ssc = true;
    }
}

struct Emission {
    XT fwd_sync, fwd_ssc, rvs_sync, rvs_ssc;
    bool has_fwd_ssc{false}, has_rvs{false}, has_rvs_ssc{false};
};

// mirror of PyModel::compute_emission (pybind/pymodel.h:922-961)
template <typename Func>
Emission compute_emission(RefModel const& m, Array const& t_obs, Array const& nu_obs, Func&& flux_func) {
    Emission e;
    Observer observer;
    if (!m.has_rvs) {
        auto [coord, fwd_shock] = solve_fwd_shock(m.jet, m.medium, t_obs, m.theta_w, m.theta_obs, m.z, m.phi_resol,
                                                  m.theta_resol, m.t_resol, m.axisymmetric, m.fwd_rad, m.rtol);
        observer.observe(coord, fwd_shock, m.lumi_dist, m.z);
        shock_emission(fwd_shock, coord, t_obs, nu_obs, observer, m.fwd_ssc, m.fwd_kn, e.fwd_sync, e.fwd_ssc,
                       e.has_fwd_ssc, flux_func);
    } else {
        auto [coord, fwd_shock, rvs_shock] =
            solve_shock_pair(m.jet, m.medium, t_obs, m.theta_w, m.theta_obs, m.z, m.phi_resol, m.theta_resol,
                             m.t_resol, m.axisymmetric, m.fwd_rad, m.rvs_rad, m.rtol);
        observer.observe(coord, fwd_shock, m.lumi_dist, m.z);
        shock_emission(fwd_shock, coord, t_obs, nu_obs, observer, m.fwd_ssc, m.fwd_kn, e.fwd_sync, e.fwd_ssc,
                       e.has_fwd_ssc, flux_func);
        shock_emission(rvs_shock, coord, t_obs, nu_obs, observer, m.rvs_ssc, m.rvs_kn, e.rvs_sync, e.rvs_ssc,
                       e.has_rvs_ssc, flux_func);
        e.has_rvs = true;
    }
    return e;
}

void store(double* out, size_t n, Emission const& e) {
    std::memset(out, 0, sizeof(double) * n * VAG_NCOMP);
    auto put = [&](int c, XT const& a) {
        size_t i = 0;
        for (auto it = a.begin(); it != a.end() && i < n; ++it, ++i) {
            out[c * n + i] = *it;
            out[VAG_C_TOTAL * n + i] += *it;
        }
    };
    put(VAG_C_FWD_SYNC, e.fwd_sync);
    if (e.has_fwd_ssc) put(VAG_C_FWD_SSC, e.fwd_ssc);
    if (e.has_rvs) put(VAG_C_RVS_SYNC, e.rvs_sync);
    if (e.has_rvs_ssc) put(VAG_C_RVS_SSC, e.rvs_ssc);
}

template <typename F>
void parallel_for(size_t n, int n_threads, F&& f) {
    if (n_threads <= 1 || n <= 1) {
        for (size_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
        pool.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n) break;
                f(i);
            }
        });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int vagref_hardware_threads() {
    return static_cast<int>(std::thread::hardware_concurrency());
}

// PyModel::flux_density_grid (pybind/pymodel.cpp:498-514) for a batch; out[n][5][n_nu][n_t]
__attribute__((visibility("default"))) int vagref_flux_density_grid(const vag_params* params, size_t n_models,
                                                                     const double* t, size_t n_t, const double* nu,
                                                                     size_t n_nu, double* out, int n_threads) {
    Array t_obs = Array::from_shape({n_t});
    Array nu_obs = Array::from_shape({n_nu});
    for (size_t i = 0; i < n_t; ++i) t_obs(i) = t[i] * unit::sec;
    for (size_t i = 0; i < n_nu; ++i) nu_obs(i) = nu[i] * unit::Hz;
    auto flux_func = [](Observer& obs, Array const& time, Array const& freq, auto& photons) -> XT {
        return obs.specific_flux(time, freq, photons) / unit::flux_den_cgs;
    };
    parallel_for(n_models, n_threads, [&](size_t i) {
        RefModel m = build(params[i]);
        Emission e = compute_emission(m, t_obs, nu_obs, flux_func);
        store(out + i * VAG_NCOMP * n_nu * n_t, n_nu * n_t, e);
    });
    return 0;
}

// PyModel::flux_density (pybind/pymodel.cpp:373-389) for a batch; out[n][5][n_pts]
__attribute__((visibility("default"))) int vagref_flux_density_series(const vag_params* params, size_t n_models,
                                                                       const double* t, const double* nu, size_t n,
                                                                       double* out, int n_threads) {
    Array t_obs = Array::from_shape({n});
    Array nu_obs = Array::from_shape({n});
    for (size_t i = 0; i < n; ++i) {
        t_obs(i) = t[i] * unit::sec;
        nu_obs(i) = nu[i] * unit::Hz;
    }
    auto flux_func = [](Observer& obs, Array const& time, Array const& freq, auto& photons) -> XT {
        return obs.specific_flux_series(time, freq, photons) / unit::flux_den_cgs;
    };
    parallel_for(n_models, n_threads, [&](size_t i) {
        RefModel m = build(params[i]);
        Emission e = compute_emission(m, t_obs, nu_obs, flux_func);
        store(out + i * VAG_NCOMP * n, n, e);
    });
    return 0;
}

// Fitter._evaluate + _chi2_sum for point data (VegasAfterglow/fitting/fitter.py:497-522) on top of
// the reference's flux_density; the 3-line chi2 formula is restated here because the Fitter class
// itself needs emcee/bilby, which are not installed.
__attribute__((visibility("default"))) int vagref_chi2_series(const vag_params* params, size_t n_models,
                                                               const double* t, const double* nu,
                                                               const double* lnF_obs, const double* sigma_ln,
                                                               const double* w, size_t n, double* chi2,
                                                               int n_threads) {
    std::vector<double> flux(n_models * VAG_NCOMP * n);
    vagref_flux_density_series(params, n_models, t, nu, n, flux.data(), n_threads);
    for (size_t m = 0; m < n_models; ++m) {
        const double* F = flux.data() + m * VAG_NCOMP * n;  // total
        double s = 0;
        for (size_t i = 0; i < n; ++i) {
            const double d = (lnF_obs[i] - std::log(std::max(F[i], 1e-300))) / sigma_ln[i];
            s += w[i] * d * d;
        }
        chi2[m] = std::isfinite(s) ? s : std::numeric_limits<double>::infinity();
    }
    return 0;
}

// Stage dump of one model: Coord + Shock tables + observer grids (the tables PyModel::details,
// pybind/pymodel.cpp:315-348, is built from, plus coord.t / theta_reps / injection_idx which it
// does not expose).  Call with all-NULL pointers to obtain *info first.
// Shock tables are dumped for the representative rows only: [7][n_reps][n_t].
// obs grids are dumped in full: [n_phi_eff][n_theta][n_t].
__attribute__((visibility("default"))) int vagref_details(const vag_params* p, double t_min, double t_max,
                                                           vag_grid_info* info, double* theta, double* phi,
                                                           int32_t* reps, double* t_rows, double* fwd_shock,
                                                           double* rvs_shock, int32_t* inj_idx, double* lg2_t,
                                                           double* lg2_doppler, double* lg2_geom) {
    RefModel m = build(*p);
    // PyModel::details uses logspace(t_min, t_max, 10): only min/max matter to auto_grid.
    Array t_obs = Array::from_shape({2});
    t_obs(0) = t_min * unit::sec;
    t_obs(1) = t_max * unit::sec;
    Coord coord;
    Shock fwd, rvs;
    if (!m.has_rvs) {
        auto res = solve_fwd_shock(m.jet, m.medium, t_obs, m.theta_w, m.theta_obs, m.z, m.phi_resol, m.theta_resol,
                                   m.t_resol, m.axisymmetric, m.fwd_rad, m.rtol);
        coord = std::move(res.first);
        fwd = std::move(res.second);
    } else {
        auto res = solve_shock_pair(m.jet, m.medium, t_obs, m.theta_w, m.theta_obs, m.z, m.phi_resol, m.theta_resol,
                                    m.t_resol, m.axisymmetric, m.fwd_rad, m.rvs_rad, m.rtol);
        coord = std::move(std::get<0>(res));
        fwd = std::move(std::get<1>(res));
        rvs = std::move(std::get<2>(res));
    }
    Observer obs;
    obs.observe(coord, fwd, m.lumi_dist, m.z);

    const size_t n_phi = coord.phi.size(), n_theta = coord.theta.size(), n_t = coord.t.shape()[2];
    const size_t n_reps = coord.theta_reps.size();
    if (info) {
        info->n_phi = (int)n_phi;
        info->n_theta = (int)n_theta;
        info->n_t = (int)n_t;
        info->n_reps = (int)n_reps;
        info->symmetry = (int)coord.symmetry;
        info->phi_mirrored = coord.phi_mirrored ? 1 : 0;
        info->n_phi_eff = (int)obs.lg2_t.shape()[0];
        info->status = 0;
    }
    if (theta) for (size_t j = 0; j < n_theta; ++j) theta[j] = coord.theta(j);
    if (phi) for (size_t i = 0; i < n_phi; ++i) phi[i] = coord.phi(i);
    if (reps) for (size_t r = 0; r < n_reps; ++r) reps[r] = (int32_t)coord.theta_reps[r];
    if (t_rows)
        for (size_t r = 0; r < n_reps; ++r)
            for (size_t k = 0; k < n_t; ++k) t_rows[r * n_t + k] = coord.t(0, coord.theta_reps[r], k);
    auto dump_shock = [&](Shock const& s, double* o) {
        MeshGrid3d const* arrs[7] = {&s.t_comv, &s.r, &s.theta, &s.Gamma, &s.Gamma_th, &s.B, &s.N_p};
        for (int a = 0; a < 7; ++a)
            for (size_t r = 0; r < n_reps; ++r)
                for (size_t k = 0; k < n_t; ++k)
                    o[(a * n_reps + r) * n_t + k] = (*arrs[a])(0, coord.theta_reps[r], k);
    };
    if (fwd_shock) dump_shock(fwd, fwd_shock);
    if (rvs_shock && m.has_rvs) dump_shock(rvs, rvs_shock);
    if (inj_idx && m.has_rvs)
        for (size_t r = 0; r < n_reps; ++r) inj_idx[r] = (int32_t)rvs.injection_idx(0, coord.theta_reps[r]);
    const size_t n_pe = obs.lg2_t.shape()[0];
    auto dump3 = [&](MeshGrid3d const& a, double* o) {
        for (size_t i = 0; i < n_pe; ++i)
            for (size_t j = 0; j < n_theta; ++j)
                for (size_t k = 0; k < n_t; ++k) o[(i * n_theta + j) * n_t + k] = a(i, j, k);
    };
    if (lg2_t) dump3(obs.lg2_t, lg2_t);
    if (lg2_doppler) dump3(obs.lg2_doppler, lg2_doppler);
    if (lg2_geom) dump3(obs.lg2_geom_factor, lg2_geom);
    return 0;
}

}  // extern "C"
