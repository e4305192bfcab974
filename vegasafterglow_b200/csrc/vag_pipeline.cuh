// vag_pipeline.cuh -- batch workspace layout in HBM and the body of every kernel of the path.
//
// Pipeline for a batch of parameter sets sharing one observation request
// (the batched form of PyModel::compute_emission, pybind/pymodel.h:922-961):
//   K0  grid        thread / model          -> GridHeader, theta[], phi[], reps[], t_dec[]
//   K0b scan        one CTA                 -> row / cell offsets of the ragged batch, totals
//   K0c rowmap      thread / model          -> row -> (model, rep) map, theta_j -> rep map
//   K1  dynamics    thread / unique row     -> time lattice + shock tables (SoA, [row][k])
//   K2  radiation   CTA / unique row        -> photon coefficients (SoA planes, coalesced in k)
//   K3  EATS        CTA / (model, split, shock) -> flux [comp][nu][t] (vag_observer.cuh)
//   K4  chi2        warp / model            -> chi^2 of the series against data
// Each body is an HD function so the CPU test-suite can execute the identical code sequentially.
#pragma once

#include "vag_grid.cuh"
#include "vag_observer.cuh"
#include "vag_radiation.cuh"
#include "vag_shock.cuh"

namespace vag {

enum { TOT_ROWS = 0, TOT_MAX_NT, TOT_MAX_NTHETA, TOT_MAX_EROWS, TOT_STATUS_OR, TOT_N };

struct BatchWs {
    int n_models;
    int cap_theta, cap_phi;
    size_t work_per_model;  // doubles
    const vag_params* params;
    ModelCfg* cfg;
    GridHeader* hdr;
    double* theta;   // [n][cap_theta]
    double* phi;     // [n][cap_phi]
    double* t_dec;   // [n][cap_theta]
    double* work;    // [n][work_per_model]
    int* reps;       // [n][cap_theta]
    int* rep_of;     // [n][cap_theta]
    int* row_off;    // [n+1]
    long long* cell_off;  // [n+1]
    int* totals;     // [TOT_N]
    int* status;     // [n]
    // per unique row
    int* row_model;
    int* row_rep;
    int* inj_idx;
    // per unique cell
    long long n_cells;
    double* t_rows;
    double* fwd[6];  // t_comv, r, Gamma, Gamma_th, B, N_p
    double* rvs[6];
    double* coef_fwd;  // [PH_NCOEF][n_cells]
    double* coef_rvs;
};

// ---- K0 ---------------------------------------------------------------------------------------
template <class Par>
VAG_HD void k0_grid_body(const Par& par, const BatchWs& w, int mi, double t_obs_min, double t_obs_max) {
    ModelCfg cfg = make_cfg(w.params[mi]);
    w.cfg[mi] = cfg;
    GridSlab s;
    s.theta = w.theta + (size_t)mi * w.cap_theta;
    s.phi = w.phi + (size_t)mi * w.cap_phi;
    s.reps = w.reps + (size_t)mi * w.cap_theta;
    s.t_dec = w.t_dec + (size_t)mi * w.cap_theta;
    s.work = w.work + (size_t)mi * w.work_per_model;
    s.cap_theta = w.cap_theta;
    s.cap_phi = w.cap_phi;
    GridHeader h;
    build_grid(par, cfg, t_obs_min * unit::sec, t_obs_max * unit::sec, h, s);
    w.hdr[mi] = h;
    w.status[mi] = h.status;
}

// ---- K0b: exclusive scan over models (single CTA; `tid`-strided two-pass is overkill for the
// few thousand models of a batch, so thread 0 of the CTA walks the headers) -------------------------
VAG_HD void k0b_scan_body(const BatchWs& w) {
    int rows = 0, max_nt = 0, max_nth = 0, max_erows = 0, st = 0;
    long long cells = 0;
    for (int mi = 0; mi < w.n_models; ++mi) {
        const GridHeader& h = w.hdr[mi];
        w.row_off[mi] = rows;
        w.cell_off[mi] = cells;
        rows += h.n_reps;
        cells += (long long)h.n_reps * h.n_t;
        max_nt = imax(max_nt, h.n_t);
        max_nth = imax(max_nth, h.n_theta);
        max_erows = imax(max_erows, h.n_theta * h.n_phi_eff);
        st |= h.status;
    }
    w.row_off[w.n_models] = rows;
    w.cell_off[w.n_models] = cells;
    w.totals[TOT_ROWS] = rows;
    w.totals[TOT_MAX_NT] = max_nt;
    w.totals[TOT_MAX_NTHETA] = max_nth;
    w.totals[TOT_MAX_EROWS] = max_erows;
    w.totals[TOT_STATUS_OR] = st;
}

// ---- K0c --------------------------------------------------------------------------------------
VAG_HD void k0c_rowmap_body(const BatchWs& w, int mi) {
    const GridHeader& h = w.hdr[mi];
    const int* reps = w.reps + (size_t)mi * w.cap_theta;
    int* rep_of = w.rep_of + (size_t)mi * w.cap_theta;
    const int ro = w.row_off[mi];
    for (int r = 0; r < h.n_reps; ++r) {
        w.row_model[ro + r] = mi;
        w.row_rep[ro + r] = r;
        const int j0 = reps[r];
        const int j1 = (r + 1 < h.n_reps) ? reps[r + 1] : h.n_theta;
        for (int j = j0; j < j1; ++j) rep_of[j] = r;
    }
}

VAG_HD ShockRow shock_row(double* const* planes, long long off) {
    ShockRow s;
    s.t_comv = planes[0] + off;
    s.r = planes[1] + off;
    s.Gamma = planes[2] + off;
    s.Gamma_th = planes[3] + off;
    s.B = planes[4] + off;
    s.N_p = planes[5] + off;
    return s;
}

// ---- K1 ---------------------------------------------------------------------------------------
VAG_HD void k1_dynamics_body(const BatchWs& w, int row) {
    const int mi = w.row_model[row];
    const int r = w.row_rep[row];
    const GridHeader& h = w.hdr[mi];
    const ModelCfg& cfg = w.cfg[mi];
    const long long off = w.cell_off[mi] + (long long)r * h.n_t;
    double* t_row = w.t_rows + off;
    const double t_dec = w.t_dec[(size_t)mi * w.cap_theta + r];
    const double theta = w.theta[(size_t)mi * w.cap_theta + w.reps[(size_t)mi * w.cap_theta + r]];
    build_row_lattice(h, t_dec, cfg.T0, t_row);
    int st = 0;
    bool finite = true;
    for (int k = 0; k < h.n_t; ++k) finite = finite && isfinite(t_row[k]);
    if (!finite) st |= VAG_ST_GRID_NONFINITE;
    const ShockRow sf = shock_row(w.fwd, off);
    if (cfg.has_rvs) {
        const ShockRow sr = shock_row(w.rvs, off);
        int inj = h.n_t;
        st |= solve_pair_row(cfg, theta, t_dec, t_row, h.n_t, sf, sr, &inj);
        w.inj_idx[row] = inj;
    } else {
        st |= solve_fwd_row(cfg, theta, t_dec, t_row, h.n_t, sf);
        w.inj_idx[row] = h.n_t;
    }
    if (st) {
#if defined(__CUDA_ARCH__)
        atomicOr(&w.status[mi], st);
#else
        w.status[mi] |= st;
#endif
    }
}

// ---- K2 ---------------------------------------------------------------------------------------
// one (row, k) cell of one shock (which = 0 forward, 1 reverse)
VAG_HD void k2_radiation_cell(const BatchWs& w, int row, int k, int which) {
    const int mi = w.row_model[row];
    const GridHeader& h = w.hdr[mi];
    const ModelCfg& cfg = w.cfg[mi];
    const long long off = w.cell_off[mi] + (long long)w.row_rep[row] * h.n_t;
    double* const* pl = which ? w.rvs : w.fwd;
    const RadCfg& rad = which ? cfg.rvs : cfg.fwd;
    auto load = [&](int kk) {
        CellShock c;
        c.t_comv = pl[0][off + kk];
        c.r = pl[1][off + kk];
        c.Gamma_th = pl[3][off + kk];
        c.B = pl[4][off + kk];
        c.N_p = pl[5][off + kk];
        return c;
    };
    const CellShock cs = load(k);
    const int inj = which ? w.inj_idx[row] : h.n_t;
    const bool relic = k >= inj;
    const CellShock ics = relic ? load(inj - 1) : cs;
    double coef[PH_NCOEF];
    radiation_cell(rad, cs, relic, ics, coef);
    double* out = (which ? w.coef_rvs : w.coef_fwd) + off + k;
#pragma unroll
    for (int c = 0; c < PH_NCOEF; ++c) out[(long long)c * w.n_cells] = coef[c];
}

VAG_HD EatsModel make_eats_model(const BatchWs& w, int mi, int which) {
    const GridHeader& h = w.hdr[mi];
    const ModelCfg& cfg = w.cfg[mi];
    const long long off = w.cell_off[mi];
    EatsModel M;
    M.h = &w.hdr[mi];
    M.theta = w.theta + (size_t)mi * w.cap_theta;
    M.phi = w.phi + (size_t)mi * w.cap_phi;
    M.rep_of = w.rep_of + (size_t)mi * w.cap_theta;
    M.t_rows = w.t_rows + off;
    M.r = w.fwd[1] + off;      // the pair solver gives both shocks identical kinematics
    M.Gamma = w.fwd[2] + off;  // (pybind/pymodel.h:943-950): one EAT geometry serves both
    M.coef = (which ? w.coef_rvs : w.coef_fwd) + off;
    M.coef_stride = (long)w.n_cells;
    photon_p_consts(which ? cfg.rvs.p : cfg.fwd.p, M.smooth_thick, M.log2_x_far);
    M.one_plus_z = 1 + cfg.z;
    M.lumi_dist = cfg.lumi_dist;
    M.theta_v = cfg.theta_v;
    (void)h;
    return M;
}

// Final scaling of Observer::specific_flux (observer.h:442) and the unit conversion of
// PyModel::flux_density_grid (pybind/pymodel.cpp:507)
VAG_HD double flux_scale(const EatsModel& M, double F) {
    return F * (M.one_plus_z / (M.lumi_dist * M.lumi_dist)) / unit::flux_den_cgs;
}

// ---- K4: chi^2 of one model (VegasAfterglow/fitting/fitter.py:497-501) ---------------------------
VAG_HD double chi2_term(double lnF_obs, double F_model, double sigma_ln, double wgt) {
    const double d = (lnF_obs - log(vmax(F_model, 1e-300))) / sigma_ln;
    return wgt * d * d;
}

}  // namespace vag
