/*
 * vag.h -- C ABI of the B200-native VegasAfterglow model-evaluation path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * torch / pybind / xtensor types.  Every entry point names the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Units at this boundary are the reference's *Python-facing* units (pybind/pymodel.cpp:498-514,
 * pybind/pymodel.h:246): seconds, Hz, cm, erg, cm^-3 in; erg cm^-2 s^-1 Hz^-1 out.
 *
 * All compute entry points run on the GPU (sm_100a kernels).  There is no CPU fallback: when no
 * CUDA device is usable they return VAG_ERR_CUDA and vag_last_error() says why.
 */
#ifndef VAG_H_
#define VAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------- */
/* status codes                                                                                 */
/* ------------------------------------------------------------------------------------------- */
enum {
    VAG_OK = 0,
    VAG_ERR_INVALID = 1, /* argument validation failed: mirrors AFTERGLOW_REQUIRE -> ValueError
                            (pybind/error_handling.h:31-69) */
    VAG_ERR_CUDA = 2,    /* CUDA runtime error / no device                                       */
    VAG_ERR_UNSUPPORTED = 3, /* a switch of the reference that this path does not implement (Python-callable
                                Ejecta / Medium profiles cannot be expressed in vag_params at all)  */
    VAG_ERR_CAPACITY = 4     /* a per-model grid exceeded the compiled capacity                  */
};

/* jet / medium enumerations: the typed variants of JetVariant / MediumVariant
 * (src/environment/jet.h:272, src/environment/medium.h:144). */
enum {
    VAG_JET_TOPHAT = 0,
    VAG_JET_GAUSSIAN = 1,
    VAG_JET_POWERLAW = 2,
    /* closed-form members of the reference's Ejecta family (pybind/pymodel.cpp:97-146,
     * src/environment/jet.h:403-470) as enumerated device profiles */
    VAG_JET_TWO_COMPONENT = 3, /* TwoComponentJet(theta_c, E_iso, Gamma0, theta_w, E_iso_w, Gamma0_w)        */
    VAG_JET_STEP_POWERLAW = 4, /* StepPowerLawJet(theta_c, E_iso, Gamma0, E_iso_w, Gamma0_w, k_e, k_g)       */
    VAG_JET_POWERLAW_WING = 5  /* PowerLawWing(theta_c, E_iso_w, Gamma0_w, k_e, k_g)                         */
};
enum { VAG_MEDIUM_ISM = 0, VAG_MEDIUM_WIND = 1 };

/* per-model status bits written by the kernels (the reference prints a
 * warning on stderr and leaves the row at its initial values; we mirror that and set a bit). */
enum {
    VAG_ST_ODE_STEP_CAP = 1,  /* forward-shock.tpp:196-200 / reverse-shock.tpp:555-559           */
    VAG_ST_ODE_STALLED = 2,   /* reverse-shock.tpp:560-566                                       */
    VAG_ST_ODE_FAIL500 = 4,   /* boost max_step_checker.hpp:99-106 (reference throws)            */
    VAG_ST_GRID_NONFINITE = 8, /* grid-refinement.h:633-635                                      */
    VAG_ST_CAPACITY = 16,     /* grid / SSC lattice larger than compiled capacity; model output is NaN */
    VAG_ST_IC_BAND = 32       /* an SSC query left the clamped output band (inverse-compton.h:622-635:
                                 the reference would rebuild that cell's full-range spectrum)      */
};

/* Radiation(eps_e, eps_B, p, xi_e=1, ssc=False, kn=False): pybind/pymodel.h:303-313 */
typedef struct vag_radiation {
    double eps_e, eps_B, p, xi_e;
    int32_t ssc, kn;
} vag_radiation;

/* One parameter set = everything Model.__init__ receives (pybind/pybind.cpp:384-422,
 * pybind/pymodel.h:613-649) for the typed jet/medium variants. */
typedef struct vag_params { /* 320 bytes, mirrored by vegasafterglow_b200/abi.py PARAMS_DTYPE */
    /* jet: TophatJet/GaussianJet/PowerLawJet(theta_c, E_iso, Gamma0[, k_e, k_g], spreading,
     * duration)  pybind/pymodel.cpp:47-95; two-component / step-power-law / power-law-wing :97-146 */
    int32_t jet_type;
    int32_t spreading; /* spreading=True of the jet factories: lateral spreading (forward-shock models) and
                          Symmetry::structured lattices; with axisymmetric = 0 one ODE row per (phi, theta) cell */
    double theta_c, E_iso, Gamma0, k_e, k_g, duration;
    double theta_w, E_iso_w, Gamma0_w; /* wing of the two-component / step-power-law / power-law-wing jets */
    double sigma0;                     /* ejecta magnetisation (constant; the reference expresses it through
                                          Ejecta(sigma0=...)); 0 = unmagnetised */
    /* medium: ISM(n_ism) / Wind(A_star, n_ism=0, n0=inf, k_m=2)  pybind/pymodel.cpp:148-186 */
    int32_t medium_type;
    int32_t pad0_;
    double n_ism, A_star, n0;
    /* Observer(lumi_dist[cm], z, theta_obs, phi_obs=0)  pybind/pymodel.h:190-204 */
    double lumi_dist, z, theta_obs, phi_obs;
    /* radiation */
    vag_radiation fwd;
    vag_radiation rvs;
    int32_t has_rvs;
    int32_t axisymmetric;       /* default 1 */
    int32_t radiative_fireball; /* default 1 */
    int32_t pad1_;
    /* resolutions=(phi, theta, t); a value <= 0 selects the reference default
     * (0.06,0.15,6) forward-only or (0.06,0.2,10) with a reverse shock
     * (src/config/simulation-defaults.h:71-83, pybind/pymodel.h:633-640) */
    double phi_resol, theta_resol, t_resol;
    double rtol; /* <= 0 selects defaults::solver::dynamics_rtol = 1e-6 */
    /* magnetar=Magnetar(L0 [erg/s], t0 [s], q) of the jet factories (pybind/pymodel.h:34-58): energy injection
     * L0 (1 + t/t0)^-q for theta <= theta_c (src/environment/jet.h:518-528).  As in the reference the jet then
     * takes the generic-Ejecta code path (pybind/pymodel.cpp:53-59). */
    int32_t has_magnetar;
    int32_t pad2_;
    double magnetar_L0, magnetar_t0, magnetar_q;
    /* Wind(..., k_m): density slope of the wind, rho = A / (r0^k + r^k) + rho_ism.  2 (or <= 0) selects the typed
     * Wind of the reference; any other value takes its generic-Medium path (pybind/pymodel.cpp:169-185:
     * numeric enclosed mass / thermal energy, CGS profile behind convert_unit_medium). */
    double wind_k_m;
} vag_params;

/* Fill *p with the reference defaults (Tophat/ISM values are NOT set, only the switches). */
void vag_params_default(vag_params* p);

/* Validate one parameter set exactly like the reference constructors do
 * (pybind/pymodel.cpp:47-186, pybind/pymodel.h:190-204,303-313,613-649).
 * Returns VAG_OK or VAG_ERR_INVALID / VAG_ERR_UNSUPPORTED; message via vag_last_error(). */
int vag_params_validate(const vag_params* p);

/* ------------------------------------------------------------------------------------------- */
/* context                                                                                      */
/* ------------------------------------------------------------------------------------------- */
typedef struct vag_context vag_context;

/* Create a context bound to CUDA device `device` (owns a stream and growable workspaces). */
int vag_create(int device, vag_context** out);
void vag_destroy(vag_context* ctx);
const char* vag_last_error(void);
const char* vag_version(void);

/* Output component order of every flux entry point: PyFlux (pybind/pymodel.h:394-400). */
enum { VAG_C_TOTAL = 0, VAG_C_FWD_SYNC = 1, VAG_C_FWD_SSC = 2, VAG_C_RVS_SYNC = 3, VAG_C_RVS_SSC = 4, VAG_NCOMP = 5 };

/* ------------------------------------------------------------------------------------------- */
/* batched model evaluation, HOST buffers (copies inside)                                       */
/* ------------------------------------------------------------------------------------------- */

/* Replaces PyModel::flux_density_grid (pybind/pymodel.cpp:498-514) for n_models parameter sets
 * sharing one (t, nu) request.  t ascending [s], nu [Hz].
 * out[n_models][VAG_NCOMP][n_nu][n_t]  (components a model does not have are 0; total = sum,
 * pybind/pymodel.cpp:350-364).  status[n_models] receives VAG_ST_* bits (may be NULL). */
int vag_flux_density_grid(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, size_t n_t,
                          const double* nu, size_t n_nu, double* out, int32_t* status);

/* Replaces PyModel::flux_density (pybind/pymodel.cpp:373-389): series of (t[i], nu[i]) points,
 * t ascending.  out[n_models][VAG_NCOMP][n]. */
int vag_flux_density_series(vag_context* ctx, const vag_params* params, size_t n_models, const double* t,
                            const double* nu, size_t n, double* out, int32_t* status);

/* Replaces PyModel::flux (pybind/pymodel.cpp:391-410) = Observer::flux (src/core/observer.h:555-567):
 * flux integrated over [nu_min, nu_max] Hz with Boole weights on num_nu log-spaced frequencies
 * (src/core/quadrature.h:138-191).  out[n_models][VAG_NCOMP][n_t] in erg cm^-2 s^-1. */
int vag_flux_band(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, size_t n_t,
                  double nu_min, double nu_max, size_t num_nu, double* out, int32_t* status);

/* Replaces Fitter._evaluate + _chi2_sum for point data (VegasAfterglow/fitting/fitter.py:497-522):
 * chi2[m] = sum_i w[i] * ((lnF_obs[i] - ln max(F_model[i], 1e-300)) / sigma_ln[i])^2
 * over the series (t[i], nu[i]); non-finite chi2 is returned as +inf (samplers.py:63-70 maps it
 * to logL = -inf).  t ascending. */
int vag_chi2_series(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, const double* nu,
                    const double* lnF_obs, const double* sigma_ln, const double* w, size_t n, double* chi2,
                    int32_t* status);

/* One band-integrated data set of the likelihood (VegasAfterglow/fitting/fitter.py BandObs, :525-531): observed
 * band fluxes [erg cm^-2 s^-1] at ascending epochs t[n], compared with Model.flux(t, nu_min, nu_max, num_nu). */
typedef struct vag_band_obs {
    const double* t;        /* [n] seconds, ascending                                   */
    const double* lnF_obs;  /* [n] ln(observed band flux)                               */
    const double* sigma_ln; /* [n] err / flux                                           */
    const double* w;        /* [n] weights                                              */
    size_t n;
    double nu_min, nu_max;  /* Hz                                                        */
    size_t num_nu;          /* frequency nodes of the Boole quadrature (>= 2)            */
} vag_band_obs;

/* Replaces Fitter._evaluate (fitter.py:503-533) for a batch: chi2[i] = sum over the point data (as vag_chi2_series;
 * n_points may be 0) + sum over every band data set of w ((lnF_obs - ln max(F_band, 1e-300)) / sigma_ln)^2.
 * A non-finite sum, or a model whose ODE hit Boost's 500-rejection limit (VAG_ST_ODE_FAIL500 -- an exception in the
 * reference, logL = -inf in its samplers, samplers.py:63-70), gives chi2 = +inf.  status[i] = OR over all terms. */
int vag_chi2(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, const double* nu,
             const double* lnF_obs, const double* sigma_ln, const double* w, size_t n_points, const vag_band_obs* bands,
             size_t n_bands, double* chi2, int32_t* status);

/* ok[i] = 1 where vag_params_validate accepts params[i], else 0 (no error is raised for rejected sets). */
int vag_params_validate_batch(const vag_params* params, size_t n, int32_t* ok);

/* ------------------------------------------------------------------------------------------- */
/* batched model evaluation, DEVICE buffers (no copies; asynchronous on `stream`)               */
/* ------------------------------------------------------------------------------------------- */
/* Same contracts as above; every pointer is a device pointer on ctx's device; `stream` is a
 * cudaStream_t passed as void* (NULL = the context's own stream).  The calls are stream-ordered;
 * they synchronise the stream once internally (to size the ragged workspace after the grid
 * kernel) and return with the remaining kernels enqueued: call vag_synchronize / sync the stream
 * before reading the outputs.  t must be ascending (not checked on the device path). */
int vag_flux_density_grid_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                              size_t n_t, const double* d_nu, size_t n_nu, double* d_out, int32_t* d_status,
                              void* stream);
int vag_flux_density_series_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                                const double* d_nu, size_t n, double* d_out, int32_t* d_status, void* stream);
int vag_chi2_series_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                        const double* d_nu, const double* d_lnF_obs, const double* d_sigma_ln, const double* d_w,
                        size_t n, double* d_chi2, int32_t* d_status, void* stream);
int vag_synchronize(vag_context* ctx);
/* The *_dev entry points cannot see the parameters on the host, so they use the context's grid
 * capacities (theta nodes / phi nodes per model; defaults 384 / 128).  The host-buffer entry
 * points derive tight capacities from the parameters themselves.  A model that needs more nodes
 * than the capacity gets VAG_ST_CAPACITY and NaN output. */
int vag_set_capacity(vag_context* ctx, int cap_theta, int cap_phi);

/* ------------------------------------------------------------------------------------------- */
/* introspection (tests, profiling): stage tables of ONE model, the analogue of                 */
/* PyModel::details (pybind/pymodel.cpp:315-348)                                                */
/* ------------------------------------------------------------------------------------------- */
typedef struct vag_grid_info {
    int32_t n_phi, n_theta, n_t;   /* Coord shape (src/core/mesh.h:55-83)                        */
    int32_t n_reps;                /* coord.theta_reps.size()                                    */
    int32_t symmetry;              /* Symmetry enum value (src/core/mesh.h:49-54)                */
    int32_t phi_mirrored;
    int32_t n_phi_eff;             /* Observer::eff_phi_grid (src/core/observer.cpp:218-222)     */
    int32_t status;
} vag_grid_info;

/* Runs grid + dynamics (+radiation) for one model and the observation window [t_min, t_max] s.
 * Any output pointer may be NULL.  Sizes: theta[n_theta], phi[n_phi], reps[n_reps],
 * t_rows[n_reps][n_t] (engine-frame lattice of each representative row, code units),
 * fwd_shock / rvs_shock [7][n_reps][n_t] in the order t_comv, r, theta, Gamma, Gamma_th, B, N_p
 * (code units, src/dynamics/shock.h:33-39), inj_idx[n_reps] (reverse shock injection_idx).
 * Call once with all NULL to get *info, then allocate. */
int vag_details(vag_context* ctx, const vag_params* p, double t_min, double t_max, vag_grid_info* info,
                double* theta, double* phi, int32_t* reps, double* t_rows, double* fwd_shock, double* rvs_shock,
                int32_t* inj_idx);

/* Photon tables of one model (same front end as vag_details): for each shock [6][n_reps][n_t] in the order
 * log2 nu_m, log2 nu_c, log2 nu_a, log2 nu_M, log2 I_nu_max (code units) and 1/nu_M -- what
 * save_photon_details exports (pybind/pymodel.cpp:263-291).  rvs may be NULL; it is untouched without a
 * reverse shock.  For a shock with ssc=True these are the photons after inverse-Compton cooling. */
int vag_details_photons(vag_context* ctx, const vag_params* p, double t_min, double t_max, double* fwd, double* rvs);

/* Electron and inverse-Compton bookkeeping of the shocks with ssc=True, after KN_cooling / Thomson_cooling
 * (pybind/pymodel.cpp:234-291, 303): for each shock [VAG_IC_DETAIL_PLANES][n_reps][n_t] in the order gamma_m, gamma_c,
 * gamma_a, gamma_M, gamma_m_hat, gamma_c_hat, Y_T.  Planes of a shock without ssc are zero; rvs may be NULL. */
#define VAG_IC_DETAIL_PLANES 7
int vag_details_ic(vag_context* ctx, const vag_params* p, double t_min, double t_max, double* fwd, double* rvs);

/* Output transfer policy of the HOST-buffer flux entry points.
 * VAG_OUT_DENSE (default): every one of the VAG_NCOMP planes of `out` is written (absent components 0).
 * VAG_OUT_PRESENT: planes of components that NO model of the batch has (reverse shock: has_rvs;
 * SSC: fwd.ssc / rvs.ssc) are neither transferred nor written -- the caller's buffer is left
 * untouched there.  This mirrors the reference's FluxDict, whose absent components are empty
 * arrays that are never materialised (pybind/pymodel.h:361-383), and saves 3/5 of the device-to-host
 * traffic of a forward-shock synchrotron batch.  VAG_C_TOTAL and VAG_C_FWD_SYNC are always present. */
enum { VAG_OUT_DENSE = 0, VAG_OUT_PRESENT = 1, VAG_OUT_PRESENT_ALIAS_TOTAL = 2 };
int vag_set_output_mode(vag_context* ctx, int mode);
/* VAG_OUT_PRESENT_ALIAS_TOTAL: as VAG_OUT_PRESENT, and when exactly ONE emission component exists in the whole batch
 * (e.g. forward-shock synchrotron only) the `total` plane -- which is then that component bit for bit -- is not
 * transferred either: vag_last_total_alias() returns the component index (VAG_C_*) whose plane holds the total of the
 * most recent host-buffer call, or -1 when `total` was written.  (The pybind mirror hands the same array out twice.) */
int vag_last_total_alias(vag_context* ctx);

/* Series requests (vag_flux_density_series, vag_chi2*) whose points share at most 8 distinct frequencies -- multi-band
 * light curves -- are evaluated "banded": the boundary luminosities of every lattice node are staged once per band and
 * a point interpolates its band's column, instead of every point evaluating both bracketing spectra itself (the
 * evaluation points are those of Observer::specific_flux_series, src/core/observer.h:494-520; results agree to
 * rounding).  mode 0 (default): chosen per request by cost; 1: never; 2: whenever the request has <= 8 frequencies. */
int vag_set_series_mode(vag_context* ctx, int mode);

/* per-stage device time of the most recent batched call on this context, milliseconds:
 * [0]=grid (K0) [1]=dynamics (K1) [2]=radiation (K2) [3]=EATS flux (K3) [4]=likelihood (K4)
 * Only filled when vag_set_profiling(ctx, 1) was called (adds event records + one sync). */
int vag_set_profiling(vag_context* ctx, int enable);
int vag_last_stage_ms(vag_context* ctx, float ms[8]);
/* Work counters of the most recent profiled pass (the denominators of bench.py's roofline entries): [0] forward-only ODE
 * rows, [1] forward+reverse ODE rows, [2] shock-table cells x shocks, [3] EATS cells (shocks x n_phi_eff x n_theta x n_t),
 * [4] EATS rows, [5] theta-quadrature attempts, [6] phi-quadrature attempts x n_theta, [7] sum of n_theta. */
int vag_last_work(vag_context* ctx, double work[8]);
/* number of kernel launches issued by the most recent batched call */
int vag_last_launch_count(vag_context* ctx);
/* dependent-free DFMA throughput of the device in TFLOP/s (FP64 roofline denominator) */
int vag_measure_fp64_peak(vag_context* ctx, double* tflops);
/* Test hook: lower the shock ODE's accepted-step limit (defaults::solver::max_ode_steps = 100000 -> VAG_ST_ODE_STEP_CAP,
 * forward-shock.tpp:196-200) and / or the consecutive-rejection limit (Boost's 500 -> VAG_ST_ODE_FAIL500) so that the failure
 * semantics can be exercised; 0 restores the reference value. */
int vag_debug_set_ode_limits(vag_context* ctx, int max_steps, int max_fails);
/* Device self-test of the libm-exact functions the grid kernel evaluates (csrc/vag_libm.cuh):
 * out[i] = fn(x[i]) with fn 0 exp, 1 exp2, 2 log, 3 log2, 4 log10, 5 pow(x[i], y[i]), 6 sin, 7 cos; host buffers.
 * The results must equal the host libm's bit for bit (the reference's theta / phi grids depend on it,
 * src/core/grid-refinement.h:137-189); y is NULL except for pow. */
int vag_selftest_libm(vag_context* ctx, int fn, const double* x, const double* y, double* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* VAG_H_ */
