// vag_common.cuh -- shared scalar definitions of the B200 model-evaluation path.
//
// Everything here is FP64 (the reference computes in `using Real = double`, src/util/macros.h:24).
// The code-unit system and physical constants restate src/util/macros.h:43-107 with the same
// expression order so the compile-time values are bit-identical to the reference's.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define VAG_HD __host__ __device__ __forceinline__
#define VAG_HD_NOINLINE __host__ __device__ __noinline__
#else
#define VAG_HD inline
#define VAG_HD_NOINLINE inline
#endif

namespace vag {

using std::isfinite;
using std::isinf;
using std::isnan;

// ---- unit system: src/util/macros.h:43-77 ---------------------------------------------------
namespace unit {
constexpr double len = 1.5e13;
constexpr double cm = 1 / len;
constexpr double sec = 3e10 / len;
constexpr double cm2 = cm * cm;
constexpr double cm3 = cm * cm * cm;
constexpr double g = 1 / 2e33;
constexpr double Hz = 1 / sec;
constexpr double erg = g * cm * cm / sec / sec;
constexpr double flux_cgs = erg / cm2 / sec;
constexpr double flux_den_cgs = erg / cm2 / sec / Hz;
}  // namespace unit

// ---- constants: src/util/macros.h:86-107 ----------------------------------------------------
namespace con {
constexpr double c = 1;
constexpr double c2 = c * c;
constexpr double mp = 1.67e-24 * unit::g;
constexpr double me = mp / 1836;
constexpr double e = 4.8e-10 / 4.472136e16 / 5.809475e19 / unit::sec;
constexpr double e2 = e * e;
constexpr double e3 = e2 * e;
constexpr double pi = 3.14159265358979323846;
constexpr double sigmaT = 6.65e-25 * unit::cm * unit::cm;
constexpr double Gamma_cut = 1.0 + 1e-6;        // src/config/simulation-defaults.h:41
constexpr double gamma_therm_cut = 1.0 + 1e-6;  // :47
constexpr double sigma_cut = 1e-6;              // :45
constexpr double sqrt3 = 1.732050807568877293527446341505872367;
constexpr double ln2 = 0.693147180559945309417232121458176568;
constexpr double log2e = 1.442695040888963407359924681001892137;
}  // namespace con

// ---- numeric defaults: src/config/simulation-defaults.h ------------------------------------
namespace dflt {
constexpr double phi_resolution = 0.06;
constexpr double theta_resolution = 0.15;
constexpr double time_resolution = 6.0;
constexpr double rvs_theta_resolution = 0.2;
constexpr double rvs_time_resolution = 10.0;
constexpr int min_theta_points = 36;
constexpr double theta_min = 1e-6;
constexpr double ode_rtol = 1e-6;
constexpr double dynamics_rtol = 1e-6;
constexpr double magnetized_rtol_factor = 0.1;
constexpr double binary_search_eps = 1e-9;
constexpr int max_ode_steps = 100000;
constexpr int theta_samples = 200;
}  // namespace dflt

constexpr double kInf = __builtin_huge_val();

// std::min / std::max / std::clamp semantics (NaN behaviour included: the first argument wins
// when the comparison is false).
VAG_HD double vmin(double a, double b) { return (b < a) ? b : a; }
VAG_HD double vmax(double a, double b) { return (a < b) ? b : a; }
VAG_HD double vclamp(double v, double lo, double hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
VAG_HD int imin(int a, int b) { return a < b ? a : b; }
VAG_HD int imax(int a, int b) { return a > b ? a : b; }

// src/util/fast-math.h with AFTERGLOW_FAST_MATH off (the default build, CMakeLists.txt:39):
// fast_log2/exp2/exp/log are libm calls and fast_pow(a,b) = exp2(b*log2(a)) (fast-math.h:147-149).
VAG_HD double fast_log2(double x) { return log2(x); }
VAG_HD double fast_exp2(double x) { return exp2(x); }
VAG_HD double fast_exp(double x) { return exp(x); }
VAG_HD double fast_pow(double a, double b) { return exp2(b * log2(a)); }
// src/core/physics.h:36-40, :59-61
VAG_HD double gamma_to_beta(double gamma) { return sqrt((gamma - 1) * (gamma + 1)) / gamma; }
VAG_HD double adiabatic_idx(double gamma) { return 4.0 / 3.0 + 1 / (3 * gamma); }

VAG_HD bool vfinite(double x) { return isfinite(x); }

// Branch-free FP64 division / square root for the ODE right-hand sides, where the compiler's IEEE
// sequences (12-14 instructions each, every one ending in a range check and a branch to a slow-path
// call) chop the dependent chain into blocks the scheduler cannot overlap.  Reciprocal / rsqrt seed
// from the MUFU unit, two Newton steps and one residual correction: faithfully rounded (<= 1 ulp).
// ONLY for operands known to be finite, normal and (divisor / radicand) positive -- zero radicands are
// handled; callers keep the IEEE operators wherever a zero or infinite operand carries meaning.
#if defined(__CUDA_ARCH__)
// MUFU.RCP64H / RSQ64H seeds carry >= 20 good bits; ONE Newton step squares that to >= 40, and the residual correction
// of the quotient / root squares it again: the result is faithful (<= 1 ulp) with two fewer dependent FMAs per
// operation than the two-step form (the ODE right-hand side is one dependent chain of ~12 such operations).
VAG_HD double vdiv(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
VAG_HD double vsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    const double s = x * y;
    const double root = fma(0.5 * y, fma(-s, s, x), s);
    return (x == 0.0) ? 0.0 : root;
}
#else
VAG_HD double vdiv(double a, double b) { return a / b; }
VAG_HD double vsqrt(double x) { return sqrt(x); }
#endif
// Pins a value as computed HERE: the compiler otherwise sinks an expensive operand of a select into a conditional
// block, and the branch ends the basic block the caller wants in one piece (device only; no code is emitted).
#if defined(__CUDA_ARCH__)
#define VAG_KEEP(x) asm volatile("" : "+d"(x))
#else
#define VAG_KEEP(x) ((void)0)
#endif
VAG_HD double adiabatic_idx_fast(double gamma) { return 4.0 / 3.0 + vdiv(1.0, 3 * gamma); }  // gamma >= 1

}  // namespace vag
