// vag_math.cuh -- FP64 exp2 / log2 for the radiation and EATS kernels.
//
// The reference calls libm's std::exp2 / std::log2 (fast-math polynomials are compiled OFF,
// src/util/fast-math.h:40-77).  CUDA's libdevice equivalents materialise every polynomial
// coefficient with two UMOV/IMAD.MOV instructions per DFMA on sm_100a (ncu: 34 % of k_eats'
// issued instructions were such moves).  These versions keep the coefficients in __constant__
// memory (uniform LDCU.128 loads, two coefficients per instruction), use MUFU.RCP64H + Newton for
// the one division, and fall back to libdevice outside the fast domain.  Accuracy: <= 1 ulp-ish
// (exp2: |rel| <= 4.5e-16, log2: |abs| <= 2e-16 + 1.2e-16*|result|), far inside the 1e-6 flux bar.
// Host builds (oracle/hostemu) run the same polynomials.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "vag_common.cuh"

namespace vag {

// 2^f on [-0.5, 0.5], degree 10 (Chebyshev-node interpolation, scripts/gen_poly.py: 4.5e-16 with the
// FP64 Horner evaluation included)
#define VAG_EXP2_DEG 10
#define VAG_EXP2_COEF                                                                                          \
    0x1.0000000000000p+0, 0x1.62e42fefa3a19p-1, 0x1.ebfbdff82c598p-3, 0x1.c6b08d703ce4ap-5, 0x1.3b2ab6fba1e2ep-7, \
        0x1.5d87fe9d7a2a0p-10, 0x1.430913095b844p-13, 0x1.ffcb5406b78d0p-17, 0x1.62bfd4af386b2p-20,             \
        0x1.b675bc23e8f7dp-24, 0x1.e605db52afdbfp-28
// log2(m) = s * sum_n c_n s^(2n), s = (m-1)/(m+1), c_n = 2 / (ln2 (2n+1)), n = 0..10
#define VAG_LOG2_COEF                                                                                          \
    2.8853900817779268147, 0.96179669392597560491, 0.57707801635558536295, 0.41219858311113240211,            \
        0.32059889797532520164, 0.26230818925253880134, 0.22195308321368667806, 0.19235933878519512098,        \
        0.16972882833987804793, 0.15186263588304878025, 0.13739952770371080070

#if defined(__CUDACC__)
__constant__ double c_exp2[VAG_EXP2_DEG + 1] = {VAG_EXP2_COEF};
__constant__ double c_log2[11] = {VAG_LOG2_COEF};
#endif
static const double h_exp2[VAG_EXP2_DEG + 1] = {VAG_EXP2_COEF};
static const double h_log2[11] = {VAG_LOG2_COEF};

#if defined(__CUDA_ARCH__)
#define VAG_CEXP2 c_exp2
#define VAG_CLOG2 c_log2
#else
#define VAG_CEXP2 h_exp2
#define VAG_CLOG2 h_log2
#endif

VAG_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    std::memcpy(&d, &u, 8);
    return d;
#endif
}
VAG_HD uint64_t double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    std::memcpy(&u, &d, 8);
    return u;
#endif
}

VAG_HD double dexp2(double x) {
    if (!(x > -1021.0 && x < 1023.0)) return exp2(x);  // overflow / subnormal / NaN: libm path
    // round to nearest integer with the 1.5*2^52 trick
    const double magic = 6755399441055744.0;
    const double t = x + magic;
    const double kf = t - magic;
    const double f = x - kf;  // [-0.5, 0.5]
    const int64_t k = (int64_t)(double_to_bits(t) & 0xFFFFFFFFull) | ((double_to_bits(t) & 0x80000000ull) ? ~0xFFFFFFFFll : 0);
    double p = VAG_CEXP2[VAG_EXP2_DEG];
#pragma unroll
    for (int j = VAG_EXP2_DEG - 1; j >= 0; --j) p = fma(p, f, VAG_CEXP2[j]);
    return bits_to_double(double_to_bits(p) + ((uint64_t)k << 52));
}

// Guard-free variants for callers that guarantee a finite argument in (-1000, 1000) / a positive normal
// finite argument: no range check, no libm fallback, hence no branch in the caller's instruction stream.
VAG_HD double dexp2_nc(double x) {
    const double magic = 6755399441055744.0;
    const double t = x + magic;
    const double kf = t - magic;
    const double f = x - kf;
    const int64_t k = (int64_t)(double_to_bits(t) & 0xFFFFFFFFull) | ((double_to_bits(t) & 0x80000000ull) ? ~0xFFFFFFFFll : 0);
    double p = VAG_CEXP2[VAG_EXP2_DEG];
#pragma unroll
    for (int j = VAG_EXP2_DEG - 1; j >= 0; --j) p = fma(p, f, VAG_CEXP2[j]);
    return bits_to_double(double_to_bits(p) + ((uint64_t)k << 52));
}
VAG_HD double dlog2_nc(double x);
// NL independent exp2 evaluations with the Horner recurrences interleaved (one coefficient fetch serves all of
// them, and the NL dependent chains overlap); arguments as for dexp2_nc
template <int NL>
VAG_HD void dexp2_nc_vec(const double* x, double* out) {
    const double magic = 6755399441055744.0;
    double f[NL], p[NL];
    uint64_t sh[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const double t = x[l] + magic;
        f[l] = x[l] - (t - magic);
        const int64_t k = (int64_t)(double_to_bits(t) & 0xFFFFFFFFull) | ((double_to_bits(t) & 0x80000000ull) ? ~0xFFFFFFFFll : 0);
        sh[l] = (uint64_t)k << 52;
        p[l] = VAG_CEXP2[VAG_EXP2_DEG];
    }
#pragma unroll
    for (int j = VAG_EXP2_DEG - 1; j >= 0; --j) {
        const double c = VAG_CEXP2[j];
#pragma unroll
        for (int l = 0; l < NL; ++l) p[l] = fma(p[l], f[l], c);
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) out[l] = bits_to_double(double_to_bits(p[l]) + sh[l]);
}

VAG_HD double dlog2(double x) {
    if (!(x >= 2.2250738585072014e-308 && x < kInf)) return log2(x);  // 0, negative, subnormal, inf, NaN
    uint64_t u = double_to_bits(x);
    int e = (int)(u >> 52) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = bits_to_double(u);  // [1, 2)
    if (m > 1.4142135623730951) {
        m *= 0.5;
        e += 1;
    }
    const double num = m - 1.0, den = m + 1.0;
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    double s = num * r;
    s = fma(fma(-den, s, num), r, s);
#else
    const double s = num / den;
#endif
    const double z = s * s;
    double q = VAG_CLOG2[10];
#pragma unroll
    for (int j = 9; j >= 0; --j) q = fma(q, z, VAG_CLOG2[j]);
    return fma(s, q, (double)e);
}

VAG_HD double dlog2_nc(double x) {
    uint64_t u = double_to_bits(x);
    int e = (int)(u >> 52) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = bits_to_double(u);
    const bool hi = m > 1.4142135623730951;
    m = hi ? m * 0.5 : m;
    e = hi ? e + 1 : e;
    const double num = m - 1.0, den = m + 1.0;
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    double s = num * r;
    s = fma(fma(-den, s, num), r, s);
#else
    const double s = num / den;
#endif
    const double z = s * s;
    double q = VAG_CLOG2[10];
#pragma unroll
    for (int j = 9; j >= 0; --j) q = fma(q, z, VAG_CLOG2[j]);
    return fma(s, q, (double)e);
}

// ---------------------------------------------------------------------------------------------
// log2(1 + 2^x) on [-20, 20] as a table of local polynomials (the EATS hot loop evaluates this
// function four times per spectrum point; computing it as log2(1 + exp2(x)) costs two polynomial
// evaluations, a division and the exponent bookkeeping of both).
// 81 rows centred at x_i = -20 + i/2, degree 7 in u = 2 (x - x_i) in [-1/2, 1/2]; the function is
// analytic with its nearest singularities at x = +-i pi/ln 2 (|Im| = 4.53), so the Chebyshev
// interpolant converges like 36^-n: truncation < 5e-13 absolute (tests/test_host_logic.py checks it).
// Layout: four planes of 16-byte coefficient pairs, plane j = (c_2j, c_2j+1) of every row, row r at pair index
// r + (r >> 3).  A spectrum point's rows advance by a near-constant stride from one lattice node (thread) to the
// next, typically 1-4 rows; with the skew the eight threads of a quarter warp then read eight distinct 16-byte
// bank groups for every stride in {1, 2, 4, 8} (the row-major 64-byte rows this replaces offered two bank groups
// per load: 237 M of 449 M shared-load wavefronts of k_eats were bank-conflict replays, L1 data pipe 74 %).
// The table is generated on the host in long double (build_softplus_lut) and staged in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int SPL_ROWS = 81;
constexpr int SPL_DEG = 7;
constexpr int SPL_PLANE = 192;  // doubles per plane: 96 pair slots >= 81 + (80 >> 3) + 1, a multiple of 128 B
constexpr int SPL_DOUBLES = ((SPL_DEG + 1) / 2) * SPL_PLANE;
// index (in doubles) of monomial coefficient q of `row`
VAG_HD constexpr int spl_index(int row, int q) { return (q >> 1) * SPL_PLANE + ((row + (row >> 3)) << 1) + (q & 1); }

inline void build_softplus_lut(double* lut) {
    constexpr int n = SPL_DEG + 1;
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int row = 0; row < SPL_ROWS; ++row) {
        const long double xc = -20.0L + 0.5L * row;
        long double f[n], cheb[n];
        for (int k = 0; k < n; ++k) {
            const long double tk = cosl(pi * (k + 0.5L) / n);  // node on [-1, 1]; x = xc + tk / 4
            f[k] = log2l(1.0L + exp2l(xc + 0.25L * tk));
        }
        for (int j = 0; j < n; ++j) {
            long double acc = 0;
            for (int k = 0; k < n; ++k) acc += f[k] * cosl(pi * j * (k + 0.5L) / n);
            cheb[j] = acc * 2.0L / n;
        }
        cheb[0] *= 0.5L;
        // Chebyshev series -> monomials in t (T_0 = 1, T_1 = t, T_{j+1} = 2 t T_j - T_{j-1})
        long double mono[n] = {0}, Tprev[n] = {0}, Tcur[n] = {0}, Tnext[n];
        Tprev[0] = 1;
        Tcur[1] = 1;
        mono[0] += cheb[0];
        for (int q = 0; q < n; ++q) mono[q] += cheb[1] * Tcur[q];
        for (int j = 2; j < n; ++j) {
            for (int q = 0; q < n; ++q) Tnext[q] = (q > 0 ? 2 * Tcur[q - 1] : 0) - Tprev[q];
            for (int q = 0; q < n; ++q) {
                mono[q] += cheb[j] * Tnext[q];
                Tprev[q] = Tcur[q];
                Tcur[q] = Tnext[q];
            }
        }
        // t = 2 u  (u = 2 (x - xc) in [-1/2, 1/2])
        long double scale = 1;
        for (int q = 0; q < n; ++q) {
            lut[spl_index(row, q)] = (double)(mono[q] * scale);
            scale *= 2;
        }
    }
}

// src/util/fast-math.h:179-185 (log2_softplus) through the table
VAG_HD double log2_softplus_lut(const double* __restrict__ lut, double x) {
    if (!(x <= 20.0)) return x;  // also a NaN argument: it must not reach the table index
    if (x < -20.0) return 0.0;
    // row = round(2 x + 40) in [0, 80], read off the low word of 2 x + (40 + 1.5 * 2^52); u = 2 x + 40 - row
    const double magic40 = 6755399441055744.0 + 40.0;
    const double t = fma(x, 2.0, magic40);
    const int row = (int)(uint32_t)(double_to_bits(t) & 0xFFFFFFFFull);
    const double u = fma(x, 2.0, magic40 - t);  // magic40 - t = 40 - row exactly; u in [-1/2, 1/2]
#if defined(__CUDA_ARCH__)
    const double2* c2 = reinterpret_cast<const double2*>(lut) + (row + (row >> 3));
    const double2 c01 = c2[0], c23 = c2[SPL_PLANE / 2], c45 = c2[SPL_PLANE], c67 = c2[3 * SPL_PLANE / 2];
    double p = fma(c67.y, u, c67.x);
    p = fma(p, u, c45.y);
    p = fma(p, u, c45.x);
    p = fma(p, u, c23.y);
    p = fma(p, u, c23.x);
    p = fma(p, u, c01.y);
    p = fma(p, u, c01.x);
#else
    double p = lut[spl_index(row, SPL_DEG)];
    for (int j = SPL_DEG - 1; j >= 0; --j) p = fma(p, u, lut[spl_index(row, j)]);
#endif
    return p;
}

}  // namespace vag
