// vag_libm.cuh -- exp / exp2 / log / log2 / log10 / pow / sin / cos that reproduce the HOST libm's results
// bit for bit on the device (and, compiled by gcc, on the host).
//
// Why this exists.  The reference's (theta, phi) grid is the inverse CDF of an adaptive dopri5 quadrature
// (src/core/grid-refinement.h:137-189) whose integrand carries the Doppler cancellation (1-beta)/(1-beta cos):
// every pdf value has ~1e-11 of position-dependent rounding noise, the step-size controller turns that noise
// into ~1e-6 relative changes of the step sizes, and a ONE-ulp difference anywhere upstream of a step
// position (a libm result, an FMA contraction) therefore decorrelates the whole step sequence: the theta
// nodes move by ~1e-8 and the flux of a structured-jet reverse shock by up to 5e-3.
// CUDA's libdevice functions differ from glibc's in the last bit for ~10 % of the arguments.  "Identical to
// the reference" on the device therefore needs the reference's own arithmetic: the reference calls glibc
// (std::exp / std::pow / std::cos ..., fast-math polynomials compiled off, src/util/fast-math.h:40-77), and
// this header restates glibc 2.39's x86-64 algorithms -- the variants its IFUNC dispatch selects on an
// AVX2+FMA host, which is what every x86-64-v3 box runs:
//   exp, log, log2, pow   S. Nagy's table-driven routines (sysdeps/ieee754/dbl-64/e_exp.c, e_log.c, e_log2.c,
//                         e_pow.c; the __*_fma multiarch builds, i.e. with the FMA contractions gcc applies)
//   exp2                  e_exp2.c (no multiarch variant: baseline SSE2 build, no contraction)
//   log10                 e_log10.c (fdlibm form around log, baseline build)
//   sin, cos              IBM accurate mathematical library, s_sin.c (__sin_fma / __cos_fma)
// with the library's look-up tables read out of the installed libm.so.6 (scripts/gen_libm_tables.py ->
// vag_libm_tables.inc).  Every multiplication / addition is spelled out through gl::mul / add / fma so that
// neither nvcc (-fmad) nor gcc (-ffp-contract=fast) can re-associate or contract differently from the
// instruction sequence of the library build.  tests/test_libm_exact.py pins the host build against the live
// libm on millions of arguments (bit equality), the -m gpu tier pins the device build against the host build.
//
// Arguments outside the range the grid builder produces (subnormal / huge / non-finite) fall back to the
// platform function; those calls are not on a bit-critical path.
#pragma once

#include "vag_common.cuh"
#include "vag_math.cuh"

namespace vag {
namespace gl {

// ---- contraction-proof IEEE operations --------------------------------------------------------------
#if defined(__CUDA_ARCH__)
VAG_HD double mul(double a, double b) { return __dmul_rn(a, b); }
VAG_HD double add(double a, double b) { return __dadd_rn(a, b); }
VAG_HD double sub(double a, double b) { return __dsub_rn(a, b); }
VAG_HD double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
VAG_HD double div(double a, double b) { return __ddiv_rn(a, b); }
VAG_HD double sqrt(double a) { return __dsqrt_rn(a); }
#else
// gcc fuses a product into a following sum only when it sees the multiplication: the empty asm makes the
// rounded product opaque.  Sums of opaque values cannot be contracted.
VAG_HD double mul(double a, double b) {
    double r = a * b;
    __asm__("" : "+x"(r));
    return r;
}
VAG_HD double add(double a, double b) { return a + b; }
VAG_HD double sub(double a, double b) { return a - b; }
VAG_HD double fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
VAG_HD double div(double a, double b) {
    double r = a / b;  // opaque as well: -freciprocal-math style rewrites must not touch it
    __asm__("" : "+x"(r));
    return r;
}
VAG_HD double sqrt(double a) { return __builtin_sqrt(a); }
#endif
// -(a*b) + c and a*b - c, single rounding (vfnmadd / vfmsub)
VAG_HD double fnma(double a, double b, double c) { return fma(-a, b, c); }
VAG_HD double fms(double a, double b, double c) { return fma(a, b, -c); }

// ---- tables -------------------------------------------------------------------------------------------
#define VAG_LIBM_TABLE(name, n) static const unsigned long long name##_h[n]
#include "vag_libm_tables.inc"
#undef VAG_LIBM_TABLE
#if defined(__CUDACC__)
#define VAG_LIBM_TABLE(name, n) static __device__ const unsigned long long name##_d[n]
#include "vag_libm_tables.inc"
#undef VAG_LIBM_TABLE
#endif
#if defined(__CUDA_ARCH__)
#define VAG_GLT(name, i) bits_to_double(name##_d[i])
#define VAG_GLTU(name, i) ((uint64_t)name##_d[i])
#else
#define VAG_GLT(name, i) bits_to_double(name##_h[i])
#define VAG_GLTU(name, i) ((uint64_t)name##_h[i])
#endif

// ---- exp: e_exp.c (FMA build) ---------------------------------------------------------------------------
namespace k {
constexpr double InvLn2N = 0x1.71547652b82fep+7, Shift = 0x1.8p52;
constexpr double NegLn2hiN = -0x1.62e42fefa0000p-8, NegLn2loN = -0x1.cf79abc9e3b3ap-47;
constexpr double EC2 = 0x1.ffffffffffdbdp-2, EC3 = 0x1.555555555543cp-3, EC4 = 0x1.55555cf172b91p-5,
                 EC5 = 0x1.1111167a4d017p-7;
}  // namespace k

// exp(x + xtail) * (sign_bias ? -1 : 1) core shared by exp and pow; x in the main range (|x| in [2^-54, 512))
VAG_HD double exp_core(double x, double xtail, bool with_tail) {
    double kd = fma(x, k::InvLn2N, k::Shift);
    const uint64_t ki = double_to_bits(kd);
    kd = sub(kd, k::Shift);
    double r = fma(kd, k::NegLn2hiN, x);
    r = fma(kd, k::NegLn2loN, r);
    if (with_tail) r = add(xtail, r);
    const int idx = 2 * (int)(ki & 127);
    const uint64_t top = ki << 45;
    const double tail = VAG_GLT(GL_EXP_TAB, idx);
    const uint64_t sbits = VAG_GLTU(GL_EXP_TAB, idx + 1) + top;
    const double p23 = fma(k::EC3, r, k::EC2);
    const double tr = add(r, tail);
    const double r2 = mul(r, r);
    const double p45 = fma(r, k::EC5, k::EC4);
    const double t = fma(p23, r2, tr);
    const double r4 = mul(r2, r2);
    const double tmp = fma(r4, p45, t);
    const double scale = bits_to_double(sbits);
    return fma(scale, tmp, scale);
}

VAG_HD double exp(double x) {
    const uint32_t abstop = (uint32_t)(double_to_bits(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) return ::exp(x);  // |x| < 2^-54, |x| >= 512, inf, nan
    return exp_core(x, 0.0, false);
}

// ---- exp2: e_exp2.c (baseline build: separate multiplications and additions) ---------------------------
VAG_HD double exp2(double x) {
    constexpr double Shift2 = 0x1.8p45;
    constexpr double C1 = 0x1.62e42fefa39efp-1, C2 = 0x1.ebfbdff82c424p-3, C3 = 0x1.c6b08d70cf4b5p-5,
                     C4 = 0x1.3b2abd24650ccp-7, C5 = 0x1.5d7e09b4e3a84p-10;
    const uint32_t abstop = (uint32_t)(double_to_bits(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) return ::exp2(x);
    double kd = add(x, Shift2);
    const uint64_t ki = double_to_bits(kd);
    kd = sub(kd, Shift2);
    const double r = sub(x, kd);
    const int idx = 2 * (int)(ki & 127);
    const uint64_t top = ki << 45;
    const double tail = VAG_GLT(GL_EXP_TAB, idx);
    const uint64_t sbits = VAG_GLTU(GL_EXP_TAB, idx + 1) + top;
    const double r2 = mul(r, r);
    double a = add(mul(C3, r), C2);
    const double b = add(mul(C1, r), tail);
    double c = add(mul(r, C5), C4);
    a = mul(a, r2);
    const double r4 = mul(r2, r2);
    a = add(a, b);
    c = mul(c, r4);
    const double tmp = add(a, c);
    const double scale = bits_to_double(sbits);
    return add(scale, mul(tmp, scale));
}

// ---- log: e_log.c (FMA build) -----------------------------------------------------------------------------
VAG_HD double log(double x) {
    constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    constexpr double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3,
                     A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
    constexpr double B0 = -0x1p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3, B3 = 0x1.999999995dd0cp-3,
                     B4 = -0x1.55555556745a7p-3, B5 = 0x1.24924a344de30p-3, B6 = -0x1.fffffa4423d65p-4,
                     B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4, B9 = 0x1.78182f7afd085p-4,
                     B10 = -0x1.5521375d145cdp-4;
    const uint64_t ix = double_to_bits(x);
    if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {  // 1 - 2^-4 <= x < 1 + 0x1.09p-4
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = sub(x, 1.0);
        const double p12 = fma(B2, r, B1);
        const double p45 = fma(B5, r, B4);
        const double r2 = mul(r, r);
        const double p78 = fma(B8, r, B7);
        const double p123 = fma(r2, B3, p12);
        const double p456 = fma(r2, B6, p45);
        const double r3 = mul(r, r2);
        double q = fma(r2, B9, p78);
        q = fma(r3, B10, q);
        q = fma(q, r3, p456);
        q = fma(q, r3, p123);
        const double t = fma(r, 0x1p27, r);
        const double rhi = fnma(0x1p27, r, t);
        const double rhi2 = mul(rhi, rhi);
        const double rlo = sub(r, rhi);
        const double hi = fma(rhi2, B0, r);
        const double d = sub(r, hi);
        const double rsum = add(r, rhi);
        double lo = fma(rhi2, B0, d);
        const double h = mul(B0, rlo);
        lo = fma(h, rsum, lo);
        const double y = fma(q, r3, lo);
        return add(hi, y);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u > 0x7fdfu) return ::log(x);  // zero, negative, subnormal, inf, nan
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int kk = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = VAG_GLT(GL_LOG_TAB, 2 * i), logc = VAG_GLT(GL_LOG_TAB, 2 * i + 1);
    const double z = bits_to_double(iz);
    const double kd = (double)kk;
    const double r = fma(z, invc, -1.0);
    const double w = fma(Ln2hi, kd, logc);
    const double p12 = fma(A2, r, A1);
    const double hi = add(r, w);
    const double r2 = mul(r, r);
    double lo = sub(w, hi);
    lo = add(lo, r);
    lo = fma(kd, Ln2lo, lo);
    const double r3 = mul(r, r2);
    const double p34 = fma(r, A4, A3);
    lo = fma(r2, A0, lo);
    const double p = fma(p34, r2, p12);
    const double y = fma(r3, p, lo);
    return add(y, hi);
}

// ---- log2: e_log2.c (FMA build) ---------------------------------------------------------------------------
VAG_HD double log2(double x) {
    constexpr double InvLn2hi = 0x1.7154765200000p+0, InvLn2lo = 0x1.705fc2eefa200p-33;
    constexpr double A0 = -0x1.71547652b8339p-1, A1 = 0x1.ec709dc3a04bep-2, A2 = -0x1.7154764702ffbp-2,
                     A3 = 0x1.2776c50034c48p-2, A4 = -0x1.ec7b328ea92bcp-3, A5 = 0x1.a6225e117f92ep-3;
    constexpr double B0 = -0x1.71547652b82fep-1, B1 = 0x1.ec709dc3a03f7p-2, B2 = -0x1.71547652b7c3fp-2,
                     B3 = 0x1.2776c50f05be4p-2, B4 = -0x1.ec709dd768fe5p-3, B5 = 0x1.a61761ec4e736p-3,
                     B6 = -0x1.7153fbc64a79bp-3, B7 = 0x1.484d154f01b4ap-3, B8 = -0x1.289e4a72c383cp-3,
                     B9 = 0x1.0b32f285aee66p-3;
    const uint64_t ix = double_to_bits(x);
    if (ix - 0x3feea4af00000000ull <= 0x210a9ffffffffull) {  // 1 - 0x1.5b51p-5 <= x < 1 + 0x1.6ab2p-5
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = sub(x, 1.0);
        const double hi = mul(InvLn2hi, r);
        const double r2 = mul(r, r);
        double lo = fms(InvLn2hi, r, hi);
        const double r4 = mul(r2, r2);
        const double p01 = fma(B1, r, B0);
        lo = fma(r, InvLn2lo, lo);
        const double y = fma(p01, r2, hi);
        const double d = sub(hi, y);
        double l2 = fma(p01, r2, d);
        const double p23 = fma(B3, r, B2);
        l2 = add(l2, lo);
        double p45 = fma(B5, r, B4);
        p45 = fma(p45, r2, p23);
        const double p67 = fma(B7, r, B6);
        double q = fma(r, B9, B8);
        q = fma(q, r2, p67);
        q = fma(q, r4, p45);
        q = fma(q, r4, l2);
        return add(y, q);
    }
    const uint32_t top = (uint32_t)(ix >> 48);
    if (top - 0x0010u > 0x7fdfu) return ::log2(x);
    const uint64_t tmp = ix - 0x3fe6000000000000ull;
    const int i = (int)((tmp >> 46) & 63);
    const int kk = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = VAG_GLT(GL_LOG2_TAB, 2 * i), logc = VAG_GLT(GL_LOG2_TAB, 2 * i + 1);
    const double z = bits_to_double(iz);
    const double kd = (double)kk;
    const double t3 = add(kd, logc);
    const double r = fma(z, invc, -1.0);
    const double p01 = fma(A1, r, A0);
    const double t1 = mul(InvLn2hi, r);
    double t2 = fms(InvLn2hi, r, t1);
    const double hi = add(t1, t3);
    double lo = sub(t3, hi);
    t2 = fma(r, InvLn2lo, t2);
    const double r2 = mul(r, r);
    lo = add(lo, t1);
    lo = add(lo, t2);
    double p23 = fma(A3, r, A2);
    const double r4 = mul(r2, r2);
    const double p45 = fma(r, A5, A4);
    p23 = fma(p23, r2, p01);
    const double p = fma(p45, r4, p23);
    const double y = fma(r2, p, lo);
    return add(y, hi);
}

// ---- log10: e_log10.c (baseline build) around the FMA log ---------------------------------------------------
VAG_HD double log10(double x) {
    constexpr double ivln10 = 0x1.bcb7b1526e50ep-2, log10_2hi = 0x1.34413509f6000p-2, log10_2lo = 0x1.9fef311f12b36p-42;
    const int64_t hx = (int64_t)double_to_bits(x);
    if (hx < 0x0010000000000000ll || hx >= 0x7ff0000000000000ll) return ::log10(x);
    const int64_t kk = (hx >> 52) - 1023;
    const int64_t i = (int64_t)((uint64_t)kk >> 63);
    const uint64_t hx2 = ((uint64_t)hx & 0x000fffffffffffffull) | ((uint64_t)(0x3ff - i) << 52);
    const double y = (double)(kk + i);
    const double lx = gl::log(bits_to_double(hx2));
    const double z = add(mul(lx, ivln10), mul(y, log10_2lo));
    return add(z, mul(y, log10_2hi));
}

// ---- pow: e_pow.c (FMA build), positive finite normal x, moderate y ------------------------------------------
VAG_HD double pow(double x, double y) {
    constexpr double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    constexpr double A0 = -0x1p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1, A3 = 0x1.999999959554ep-1,
                     A4 = -0x1.555555529a47ap-1, A5 = -0x1.2495b9b4845e9p+0, A6 = 0x1.0002b8b263fc3p+0;
    const uint64_t ix = double_to_bits(x), iy = double_to_bits(y);
    const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
    // x zero / subnormal / negative / inf / nan, or |y| outside [2^-65, 2^63): not on the bit-exact path
    if (topx - 1u > 0x7fdu || (topy & 0x7ff) - 0x3beu > 0x7fu) return ::pow(x, y);
    // log_inline
    const uint64_t tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const int kk = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double z = bits_to_double(iz);
    const double kd = (double)kk;
    const double invc = VAG_GLT(GL_POWLOG_TAB, 3 * i), logc = VAG_GLT(GL_POWLOG_TAB, 3 * i + 1),
                 logctail = VAG_GLT(GL_POWLOG_TAB, 3 * i + 2);
    const double t1 = fma(Ln2hi, kd, logc);
    const double lo1 = fma(Ln2lo, kd, logctail);
    const double r = fma(z, invc, -1.0);
    const double ar = mul(r, A0);
    const double p12 = fma(A2, r, A1);
    const double p34 = fma(A4, r, A3);
    const double t2 = add(r, t1);
    const double lo2 = add(sub(t1, t2), r);
    const double ar2 = mul(r, ar);
    const double ar3 = mul(r, ar2);
    const double lo3 = fms(ar, r, ar2);
    const double hi = add(t2, ar2);
    double p56 = fma(r, A6, A5);
    const double lo4 = add(sub(t2, hi), ar2);
    p56 = fma(p56, ar2, p34);
    const double p = fma(ar2, p56, p12);
    double lo = add(lo1, lo2);
    lo = add(lo, lo3);
    lo = add(lo, lo4);
    lo = fma(ar3, p, lo);
    const double lhi = add(hi, lo);
    const double llo = add(sub(hi, lhi), lo);
    // exp_inline(y * log x)
    const double ehi = mul(y, lhi);
    double elo = fms(lhi, y, ehi);
    elo = fma(y, llo, elo);
    const uint32_t abstop = (uint32_t)(double_to_bits(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) return ::pow(x, y);  // result ~1, overflow or underflow
    return exp_core(ehi, elo, true);
}

// ---- sin / cos: s_sin.c (FMA build) ----------------------------------------------------------------------------
namespace k {
constexpr double big = 0x1.8p45, toint = 0x1.8p52;
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
constexpr double cs2 = 0x1p-1, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
                 s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
constexpr double hpinv = 0x1.45f306dc9c883p-1;
constexpr double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27;
constexpr double pp3 = -0x1.cb3b398000000p-55, pp4 = -0x1.d747f23e32ed7p-83;
}  // namespace k

VAG_HD double do_cos(double x, double dx) {
    if (x < 0) dx = -dx;
    const double ax = fabs(x);
    const double u = add(ax, k::big);
    const int kq = (int)(uint32_t)(double_to_bits(u) & 0xffffffffull) * 4;
    x = sub(ax, sub(u, k::big));
    x = add(x, dx);
    const double xx = mul(x, x);
    const double ps = fma(k::sn5, xx, k::sn3);
    const double x3 = mul(x, xx);
    const double s = fma(x3, ps, x);
    double pc = fma(k::cs6, xx, k::cs4);
    pc = fma(pc, xx, k::cs2);
    const double c = mul(xx, pc);
    const double sn = VAG_GLT(GL_SINCOS_TAB, kq), ssn = VAG_GLT(GL_SINCOS_TAB, kq + 1),
                 cs = VAG_GLT(GL_SINCOS_TAB, kq + 2), ccs = VAG_GLT(GL_SINCOS_TAB, kq + 3);
    double cor = fnma(ssn, s, ccs);
    cor = fnma(c, cs, cor);
    cor = fnma(s, sn, cor);
    return add(cs, cor);
}

VAG_HD double do_sin(double x, double dx) {
    const double xold = x;
    if (fabs(x) < 0.126) {  // TAYLOR_SIN
        const double xx = mul(x, x);
        double p = fma(k::s5, xx, k::s4);
        p = fma(p, xx, k::s3);
        p = fma(p, xx, k::s2);
        p = fma(p, xx, k::s1);
        const double hdx = mul(dx, 0.5);
        const double q = fms(p, x, hdx);
        const double t = fma(xx, q, dx);
        return add(x, t);
    }
    if (x <= 0) dx = -dx;
    const double ax = fabs(x);
    const double u = add(ax, k::big);
    const int kq = (int)(uint32_t)(double_to_bits(u) & 0xffffffffull) * 4;
    x = sub(ax, sub(u, k::big));
    const double xx = mul(x, x);
    const double ps = fma(k::sn5, xx, k::sn3);
    const double x3 = mul(x, xx);
    const double sd = fma(x3, ps, dx);
    double pc = fma(k::cs6, xx, k::cs4);
    pc = fma(pc, xx, k::cs2);
    const double s = add(x, sd);
    const double c0 = mul(xx, pc);
    const double c = fma(x, dx, c0);
    const double sn = VAG_GLT(GL_SINCOS_TAB, kq), ssn = VAG_GLT(GL_SINCOS_TAB, kq + 1),
                 cs = VAG_GLT(GL_SINCOS_TAB, kq + 2), ccs = VAG_GLT(GL_SINCOS_TAB, kq + 3);
    double cor = fma(ccs, s, ssn);
    cor = fnma(c, sn, cor);
    cor = fma(s, cs, cor);
    const double res = add(sn, cor);
    return copysign(res, xold);
}

// x = n pi/2 + a + da, |x| < 105414350
VAG_HD int reduce_sincos(double x, double& a, double& da) {
    const double t = fma(x, k::hpinv, k::toint);
    const double xn = sub(t, k::toint);
    const int n = (int)(double_to_bits(t) & 3);
    double y = fnma(xn, k::mp1, x);
    y = fnma(xn, k::mp2, y);
    const double t2 = fnma(xn, k::pp3, y);
    double db = sub(y, t2);
    db = fnma(xn, k::pp3, db);
    const double b = fnma(xn, k::pp4, t2);
    double d2 = sub(t2, b);
    d2 = fnma(xn, k::pp4, d2);
    a = b;
    da = add(db, d2);
    return n;
}

VAG_HD double do_sincos(double a, double da, int n) {
    const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
    return (n & 2) ? -r : r;
}

VAG_HD double cos(double x) {
    const uint32_t kx = (uint32_t)(double_to_bits(x) >> 32) & 0x7fffffffu;
    if (kx < 0x3e400000u) return 1.0;
    if (kx < 0x3feb6000u) return do_cos(x, 0.0);
    if (kx < 0x400368fdu) {
        const double y = sub(k::hp0, fabs(x));
        const double a = add(y, k::hp1);
        const double da = add(sub(y, a), k::hp1);
        return do_sin(a, da);
    }
    if (kx < 0x419921fbu) {
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n + 1);
    }
    return ::cos(x);
}

VAG_HD double sin(double x) {
    const uint32_t kx = (uint32_t)(double_to_bits(x) >> 32) & 0x7fffffffu;
    if (kx < 0x3e500000u) return x;
    if (kx < 0x3feb6000u) return do_sin(x, 0.0);
    if (kx < 0x400368fdu) {
        const double t = sub(k::hp0, fabs(x));
        return copysign(do_cos(t, k::hp1), x);
    }
    if (kx < 0x419921fbu) {
        double a, da;
        const int n = reduce_sincos(x, a, da);
        return do_sincos(a, da, n);
    }
    return ::sin(x);
}

}  // namespace gl
}  // namespace vag
