// vag_dopri5.cuh -- register-resident Dormand-Prince 5(4) with FSAL, dense output and the step
// controller of Boost.odeint 1.82 (the integrator the reference instantiates as
// make_dense_output(rtol, rtol, runge_kutta_dopri5<State>()), src/dynamics/forward-shock.tpp:191,
// src/dynamics/reverse-shock.tpp:539, src/core/grid-refinement.h:147).
//
// Semantics restated from the published algorithm as vendored by the reference:
//   tableau / stages        external/boost/numeric/odeint/stepper/runge_kutta_dopri5.hpp:92-158
//   error coefficients      :163-198
//   dense output            :229-275 (Hairer-Norsett-Wanner I, p.191)
//   error norm              external/boost/numeric/odeint/stepper/controlled_runge_kutta.hpp:64-90,
//                           algebra/default_operations.hpp:431-444 (inf-norm, a_x = a_dxdt = 1)
//   step decrease/increase  controlled_runge_kutta.hpp:114-153 (order 5, error order 4)
//   do_step retry loop      stepper/dense_output_runge_kutta.hpp:324-345, 500-failure cap from
//                           integrate/max_step_checker.hpp:84-107
//
// Layout: all N-vectors are fixed-size arrays indexed with compile-time-unrolled loops so they
// live in registers (N = 1 grid CDF, 5 forward shock, 11 forward+reverse shock pair).
#pragma once

#include "vag_common.cuh"
#include "vag_math.cuh"

namespace vag {

#if defined(VAG_INSTRUMENT) && !defined(__CUDA_ARCH__)
// host-only analysis hook (scripts/step_stats.cpp): attempts / rejections of the current row
struct StepStats { long attempts = 0, rejects = 0, attempts_s = 0, rejects_s = 0; };
inline StepStats g_step_stats;
#define VAG_COUNT_ATTEMPT() (++g_step_stats.attempts)
#define VAG_COUNT_REJECT() (++g_step_stats.rejects)
#define VAG_COUNT_ATTEMPT_S() (++g_step_stats.attempts_s)
#define VAG_COUNT_REJECT_S() (++g_step_stats.rejects_s)
#else
#define VAG_COUNT_ATTEMPT() ((void)0)
#define VAG_COUNT_REJECT() ((void)0)
#define VAG_COUNT_ATTEMPT_S() ((void)0)
#define VAG_COUNT_REJECT_S() ((void)0)
#endif

// err^p of the step-size controller (err finite and > 0 at both call sites): exp2(p log2 err) with the
// constant-memory polynomials of vag_math.cuh instead of libdevice's pow (~4x fewer instructions)
// EXACT selects libm / IEEE arithmetic: the grid builder's CDF quadrature (N = 1) amplifies last-bit
// differences of its step sizes into the theta grid, so it stays on the operators the reference uses.
template <bool EXACT>
VAG_HD double ctrl_pow(double err, double p) {
    if (EXACT) return pow(err, p);
    // both call sites pass a positive err (> 1 when a step is rejected, in [5^-5, 1/2) when the next step grows) and a
    // negative exponent: clamped to the range of the guard-free polynomials, an infinite error gives the same 0 the
    // library call would (the caller's max(., 0.2) then applies) -- no range checks, no fallback call in the step loop
    return dexp2_nc(vmax(p * dlog2_nc(vmin(err, 1e300)), -1000.0));
}
template <bool EXACT>
VAG_HD double ctrl_div(double a, double b) {
    if (EXACT) return a / b;
    return vdiv(a, b);
}

template <int N>
struct Dopri5 {
    double x[N];     // current state
    double k1[N];    // derivative at current state (FSAL)
    double xo[N];    // state at the start of the last accepted step
    double ko[N];    // derivative at the start of the last accepted step
    double k3[N], k4[N], k5[N], k6[N];  // stages of the last attempted step
    double t, t_old, dt;
    double eps;      // eps_abs = eps_rel
    bool deriv_ready;

    VAG_HD void initialize(const double* x0, double t0, double dt0, double tol) {
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = x0[i];
        t = t0;
        t_old = t0;
        dt = dt0;
        eps = tol;
        deriv_ready = false;
    }

    // One accepted step (retrying with smaller dt on rejection).  Returns false when 500
    // consecutive attempts were rejected (Boost throws std::runtime_error there).
    template <class Sys>
    VAG_HD bool do_step(Sys& sys) {
        begin_step(sys);
        for (int fails = 0;;) {
            if (try_step(sys)) return true;
            if (++fails >= 500) return false;
        }
    }

    // First half of dense_output_runge_kutta::do_step: lazily evaluate the FSAL derivative.
    template <class Sys>
    VAG_HD void begin_step(Sys& sys) {
        if (!deriv_ready) {
            sys(x, k1, t);
            deriv_ready = true;
        }
        t_old = t;
    }

    // One attempt with the current dt (controlled_runge_kutta::try_step): true = accepted.
    // The six stage evaluations are sys(.., t + a_i dt) for a_i = 1/5, 3/10, 4/5, 8/9, 1, 1.
    template <class Sys>
    VAG_HD bool try_step(Sys& sys) {
        constexpr double a2 = 1.0 / 5, a3 = 3.0 / 10, a4 = 4.0 / 5, a5 = 8.0 / 9;
        constexpr double b21 = 1.0 / 5;
        constexpr double b31 = 3.0 / 40, b32 = 9.0 / 40;
        constexpr double b41 = 44.0 / 45, b42 = -56.0 / 15, b43 = 32.0 / 9;
        constexpr double b51 = 19372.0 / 6561, b52 = -25360.0 / 2187, b53 = 64448.0 / 6561, b54 = -212.0 / 729;
        constexpr double b61 = 9017.0 / 3168, b62 = -355.0 / 33, b63 = 46732.0 / 5247, b64 = 49.0 / 176,
                         b65 = -5103.0 / 18656;
        constexpr double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
        constexpr double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                         dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;

        {
            VAG_COUNT_ATTEMPT();
            double xt[N], k2[N], k7[N];
#pragma unroll
            for (int i = 0; i < N; ++i) xt[i] = 1.0 * x[i] + (dt * b21) * k1[i];
            sys(xt, k2, t + dt * a2);
#pragma unroll
            for (int i = 0; i < N; ++i) xt[i] = 1.0 * x[i] + (dt * b31) * k1[i] + (dt * b32) * k2[i];
            sys(xt, k3, t + dt * a3);
#pragma unroll
            for (int i = 0; i < N; ++i)
                xt[i] = 1.0 * x[i] + (dt * b41) * k1[i] + (dt * b42) * k2[i] + (dt * b43) * k3[i];
            sys(xt, k4, t + dt * a4);
#pragma unroll
            for (int i = 0; i < N; ++i)
                xt[i] = 1.0 * x[i] + (dt * b51) * k1[i] + (dt * b52) * k2[i] + (dt * b53) * k3[i] + (dt * b54) * k4[i];
            sys(xt, k5, t + dt * a5);
#pragma unroll
            for (int i = 0; i < N; ++i)
                xt[i] = 1.0 * x[i] + (dt * b61) * k1[i] + (dt * b62) * k2[i] + (dt * b63) * k3[i] +
                        (dt * b64) * k4[i] + (dt * b65) * k5[i];
            sys(xt, k6, t + dt);
#pragma unroll
            for (int i = 0; i < N; ++i)
                xt[i] = 1.0 * x[i] + (dt * c1) * k1[i] + (dt * c3) * k3[i] + (dt * c4) * k4[i] + (dt * c5) * k5[i] +
                        (dt * c6) * k6[i];
            sys(xt, k7, t + dt);

            // error estimate and inf-norm of err_i / (eps + eps * (|x_i| + |dt| |k1_i|))
            double err = 0;
            const double adt = fabs(dt);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double e = (dt * dc1) * k1[i] + (dt * dc3) * k3[i] + (dt * dc4) * k4[i] + (dt * dc5) * k5[i] +
                                 (dt * dc6) * k6[i] + (dt * dc7) * k7[i];
                const double r = ctrl_div<N == 1>(fabs(e), eps + eps * (1.0 * fabs(x[i]) + adt * fabs(k1[i])));
                // boost norm_inf: max over |r| starting from 0 with std::max(init, |r|)
                err = vmax(err, fabs(r));
            }

            if (err > 1.0) {
                VAG_COUNT_REJECT();
                dt *= vmax(0.9 * ctrl_pow<N == 1>(err, -1.0 / 3.0), 0.2);
                return false;
            }
            // accept
#pragma unroll
            for (int i = 0; i < N; ++i) {
                xo[i] = x[i];
                ko[i] = k1[i];
                x[i] = xt[i];
                k1[i] = k7[i];
            }
            t += dt;
            if (err < 0.5) {
                err = vmax(pow(5.0, -5.0), err);
                dt *= 0.9 * ctrl_pow<N == 1>(err, -1.0 / 5.0);
            }
            return true;
        }
    }

    // Dense-output weights of time tq inside the last accepted step [t_old, t]:
    // out_i = xo_i + w[0] ko_i + w[1] k3_i + w[2] k4_i + w[3] k5_i + w[4] k6_i + w[5] k1_i
    VAG_HD void dense_weights(double tq, double* w) const {
        constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
        const double h = t - t_old;
        const double th = (tq - t_old) / h;
        const double X1 = 5.0 * (2558722523.0 - 31403016.0 * th) / 11282082432.0;
        const double X3 = 100.0 * (882725551.0 - 15701508.0 * th) / 32700410799.0;
        const double X4 = 25.0 * (443332067.0 - 31403016.0 * th) / 1880347072.0;
        const double X5 = 32805.0 * (23143187.0 - 3489224.0 * th) / 199316789632.0;
        const double X6 = 55.0 * (29972135.0 - 7076736.0 * th) / 822651844.0;
        const double X7 = 10.0 * (7414447.0 - 829305.0 * th) / 29380423.0;
        const double thm1 = th - 1.0;
        const double thsq = th * th;
        const double A = thsq * (3.0 - 2.0 * th);
        const double B = thsq * thm1;
        const double C = thsq * thm1 * thm1;
        const double D = th * thm1 * thm1;
        w[0] = h * (A * b1 - C * X1 + D);
        w[1] = h * (A * b3 + C * X3);
        w[2] = h * (A * b4 - C * X4);
        w[3] = h * (A * b5 + C * X5);
        w[4] = h * (A * b6 - C * X6);
        w[5] = h * (B + C * X7);
    }
    // one component of the dense output (same expression as calc_state)
    template <int I>
    VAG_HD double dense_component(const double* w) const {
        return 1.0 * xo[I] + w[0] * ko[I] + w[1] * k3[I] + w[2] * k4[I] + w[3] * k5[I] + w[4] * k6[I] + w[5] * k1[I];
    }

    // Dense output at time tq inside the last accepted step [t_old, t].
    VAG_HD void calc_state(double tq, double* out) const {
        constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
        // the rational constants of the published formula as reciprocal multipliers (<= 1 ulp per weight, as in
        // Dopri5S::dense_weights): six IEEE divisions by constants per lattice node were ~9 % of the forward-shock kernel
        constexpr double r1 = 5.0 / 11282082432.0, r3 = 100.0 / 32700410799.0, r4 = 25.0 / 1880347072.0,
                         r5 = 32805.0 / 199316789632.0, r6 = 55.0 / 822651844.0, r7 = 10.0 / 29380423.0;
        const double h = t - t_old;
        const double th = vdiv(tq - t_old, h);  // h > 0: an accepted step
        const double X1 = r1 * (2558722523.0 - 31403016.0 * th);
        const double X3 = r3 * (882725551.0 - 15701508.0 * th);
        const double X4 = r4 * (443332067.0 - 31403016.0 * th);
        const double X5 = r5 * (23143187.0 - 3489224.0 * th);
        const double X6 = r6 * (29972135.0 - 7076736.0 * th);
        const double X7 = r7 * (7414447.0 - 829305.0 * th);
        const double thm1 = th - 1.0;
        const double thsq = th * th;
        const double A = thsq * (3.0 - 2.0 * th);
        const double B = thsq * thm1;
        const double C = thsq * thm1 * thm1;
        const double D = th * thm1 * thm1;
        const double w1 = h * (A * b1 - C * X1 + D);
        const double w3 = h * (A * b3 + C * X3);
        const double w4 = h * (A * b4 - C * X4);
        const double w5 = h * (A * b5 + C * X5);
        const double w6 = h * (A * b6 - C * X6);
        const double w7 = h * (B + C * X7);
#pragma unroll
        for (int i = 0; i < N; ++i)
            out[i] = 1.0 * xo[i] + w1 * ko[i] + w3 * k3[i] + w4 * k4[i] + w5 * k5[i] + w6 * k6[i] + w7 * k1[i];
    }
};

// ---------------------------------------------------------------------------------------------
// Dopri5S: the same integrator with the stage vectors in a caller-provided column of (shared)
// memory instead of registers.  Used by the shock ODE kernel (N = 5 / 11): only the state x lives in
// registers, the seven stage derivatives and the step-start state sit in the thread's private column
// `col[(slot * N + i) * stride]`, and the six stage evaluations run as a rolled loop, so the RHS exists
// once in the instruction stream (the fully unrolled register version was 40 k SASS instructions and
// ran out of the instruction cache with one warp per SM).  Same tableau, same operation order.
// ---------------------------------------------------------------------------------------------
#define VAG_DP_A {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0}
#define VAG_DP_B                                                                                             \
    {                                                                                                        \
        {1.0 / 5, 0, 0, 0, 0, 0}, {3.0 / 40, 9.0 / 40, 0, 0, 0, 0}, {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0}, \
            {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},                           \
            {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},                    \
            {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}                            \
    }
#if defined(__CUDACC__)
__constant__ double c_dp_a[6] = VAG_DP_A;
__constant__ double c_dp_b[6][6] = VAG_DP_B;
#endif
static const double h_dp_a[6] = VAG_DP_A;
static const double h_dp_b[6][6] = VAG_DP_B;
#if defined(__CUDA_ARCH__)
#define VAG_DPA c_dp_a
#define VAG_DPB c_dp_b
#else
#define VAG_DPA h_dp_a
#define VAG_DPB h_dp_b
#endif

template <int N>
struct Dopri5S {
    enum { S_K1 = 0, S_K2, S_K3, S_K4, S_K5, S_K6, S_K7, S_XO, NSLOT };
    static constexpr int kDoublesPerThread = NSLOT * N;
    double x[N];  // current state
    double* col;  // this thread's column
    int stride;
    double t, t_old, dt, eps;

    VAG_HD double& K(int slot, int i) const { return col[(size_t)(slot * N + i) * stride]; }

    VAG_HD void initialize(double* column, int column_stride, const double* x0, double t0, double dt0, double tol) {
        col = column;
        stride = column_stride;
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = x0[i];
        t = t0;
        t_old = t0;
        dt = dt0;
        eps = tol;
    }

    // derivative at the initial state (the FSAL slot of the first step)
    template <class Sys>
    VAG_HD void begin(Sys& sys) {
        double kk[N];
        sys(x, kk, t);
#pragma unroll
        for (int i = 0; i < N; ++i) K(S_K1, i) = kk[i];
        t_old = t;
    }

    // One attempt with the current dt (controlled_runge_kutta::try_step): true = accepted.  After an
    // accepted attempt the dense output of [t_old, t] is available until advance() is called.
    template <class Sys>
    VAG_HD bool try_step(Sys& sys) {
        constexpr double c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
        constexpr double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
                         dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
        VAG_COUNT_ATTEMPT_S();
        double xt[N];
#pragma unroll 1
        for (int s = 0; s < 6; ++s) {
#pragma unroll
            for (int i = 0; i < N; ++i) xt[i] = 1.0 * x[i];
            // All six tableau columns as straight-line code, the absent ones (zero above the diagonal, and k2 in the
            // 5th-order row) deselected: a rolled j loop cuts the stage into basic blocks of one column each, and
            // with one warp per scheduler the N independent accumulations only overlap inside a block.  The
            // deselected term is never added (a stale slot may hold anything), so the sum is the reference's.
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double b = VAG_DPB[s][j];
                const bool use = b != 0;
                const double db = dt * b;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double v = xt[i] + db * K(j, i);
                    xt[i] = use ? v : xt[i];
                }
            }
            double kk[N];
            sys(xt, kk, t + dt * VAG_DPA[s]);
#pragma unroll
            for (int i = 0; i < N; ++i) K(s + 1, i) = kk[i];
        }
        // error estimate and inf-norm of err_i / (eps + eps * (|x_i| + |dt| |k1_i|))
        double err = 0;
        const double adt = fabs(dt);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double k1i = K(S_K1, i);
            const double e = (dt * dc1) * k1i + (dt * dc3) * K(S_K3, i) + (dt * dc4) * K(S_K4, i) + (dt * dc5) * K(S_K5, i) +
                             (dt * dc6) * K(S_K6, i) + (dt * dc7) * K(S_K7, i);
            // the scale eps + eps (|x| + |dt| |k1|) is positive and finite for a finite state: branch-free
            // division (a non-finite state gives a non-finite ratio either way)
            const double r = vdiv(fabs(e), eps + eps * (1.0 * fabs(x[i]) + adt * fabs(k1i)));
            err = vmax(err, fabs(r));
        }
        if (err > 1.0) {
            VAG_COUNT_REJECT_S();
            dt *= vmax(0.9 * ctrl_pow<false>(err, -1.0 / 3.0), 0.2);
            return false;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            K(S_XO, i) = x[i];
            x[i] = xt[i];
        }
        t += dt;
        if (err < 0.5) {
            err = vmax(pow(5.0, -5.0), err);
            dt *= 0.9 * ctrl_pow<false>(err, -1.0 / 5.0);
        }
        return true;
    }

    // start of the next step: the FSAL derivative k7 becomes k1
    VAG_HD void advance() {
#pragma unroll
        for (int i = 0; i < N; ++i) K(S_K1, i) = K(S_K7, i);
        t_old = t;
    }

    // Dense-output weights (runge_kutta_dopri5.hpp:229-275).  The rational constants of the published
    // formula are folded into reciprocal multipliers (<= 1 ulp per weight against the divisions).
    VAG_HD double dense_theta(double tq) const { return vdiv(tq - t_old, t - t_old); }  // t > t_old: an accepted step
    VAG_HD void dense_weights(double tq, double* w) const {
        constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
        constexpr double r1 = 5.0 / 11282082432.0, r3 = 100.0 / 32700410799.0, r4 = 25.0 / 1880347072.0,
                         r5 = 32805.0 / 199316789632.0, r6 = 55.0 / 822651844.0, r7 = 10.0 / 29380423.0;
        const double h = t - t_old;
        const double th = vdiv(tq - t_old, h);
        const double X1 = r1 * (2558722523.0 - 31403016.0 * th);
        const double X3 = r3 * (882725551.0 - 15701508.0 * th);
        const double X4 = r4 * (443332067.0 - 31403016.0 * th);
        const double X5 = r5 * (23143187.0 - 3489224.0 * th);
        const double X6 = r6 * (29972135.0 - 7076736.0 * th);
        const double X7 = r7 * (7414447.0 - 829305.0 * th);
        const double thm1 = th - 1.0;
        const double thsq = th * th;
        const double A = thsq * (3.0 - 2.0 * th);
        const double B = thsq * thm1;
        const double C = thsq * thm1 * thm1;
        const double D = th * thm1 * thm1;
        w[0] = h * (A * b1 - C * X1 + D);
        w[1] = h * (A * b3 + C * X3);
        w[2] = h * (A * b4 - C * X4);
        w[3] = h * (A * b5 + C * X5);
        w[4] = h * (A * b6 - C * X6);
        w[5] = h * (B + C * X7);
    }
    // The same interpolant of ONE component as a polynomial in theta = (tq - t_old) / h, for callers
    // that evaluate it many times inside one step (the crossing-time bisection):
    //   x(theta) = p0 + A p1 + C (p2 + p3 theta) + D p4 + B p5,  A..D as in dense_weights
    VAG_HD void dense_poly(int i, double* p) const {
        constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
        constexpr double r1 = 5.0 / 11282082432.0, r3 = 100.0 / 32700410799.0, r4 = 25.0 / 1880347072.0,
                         r5 = 32805.0 / 199316789632.0, r6 = 55.0 / 822651844.0, r7 = 10.0 / 29380423.0;
        const double h = t - t_old;
        const double ko = K(S_K1, i), k3 = K(S_K3, i), k4 = K(S_K4, i), k5 = K(S_K5, i), k6 = K(S_K6, i), k7 = K(S_K7, i);
        p[0] = K(S_XO, i);
        p[1] = h * (b1 * ko + b3 * k3 + b4 * k4 + b5 * k5 + b6 * k6);
        p[2] = h * (-(r1 * 2558722523.0) * ko + (r3 * 882725551.0) * k3 - (r4 * 443332067.0) * k4 + (r5 * 23143187.0) * k5 -
                    (r6 * 29972135.0) * k6 + (r7 * 7414447.0) * k7);
        p[3] = h * ((r1 * 31403016.0) * ko - (r3 * 15701508.0) * k3 + (r4 * 31403016.0) * k4 - (r5 * 3489224.0) * k5 +
                    (r6 * 7076736.0) * k6 - (r7 * 829305.0) * k7);
        p[4] = h * ko;
        p[5] = h * k7;
    }
    VAG_HD double dense_poly_eval(const double* p, double th) const {
        const double thm1 = th - 1.0;
        const double thsq = th * th;
        const double A = thsq * (3.0 - 2.0 * th);
        const double B = thsq * thm1;
        const double C = thsq * thm1 * thm1;
        const double D = th * thm1 * thm1;
        return p[0] + A * p[1] + C * (p[2] + p[3] * th) + D * p[4] + B * p[5];
    }
    VAG_HD double dense_component(const double* w, int i) const {
        return 1.0 * K(S_XO, i) + w[0] * K(S_K1, i) + w[1] * K(S_K3, i) + w[2] * K(S_K4, i) + w[3] * K(S_K5, i) +
               w[4] * K(S_K6, i) + w[5] * K(S_K7, i);
    }
    // Dense output at time tq inside the last accepted step [t_old, t] (before advance()).
    VAG_HD void calc_state(double tq, double* out) const {
        double w[6];
        dense_weights(tq, w);
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = dense_component(w, i);
    }
};

}  // namespace vag
