// lfb_math.cuh -- device arithmetic of the filter-PDE right-hand side.
//
// Two arithmetic flavours, selected by the template parameter T:
//   * T = sd  ("strict double"): every operation is an individually rounded IEEE-754 operation
//     (__dadd_rn/__dmul_rn/__ddiv_rn never contract into FMAs) written in the reference's order of
//     operations, so the result is bit-comparable with the CPU oracle (oracle/lf_oracle.c, built
//     with -ffp-contract=off).  Used by parity tests and by handles created with desc.strict = 1.
//   * T = double: the production flavour.  WENO5-Z is evaluated in difference form with a common
//     denominator (one division per face instead of four); algebraically identical, differs from
//     the strict flavour by O(1e-16) relative rounding.
//
// Reference numerics (third-party Oceananigans WENO(order=5) reached through div_Uc in
// lagrangian_filter_tendency_kernel_functions.jl:56): SURVEY.md Appendix A.4-A.5.
#pragma once

#include <cuda_runtime.h>

struct sd {
    double v;
    __device__ __forceinline__ sd() {}
    __device__ __forceinline__ sd(double x) : v(x) {}
};
__device__ __forceinline__ sd operator+(sd a, sd b) { return sd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a, sd b) { return sd(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator*(sd a, sd b) { return sd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator/(sd a, sd b) { return sd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a) { return sd(-a.v); }

__device__ __forceinline__ double val(double x) { return x; }
__device__ __forceinline__ double val(sd x) { return x.v; }
__device__ __forceinline__ double absT(double x) { return fabs(x); }
__device__ __forceinline__ sd absT(sd x) { return sd(fabs(x.v)); }

#define LFB_WENO_EPS 1e-8

// ---- literal formulas (reference order of operations) -----------------------------------------
// WENO5-Z, left-biased argument order (p, o, n | e, f): upwind cell is n, face lies between n and e.
template <class T>
__device__ __forceinline__ T weno5z_ref(T p, T o, T n, T e, T f)
{
    const T eps(LFB_WENO_EPS);
    T b0 = n * (T(10.0) * n - T(31.0) * e + T(11.0) * f) + e * (T(25.0) * e - T(19.0) * f) + f * (T(4.0) * f);
    T b1 = o * (T(4.0) * o - T(13.0) * n + T(5.0) * e) + n * (T(13.0) * n - T(13.0) * e) + e * (T(4.0) * e);
    T b2 = p * (T(4.0) * p - T(19.0) * o + T(11.0) * n) + o * (T(25.0) * o - T(31.0) * n) + n * (T(10.0) * n);
    T q0 = n / T(3.0) + T(5.0) * e / T(6.0) - f / T(6.0);
    T q1 = -o / T(6.0) + T(5.0) * n / T(6.0) + e / T(3.0);
    T q2 = p / T(3.0) - T(7.0) * o / T(6.0) + T(11.0) * n / T(6.0);
    T tau = absT(b0 - b2);
    T r0 = tau / (b0 + eps), r1 = tau / (b1 + eps), r2 = tau / (b2 + eps);
    T a0 = T(0.3) * (T(1.0) + r0 * r0), a1 = T(0.6) * (T(1.0) + r1 * r1), a2 = T(0.1) * (T(1.0) + r2 * r2);
    return (a0 * q0 + a1 * q1 + a2 * q2) / (a0 + a1 + a2);
}

// WENO3-Z buffer scheme, left-biased order (o, n | e).
template <class T>
__device__ __forceinline__ T weno3z_ref(T o, T n, T e)
{
    const T eps(LFB_WENO_EPS);
    T b0 = (e - n) * (e - n), b1 = (n - o) * (n - o);
    T q0 = (n + e) / T(2.0), q1 = -o / T(2.0) + T(3.0) * n / T(2.0);
    T tau = absT(b0 - b1);
    T r0 = tau / (b0 + eps), r1 = tau / (b1 + eps);
    T a0 = T(2.0 / 3.0) * (T(1.0) + r0 * r0), a1 = T(1.0 / 3.0) * (T(1.0) + r1 * r1);
    return (a0 * q0 + a1 * q1) / (a0 + a1);
}

// ---- production formulas ------------------------------------------------------------------------
// WENO5-Z from the four first differences of the upwind-ordered stencil
//   dp = o - p, do_ = n - o, dn = e - n, de = f - e          (face between n and e, upwind cell n)
// Smoothness indicators in sum-of-squares form (3x Jiang-Shu, as Oceananigans):
//   b0 = 13/4 (de - dn)^2 + 3/4 (de - 3 dn)^2
//   b1 = 13/4 (dn - do)^2 + 3/4 (dn +  do)^2
//   b2 = 13/4 (do - dp)^2 + 3/4 (3 do - dp)^2
// Z weights a_s = C_s (1 + tau^2 / (b_s + eps)^2) brought to the common denominator
// prod_s (b_s + eps)^2, so that face value = n + sum_s A_s X_s / sum_s A_s needs one division,
// with X_s = q_s - n the candidate increments; written as X_1 - (A0 y0 / 2 + A2 y2 / 3) / sum C_s A_s.
__device__ __forceinline__ double weno5z_increment(double dp, double do_, double dn, double de)
{
    // the short form of the TMA z-march kernel (lfb_ztma.cu, zt_inc / zt_blend): indicators scaled by 4/3 with eps
    // folded into the curvature term, tau^2 folded into the products, one reciprocal (hardware seed + one cubic step)
    const double k133 = 13.0 / 3.0, eps43 = LFB_WENO_EPS * (4.0 / 3.0);
    const double s0 = de - dn, s1 = dn - do_, s2 = do_ - dp;
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double g0 = fma(t0, t0, fma(k133 * s0, s0, eps43));
    const double g1 = fma(t1, t1, fma(k133 * s1, s1, eps43));
    const double g2 = fma(t2, t2, fma(k133 * s2, s2, eps43));
    const double tau = g0 - g2;
    const double G0 = g0 * g0, G1 = g1 * g1, G2 = g2 * g2;
    const double A0 = fma(tau, tau, G0) * (G1 * G2);
    const double A1 = fma(tau, tau, G1) * (G0 * G2);
    const double A2 = fma(tau, tau, G2) * (G0 * G1);
    const double den = fma(3.0, A0, fma(6.0, A1, A2));
    const double num = fma(0.5, A0 * (s0 - s1), (1.0 / 3.0) * (A2 * (s1 - s2)));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    const double e = fma(-den, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    return fma(-num, r, fma(1.0 / 3.0, dn, (1.0 / 6.0) * do_));
}

// Upwind face value at face m (1-based; between cells m-1 and m) along one axis.
//   cm: pointer to cell m; s: stride along the axis; N: global cell count; topo: axis topology;
//   q: advecting face velocity.
// Bounded axes degrade WENO5 -> WENO3 -> Centered2 by face index; on the wall
// faces m = 1, N+1 the zero-flux halo cell mirrors the interior, so Centered2 returns the
// interior value.
template <class T, bool STRICT>
__device__ __forceinline__ T face_value(const double *__restrict__ cm, int64_t s, int m, int N, int topo, T q)
{
    int order = 5;
    if (topo == LFB_BOUNDED) {
        if (m >= 4 && m <= N - 2) order = 5;
        else if (m >= 3 && m <= N - 1) order = 3;
        else order = 2;
    }
    if (order == 2) {
        if (m == 1) return T(cm[0]);
        if (m == N + 1) return T(cm[-s]);
        return (T(cm[-s]) + T(cm[0])) / T(2.0);
    }
    const bool left = val(q) > 0.0;
    if (order == 3) {
        T o = left ? T(cm[-2 * s]) : T(cm[s]);
        T n = left ? T(cm[-s]) : T(cm[0]);
        T e = left ? T(cm[0]) : T(cm[-s]);
        return weno3z_ref<T>(o, n, e);
    }
    double p = left ? cm[-3 * s] : cm[2 * s];
    double o = left ? cm[-2 * s] : cm[s];
    double n = left ? cm[-s] : cm[0];
    double e = left ? cm[0] : cm[-s];
    double f = left ? cm[s] : cm[-2 * s];
    if (STRICT) return weno5z_ref<T>(T(p), T(o), T(n), T(e), T(f));
    return T(n + weno5z_increment(o - p, n - o, e - n, f - e));
}
