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

#include "lfb200.h"

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

// Reconstruction kinds of a scheme chain.  Every Oceananigans scheme carries a buffer scheme that takes over where
// its own stencil would reach outside a Bounded domain or into the immersed boundary (chains, highest order first:
// WENO5 -> WENO3 -> Centered2, UpwindBiased5 -> 3 -> 1, Centered4 -> Centered2; the first one is printed at
// OfflineLagrangianFilter.jl:212-214).
enum { LFB_K_C2, LFB_K_C4, LFB_K_U1, LFB_K_U3, LFB_K_U5, LFB_K_W3, LFB_K_W5 };

// Kind used at face m (1-based; between cells m-1 and m) of an axis of N cells: the first scheme of the chain whose
// symmetric stencil of `buffer` cells on each side of the face (cells m-b .. m+b-1) lies inside a Bounded domain and
// outside the immersed boundary (imm: flag of cell m, or null).  For WENO(order=5) without an immersed boundary this
// is the face-index rule of SURVEY.md App. A.5.
__device__ __forceinline__ int lfb_face_kind(int scheme, int m, int N, int topo, const double *__restrict__ imm, int64_t s)
{
    int kind[3], buf[3], n;
    switch (scheme) {
    case LFB_ADVECTION_WENO3:     kind[0] = LFB_K_W3; buf[0] = 2; kind[1] = LFB_K_C2; buf[1] = 1; n = 2; break;
    case LFB_ADVECTION_CENTERED4: kind[0] = LFB_K_C4; buf[0] = 2; kind[1] = LFB_K_C2; buf[1] = 1; n = 2; break;
    case LFB_ADVECTION_CENTERED2: kind[0] = LFB_K_C2; buf[0] = 1; n = 1; break;
    case LFB_ADVECTION_UPWIND5:   kind[0] = LFB_K_U5; buf[0] = 3; kind[1] = LFB_K_U3; buf[1] = 2; kind[2] = LFB_K_U1; buf[2] = 1; n = 3; break;
    case LFB_ADVECTION_UPWIND3:   kind[0] = LFB_K_U3; buf[0] = 2; kind[1] = LFB_K_U1; buf[1] = 1; n = 2; break;
    case LFB_ADVECTION_UPWIND1:   kind[0] = LFB_K_U1; buf[0] = 1; n = 1; break;
    default:                      kind[0] = LFB_K_W5; buf[0] = 3; kind[1] = LFB_K_W3; buf[1] = 2; kind[2] = LFB_K_C2; buf[2] = 1; n = 3; break;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (c >= n) break;
        const int b = buf[c];
        bool ok = !(topo == LFB_BOUNDED && (m - b < 1 || m + b - 1 > N));
        if (ok && imm)
            for (int r = -b; r < b; ++r) if (imm[r * s] != 0.0) { ok = false; break; }
        if (ok) return kind[c];
    }
    return LFB_K_C2;
}

// Face value at face m (1-based; between cells m-1 and m) along one axis.
//   cm: pointer to cell m; s: stride along the axis; N: global cell count; topo: axis topology;
//   q: advecting face velocity; imm: immersed flag of cell m (or null).
// On the wall faces m = 1, N+1 of a Bounded axis the zero-flux halo cell mirrors the interior, so Centered2 returns
// the interior value (the device layout does not keep Bounded halos).
template <class T, bool STRICT>
__device__ __forceinline__ T face_value(const double *__restrict__ cm, int64_t s, int m, int N, int topo, T q, int scheme,
                                        const double *__restrict__ imm)
{
    const int kind = lfb_face_kind(scheme, m, N, topo, imm, s);
    if (kind == LFB_K_C2) {
        if (topo == LFB_BOUNDED && m == 1) return T(cm[0]);
        if (topo == LFB_BOUNDED && m == N + 1) return T(cm[-s]);
        return (T(cm[-s]) + T(cm[0])) / T(2.0);
    }
    if (kind == LFB_K_C4)
        return T(-1.0 / 12.0) * T(cm[-2 * s]) + T(7.0 / 12.0) * T(cm[-s]) + T(7.0 / 12.0) * T(cm[0]) + T(-1.0 / 12.0) * T(cm[s]);
    // upwind-ordered stencil (p, o, n | e, f): n is the upwind cell
    const bool left = val(q) > 0.0;
    const double n = left ? cm[-s] : cm[0];
    if (kind == LFB_K_U1) return T(n);
    const double o = left ? cm[-2 * s] : cm[s];
    const double e = left ? cm[0] : cm[-s];
    if (kind == LFB_K_W3) return weno3z_ref<T>(T(o), T(n), T(e));
    if (kind == LFB_K_U3) return T(-1.0 / 6.0) * T(o) + T(5.0 / 6.0) * T(n) + T(1.0 / 3.0) * T(e);
    const double p = left ? cm[-3 * s] : cm[2 * s];
    const double f = left ? cm[s] : cm[-2 * s];
    if (kind == LFB_K_U5)
        return T(2.0 / 60.0) * T(p) + T(-13.0 / 60.0) * T(o) + T(47.0 / 60.0) * T(n) + T(27.0 / 60.0) * T(e) + T(-3.0 / 60.0) * T(f);
    if (STRICT) return weno5z_ref<T>(T(p), T(o), T(n), T(e), T(f));
    return T(n + weno5z_increment(o - p, n - o, e - n, f - e));
}
