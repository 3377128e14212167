// lfb_zm_common.cuh -- device helpers of the z-marching stage kernel (lfb_ztma.cu):
// 32-bit shared-window loads/stores, cp.async, the block barrier, the branch-free division and the
// difference-form WENO5-Z / WENO3-Z face values.
#pragma once

#include <cuda_runtime.h>

#include "lfb_internal.h"
#include "lfb_math.cuh"

#define ZM_TX 32
#define ZM_CW (ZM_TX + 6)

// ---- shared memory through 32-bit shared-window addresses -----------------------------------------
__device__ __forceinline__ double lds(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void cp_async8(unsigned smem_dst, const double *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// WENO3-Z increment from dn = e - n, do_ = n - o (buffer scheme next to a wall)
__device__ __forceinline__ double weno3z_inc(double do_, double dn)
{
    const double eps = LFB_WENO_EPS;
    const double b0 = dn * dn, b1 = do_ * do_;
    const double tau = fabs(b0 - b1);
    const double r0 = tau / (b0 + eps), r1 = tau / (b1 + eps);
    const double a0 = (2.0 / 3.0) * (1.0 + r0 * r0), a1 = (1.0 / 3.0) * (1.0 + r1 * r1);
    return (a0 * (0.5 * dn) + a1 * (0.5 * do_)) / (a0 + a1);
}

__device__ __forceinline__ void block_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// num / den for den > 0 in the normal range: reciprocal seed + two Newton steps (a couple of ulp, far inside
// the 1e-12 parity budget; no special-case branch, unlike the IEEE division sequence)
__device__ __forceinline__ double div_pos(double num, double den)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    double e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    e = fma(-den, r, 1.0);
    r = fma(r, e, r);
    return num * r;
}

// Z weights and candidate increments from the smoothness indicators; returns (face value - n).
__device__ __forceinline__ double zm_from_betas(double b0, double b1, double b2, double dp, double do_, double dn, double de)
{
    const double eps = LFB_WENO_EPS;
    const double x0 = fma(2.0 / 3.0, dn, (-1.0 / 6.0) * de);
    const double x1 = fma(1.0 / 3.0, dn, (1.0 / 6.0) * do_);
    const double x2 = fma(5.0 / 6.0, do_, (-1.0 / 3.0) * dp);
    const double tau = b0 - b2;
    const double tt = tau * tau;
    const double g0 = b0 + eps, g1 = b1 + eps, g2 = b2 + eps;
    const double G0 = g0 * g0, G1 = g1 * g1, G2 = g2 * g2;
    const double A0 = (G0 + tt) * (G1 * G2);
    const double A1 = (G1 + tt) * (G0 * G2);
    const double A2 = (G2 + tt) * (G0 * G1);
    const double den = fma(3.0, A0, fma(6.0, A1, A2));
    const double num = fma(3.0, A0 * x0, fma(6.0, A1 * x1, A2 * x2));
    return div_pos(num, den);
}

// WENO5-Z increment from the four first differences of the upwind-ordered stencil
__device__ __forceinline__ double zm_inc(double dp, double do_, double dn, double de)
{
    const double s0 = de - dn, s1 = dn - do_, s2 = do_ - dp;
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double b0 = fma(3.25 * s0, s0, (0.75 * t0) * t0);
    const double b1 = fma(3.25 * s1, s1, (0.75 * t1) * t1);
    const double b2 = fma(3.25 * s2, s2, (0.75 * t2) * t2);
    return zm_from_betas(b0, b1, b2, dp, do_, dn, de);
}

// ... with the curvature terms Q_s = 13/4 s_s^2 supplied (z direction, register window)
__device__ __forceinline__ double zm_inc_q(double dp, double do_, double dn, double de, double Q2, double Q1, double Q0)
{
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double b0 = fma(0.75 * t0, t0, Q0);
    const double b1 = fma(0.75 * t1, t1, Q1);
    const double b2 = fma(0.75 * t2, t2, Q2);
    return zm_from_betas(b0, b1, b2, dp, do_, dn, de);
}

// A line of the shared tile around a face: `a` is the 32-bit address of the cell on the high side of the
// face (cell m), S the byte stride along the axis.  Loads are issued first (for both in-plane faces) so the
// arithmetic of the two faces can overlap.
struct ZmLine { double m1, m2, p1, far; };

template <int S>
__device__ __forceinline__ ZmLine zm_load_line(unsigned a, bool left)
{
    ZmLine l;
    l.m1 = lds(a - S); l.m2 = lds(a - 2 * S); l.p1 = lds(a + S);
    l.far = lds(left ? a - 3 * S : a + 2 * S);
    return l;
}

// Upwind WENO5-Z face value.  The increment is odd in the differences, so the right-biased value is
// n - inc(mirrored differences) and no operand has to be negated.
//   left : p,o,n,e,f = far,m2,m1,cm,p1      right: p,o,n,e,f = far,p1,cm,m1,m2 (mirrored)
__device__ __forceinline__ double zm_face(const ZmLine &l, double cm, bool left)
{
    const double d_lo = l.m1 - l.m2, d_mid = cm - l.m1, d_hi = l.p1 - cm;      // differences in storage order
    const double d_far = left ? l.m2 - l.far : l.far - l.p1;
    const double inc = zm_inc(d_far, left ? d_lo : d_hi, d_mid, left ? d_hi : d_lo);
    return left ? l.m1 + inc : cm - inc;
}

template <int S>
__device__ __forceinline__ double zm_face_from_tile(unsigned a, double cm, double q)
{
    const bool left = q > 0.0;
    const ZmLine l = zm_load_line<S>(a, left);
    return zm_face(l, cm, left);
}

