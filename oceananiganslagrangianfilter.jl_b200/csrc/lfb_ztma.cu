// lfb_ztma.cu -- TMA-staged z-marching RK3 stage kernel (production path on 3-D grids: Periodic, Bounded or
// slab-decomposed in x and y, Bounded or Periodic in z).
//
// Same fused work as lfb_stage_generic (WENO5-Z flux divergence, c div(U), forcing, RK3
// substep, tendency caching, periodic halo images; reference call sites: compute_Gc! / tracer_tendency,
// lagrangian_filter_tendency_kernel_functions.jl:36-61, lagrangian_filter_rk3_substep.jl:31-61,
// cache_lagrangian_filter_tendencies.jl:8-31, update_lagrangian_filter_state.jl:33-34).
//
// What bounds this stencil on B200 is the warp schedulers' issue port, not HBM: an FP64 instruction holds the
// sub-partition's port for two cycles and nothing else issues in its shadow (measured: 1 DFMA + 1 FFMA per
// thread cost 3 cycles, scratch microbenchmark quoted in DESIGN.md), so the time of a plane is about
// 2 x (FP64 instructions) + (all other instructions), per warp, summed over the warps of a sub-partition.  The
// kernel is therefore organised to execute as few instructions as possible:
//   * every input tile is brought into shared memory by TMA (cp.async.bulk.tensor + mbarrier): the haloed
//     40 x (TY+6) plane of each tracer, the u / v / w face-velocity tiles, G^- and the forcing variable.  One
//     elected lane issues them two planes ahead; no thread computes a global load address, stages a halo
//     element or publishes a velocity.  The tracer planes live in an 8-deep ring, so the plane that feeds the
//     z window (k+4) and the plane the in-plane stencils read (k) are the same TMA load;
//   * upwinding is done on addresses, not values: the five cells of a face stencil are loaded in upwind order
//     (mirrored shared-memory addressing), so no FP64 select or negation is executed; along z the register
//     window is used by a warp-uniform branch per direction;
//   * the arithmetic per face is short (46 FP64 instructions in x / y, 35 in z): smoothness indicators scaled
//     by 4/3 with eps folded into the curvature term, the blend written as x1 - (A0 y0 / 2 + A2 y2 / 3) / den
//     with y = second differences already at hand, a cubic reciprocal refinement folded into the last FMA,
//     fluxes and divergences in units of face velocity (areas / volume applied once per axis);
//   * the march is unrolled over the 8-plane period of the rings, so every ring offset, flux buffer and
//     mbarrier parity is an immediate; all shared-memory addresses of a thread derive from four registers;
//   * two planes are computed per block barrier.
// One block = the tracers of one (field, coefficient-set) item on a 32 x TY tile, marching over all Nz planes;
// a thread owns one cell column of one tracer and computes its west, south and top faces; the faces on the
// east / north tile edge are computed by helper warps, fluxes are exchanged through shared memory (that
// barrier also orders the reuse of the TMA ring slots).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <type_traits>

#include "lfb_zm_common.cuh"

#define ZT_TX 32
#define ZT_CW 40          // width of the pitch-40 tiles: 4 + 32 + 4 columns (TMA needs a 16-byte aligned inner start
                          // coordinate, so the 3-wide halo is widened to 4)
#define ZT_HX 4
#define ZT_RC 8           // ring of haloed tracer planes: k .. k+4 live, 2 in flight, 1 being released
#define ZT_RI 4           // ring of per-plane input tiles, of flux buffers and of group mbarriers
#define ZT_D  2           // planes the producer runs ahead
#ifndef ZT_L2_PROMOTION
#define ZT_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
#ifndef ZT_PAIR
#define ZT_PAIR 0         // 1: the cosine and the sine tracer of an item share a main thread (see the main warps); measured slower (-8 %)
#endif
#ifndef ZT_PB
#define ZT_PB 2           // planes per block barrier
#endif
#ifndef ZT_TMA_STORE
#define ZT_TMA_STORE 0    // 1: the new state and G^n leave through shared memory and cp.async.bulk.tensor stores (UTMASTG);
                          // measured 2.4 % slower than per-thread STG (profiles/microbench/r02_tma_store_experiment.txt)
#endif
#ifndef ZT_PFORM
#define ZT_PFORM 1        // Z-weight products as P + tau^2 R_s (one FP64 instruction less per face)
#endif
#ifndef ZT_UPS
#define ZT_UPS 1          // upwind stencil addresses as a + s * const (sign mask of the face velocity): 6 integer instructions per face instead of 10
#endif
#ifndef ZT_ZSPLIT
#define ZT_ZSPLIT 1       // z face: separate code for warps whose w is positive in every lane (no dummy value, no select)
#endif
#ifndef ZT_OPAQUE
#define ZT_OPAQUE 2       // 1: the partner tracer's cell address lives in a register instead of being re-selected per plane; 2: the thread's base addresses too
#endif
#ifndef ZT_PF
#define ZT_PF 0           // planes an L2 prefetch (cp.async.bulk.prefetch.tensor) runs ahead; 0 = off: measured slower (6: +2 %, 10: +7 %)
#endif

struct LfbZtParams {
    LfbStageParams S;
    alignas(64) CUtensorMap mC;                  // U_old stack  (x, y, z, tracer), box 40 x (TY+6)
    alignas(64) CUtensorMap mG;                  // G stack      (x, y, z, tracer), box 32 x TY
    alignas(64) CUtensorMap mU;                  // u            (x, y, z),         box 40 x TY
    alignas(64) CUtensorMap mV;                  // v                               box 32 x (TY+1)
    alignas(64) CUtensorMap mW;                  // w                               box 32 x TY x 2 (bottom and top faces)
    alignas(64) CUtensorMap mF[LFB_MAX_VARS];    // original variables, snapshot at or before the stage time, box 32 x TY
    alignas(64) CUtensorMap mF2[LFB_MAX_VARS];   // ... snapshot after it (linear interpolation in the forcing term)
    alignas(64) CUtensorMap mM;                  // relaxation mask (boundary_relaxation), box 32 x TY
    alignas(64) CUtensorMap mK;                  // face codes of an immersed grid, box 34 x (TY+1)
    // store maps (box 32 x TY) of U_new and G for the two row ranges of a launch: their extents end at column Nx and at
    // row j1 of the range, so the surplus columns / rows of partial tiles are clipped by the TMA unit
    alignas(64) CUtensorMap mSU[2];
    alignas(64) CUtensorMap mSG[2];
    double k133, k13, k16, eps43, rd[3];         // 13/3, 1/3, 1/6, 4/3 eps, area / volume per axis
    double dtg, dtz;                             // dt gamma, dt zeta of the stage (constant-bank operands: no register, no per-plane product)
    double *scratch;                             // one field: store target of the rows beyond j1 in partial tiles
};

// Shared memory (doubles).  Two tile geometries: pitch 40 (x halo of 4: tracer planes, u, x fluxes; a thread's
// own cell sits at (ty [+3]) * 40 + tx + 4) and pitch 32 (v, w, G^-, forcing variable, y fluxes; own cell at
// ty * 32 + tx).  Per-tracer tiles are strided by CTP (pitch 40) or VT (pitch 32) doubles.
template <int NSUB, int TY, bool IMMK = false>
struct ZtSmem {
    static constexpr int CH = TY + 6;
    static constexpr int CTP = CH * ZT_CW;                       // haloed tracer tile = the TMA box (multiple of 16)
    static constexpr int NT = ZT_TX * TY;
    static constexpr int UT = TY * ZT_CW;                        // [TY][40]
    static constexpr int VT = (TY + 1) * 32;                     // [TY+1][32]
    static constexpr int OFF_U = 0, OFF_V = UT, OFF_W = OFF_V + VT, OFF_G = OFF_W + 2 * NT, OFF_F = OFF_G + NSUB * VT,
                         OFF_F2 = OFF_F + NT,                    // forcing variable: the two bracketing snapshots
                         OFF_M = OFF_F2 + NT,                    // relaxation mask
                         OFF_K = OFF_M + NT;                     // face codes of an immersed grid: [TY+1][34] (pitch 34), padded to 128 B
    static constexpr int KP = 34, KT = IMMK ? ((TY + 1) * KP + 15) / 16 * 16 : 0;     // only the immersed instantiations carry it
    static constexpr int ISLOT = OFF_K + KT;                     // doubles per input slot
    static constexpr int CSLOT = NSUB * CTP;
    static constexpr int FXT = ZT_TMA_STORE ? UT : CTP;          // x-flux tile of a tracer: [TY][40] (or the tracer-tile geometry)
    static constexpr int FBUF = NSUB * (FXT + VT);               // Fx[NSUB], Fy[NSUB][TY+1][32]
    static constexpr int OUTBUF = ZT_TMA_STORE ? ZT_PB * NSUB * 2 * NT : 0;   // per barrier interval: planes x tracers x (U_new, G^n) tiles [TY][32]
    static constexpr int O_C = 0;
    static constexpr int O_IN = O_C + ZT_RC * CSLOT;
    static constexpr int O_FL = O_IN + ZT_RI * ISLOT;
    static constexpr int O_OUT = O_FL + ZT_RI * FBUF;            // two interval buffers (written while the other one is being stored)
    static constexpr int O_BAR = O_OUT + 2 * OUTBUF;             // ZT_RI group barriers + the prologue barrier
    static constexpr int DOUBLES = O_BAR + 8 + 8;                // + 2 x 4 relaxation coefficients
    static constexpr int BYTES = DOUBLES * 8;
    static constexpr int NS = (ZT_PAIR && NSUB == 2) ? 2 : 1;    // tracers per main thread
    static constexpr int NMAIN = NSUB * NT / NS;
    static constexpr int THREADS = NMAIN + NSUB * 32 + 32;
    static_assert(NSUB * TY < 32, "east-edge faces and the producer lane share one warp");
    static_assert(CTP % 16 == 0 && NT % 16 == 0 && VT % 16 == 0 && UT % 16 == 0, "TMA destinations must stay 128-byte aligned");
    static_assert(BYTES <= 227 * 1024, "shared memory");
};

// ---- mbarrier / TMA -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZT_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZT_DONE_%=;\n"
        "bra ZT_WAIT_%=;\n"
        "ZT_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap *map, int x, int y, int z, int t, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(t), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int x, int y, int z, int t)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(x), "r"(y), "r"(z), "r"(t) : "memory");
}

// shared -> global tensor store (bulk async group), and the fences / waits around it
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, unsigned src, int x, int y, int z, int t)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(x), "r"(y), "r"(z), "r"(t) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// streaming (evict-first) global accesses of the side job: its data is touched once per stage and must not push the
// velocity tiles the sibling blocks of a tile share out of L2
__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(double2 *p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- arithmetic ---------------------------------------------------------------------------------------------------
struct ZtK { double k133, k13, k16, eps43; };

// WENO5-Z blend.  g_s = 4/3 (beta_s + eps) (eps is folded into the curvature terms by the callers), y0 = s0 - s1 and
// y2 = s1 - s2 the second differences of the differences, x1 = dn/3 + do/6 the increment of the central candidate.
// With G_s = g_s^2, tau = g0 - g2, A_s = (G_s + tau^2) prod_{j != s} G_j:
//     face - n = x1 - (A0 y0 / 2 + A2 y2 / 3) / (3 A0 + 6 A1 + A2)
// which is sum alpha_s q_s / sum alpha_s of SURVEY.md App. A.5 brought to one denominator (the common scale 4/3 of
// beta, eps and tau cancels).  One reciprocal: hardware seed (~20 bits) + one cubic step, folded into the last FMA.
__device__ __forceinline__ double zt_blend(double g0, double g1, double g2, double y0, double y2, double x1, const ZtK &K)
{
    const double tau = g0 - g2;
    const double G0 = g0 * g0, G1 = g1 * g1, G2 = g2 * g2;
#if ZT_PFORM
    // A_s = (G_s + tau^2) prod_{j != s} G_j = P + tau^2 R_s with R_s = prod_{j != s} G_j, P = G0 G1 G2: one FP64 instruction less
    const double tt = tau * tau;
    const double R0 = G1 * G2, R1 = G0 * G2, R2 = G0 * G1, P = G0 * R0;
    const double A0 = fma(tt, R0, P), A1 = fma(tt, R1, P), A2 = fma(tt, R2, P);
#else
    const double A0 = fma(tau, tau, G0) * (G1 * G2);
    const double A1 = fma(tau, tau, G1) * (G0 * G2);
    const double A2 = fma(tau, tau, G2) * (G0 * G1);
#endif
    const double den = fma(3.0, A0, fma(6.0, A1, A2));
    const double num = fma(0.5, A0 * y0, K.k13 * (A2 * y2));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    const double e = fma(-den, r, 1.0);
    const double t = fma(e, e, e);
    r = fma(r, t, r);
    return fma(-num, r, x1);
}

// increment (face - n) from the four first differences in upwind order
__device__ __forceinline__ double zt_inc(double dp, double do_, double dn, double de, const ZtK &K)
{
    const double s0 = de - dn, s1 = dn - do_, s2 = do_ - dp;
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double g0 = fma(t0, t0, fma(K.k133 * s0, s0, K.eps43));
    const double g1 = fma(t1, t1, fma(K.k133 * s1, s1, K.eps43));
    const double g2 = fma(t2, t2, fma(K.k133 * s2, s2, K.eps43));
    return zt_blend(g0, g1, g2, s0 - s1, s1 - s2, fma(K.k13, dn, K.k16 * do_), K);
}

// ... with the curvature terms Q_s = 13/3 s_s^2 + 4/3 eps supplied (z direction, register window)
__device__ __forceinline__ double zt_inc_q(double dp, double do_, double dn, double de, double Q2, double Q1, double Q0, const ZtK &K)
{
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double g0 = fma(t0, t0, Q0);
    const double g1 = fma(t1, t1, Q1);
    const double g2 = fma(t2, t2, Q2);
    const double y0 = fma(-2.0, dn, de) + do_, y2 = fma(-2.0, do_, dn) + dp;
    return zt_blend(g0, g1, g2, y0, y2, fma(K.k13, dn, K.k16 * do_), K);
}

// Upwind face value at the low face of the cell at shared byte address `a` along the axis with byte stride S:
// the stencil is read in upwind order (p, o, n | e, f), mirrored when the face velocity is not positive.
// q > 0 decided on the high word (an integer compare instead of a two-cycle FP64 one).  Differs from q > 0.0 only for
// positive q below 2^-1022 * 2^-20 (high word zero), where the flux q * face vanishes anyway.
__device__ __forceinline__ bool zt_positive(double q) { return __double2hiint(q) > 0; }

template <int S>
__device__ __forceinline__ double zt_face(unsigned a, double q, const ZtK &K)
{
#if ZT_UPS
    // sign-mask addressing, see zt_face_up
    const int s = __double2hiint(q) >> 31;
    const double n = lds(a - S - s * S), e = lds(a + s * S), o = lds(a - 2 * S - s * (3 * S));
    const double f = lds(a + S + s * (3 * S)), p = lds(a - 3 * S - s * (5 * S));
#else
    const bool left = zt_positive(q);
    const int step = left ? S : -S;
    const unsigned an = left ? a - S : a;
    const double n = lds(an), e = lds(an + step), o = lds(an - step);
    const double f = lds(an + 2 * step), p = lds(an - 2 * step);
#endif
    return n + zt_inc(o - p, n - o, e - n, f - e, K);
}

// The upwind decision of a face, shared by every tracer advected by the same velocity: byte offset of the upwind cell n
// relative to the cell on the high side of the face, and the signed stride that walks the stencil in upwind order.
#if ZT_UPS
// s = 0 where the face velocity is not negative (high word >= 0: upwind cell on the low side), -1 where it is negative.
// The five loads of the stencil sit at a + s * (k S) + const: one IMAD per load, no select and no register constants.
// (Differs from q > 0 only for q = +0, where the flux q * face vanishes for either stencil.)
struct ZtUp { int s; };
template <int S>
__device__ __forceinline__ ZtUp zt_up(double q)
{
    ZtUp u;
    u.s = __double2hiint(q) >> 31;
    return u;
}
template <int S>
__device__ __forceinline__ double zt_face_up(unsigned a, const ZtUp &u, const ZtK &K)
{
    const int s = u.s;
    const double n = lds(a - S - s * S), e = lds(a + s * S), o = lds(a - 2 * S - s * (3 * S));
    const double f = lds(a + S + s * (3 * S)), p = lds(a - 3 * S - s * (5 * S));
    return n + zt_inc(o - p, n - o, e - n, f - e, K);
}
#else
struct ZtUp { int an, step; };
template <int S>
__device__ __forceinline__ ZtUp zt_up(double q)
{
    const bool left = zt_positive(q);
    ZtUp u;
    u.an = left ? -S : 0; u.step = left ? S : -S;
    return u;
}
template <int S>
__device__ __forceinline__ double zt_face_up(unsigned a, const ZtUp &u, const ZtK &K)
{
    const unsigned an = a + u.an;
    const double n = lds(an), e = lds(an + u.step), o = lds(an - u.step);
    const double f = lds(an + 2 * u.step), p = lds(an - 2 * u.step);
    return n + zt_inc(o - p, n - o, e - n, f - e, K);
}
#endif

// Scheme of the face with 1-based global index m on an axis of N cells: 5 = WENO5, 3 = buffer
// WENO3, 2 = Centered2, 1 / 0 = lower / upper wall face (the value next to the wall; the wall velocity is zero).
// Periodic (or slab-connected) axes and everything outside 1 .. N+1 (surplus lanes of partial tiles): WENO5.
__device__ __forceinline__ int zt_face_order(int m, int N, bool bounded)
{
    if (!bounded || m < 1 || m > N + 1) return 5;
    if (m >= 4 && m <= N - 2) return 5;
    if (m == 1) return 1;
    if (m == N + 1) return 0;
    if (m == 3 || m == N - 1) return 3;
    return 2;
}

// in-plane face next to a wall of a Bounded axis: the same formulas as the generic kernel's production flavour
// (weno3z_ref, lfb_math.cuh).  `a` = shared byte address of the cell on the high side of the face.
template <int S>
__device__ __noinline__ double zt_face_buffer(unsigned a, double q, int order)
{
    const double cm = lds(a), cm1 = lds(a - S);
    if (order == 1) return cm;
    if (order == 0) return cm1;
    if (order == 2) return (cm1 + cm) / 2.0;
    const bool left = q > 0.0;
    const double o = left ? lds(a - 2 * S) : lds(a + S);
    return weno3z_ref<double>(o, left ? cm1 : cm, left ? cm : cm1);
}

// top face of cell k next to the walls (1-based face index m = k + 2 outside 4 .. Nz - 2): WENO3-Z, Centered2, wall
// 
__device__ __noinline__ double zt_top_face_buffer(int m, int Nz, double cw, double d1, double d2, double d3, double ck, double ck1)
{
    if (m == Nz + 1) return ck;
    if (m == 3 || m == Nz - 1) {
        const bool left = cw > 0.0;
        const double dn = d2, do_ = left ? d1 : d3;
        const double eps = LFB_WENO_EPS;
        const double b0 = dn * dn, b1 = do_ * do_;
        const double tau = b0 - b1, tt = tau * tau;
        const double g0 = b0 + eps, g1 = b1 + eps;
        const double G0 = g0 * g0, G1 = g1 * g1;
        const double a0 = 2.0 * ((G0 + tt) * G1), a1 = (G1 + tt) * G0;
        const double inc = div_pos(fma(a0, 0.5 * dn, a1 * (0.5 * do_)), a0 + a1);
        return left ? ck + inc : ck1 - inc;
    }
    return 0.5 * (ck + ck1);
}

// ... on an immersed grid the scheme comes from the face code (3 = WENO3-Z, 2 = Centered2, 0 / 1 = domain wall, 7 = no flux)
__device__ __noinline__ double zt_top_face_coded(int order, double cw, double d1, double d2, double d3, double ck, double ck1)
{
    if (order == 3) return zt_top_face_buffer(3, 1 << 30, cw, d1, d2, d3, ck, ck1);
    if (order == 2) return 0.5 * (ck + ck1);
    return ck;
}

// WALLS: the general instantiation -- Bounded x or y axes (buffer schemes next to those walls), Periodic z and boundary relaxation;
// compiled separately so that the all-periodic form keeps its register allocation.
// FLATY: the folded form for grids with a Flat y axis (LfbStageParams::fold): no y faces, no v tile.
// WALLS = 2: the general instantiation on an immersed grid (whole cells): the scheme of every face and the immersed flag of
// every cell come from the packed code field (LfbStageParams::codes), one more TMA tile per plane.
template <int NSUB, int TY, int WALLS, bool FLATY>
__global__ void __launch_bounds__(ZtSmem<NSUB, TY>::THREADS, 1) lfb_stage_ztma(const __grid_constant__ LfbZtParams P)
{
    constexpr bool IMM = WALLS == 2;
    using SM = ZtSmem<NSUB, TY, WALLS == 2>;
    extern __shared__ __align__(1024) double zt_smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(zt_smem);
    constexpr int ROW_B = 8 * ZT_CW;
    constexpr unsigned CSLOT_B = 8u * SM::CSLOT, ISLOT_B = 8u * SM::ISLOT, FBUF_B = 8u * SM::FBUF, CTP_B = 8u * SM::CTP,
                       VT_B = 8u * SM::VT, NT_B = 8u * SM::NT;
    constexpr unsigned oC = 8u * SM::O_C, oIN = 8u * SM::O_IN, oFL = 8u * SM::O_FL, oBAR = 8u * SM::O_BAR;
    constexpr unsigned oU = oIN + 8u * SM::OFF_U, oV = oIN + 8u * SM::OFF_V, oW = oIN + 8u * SM::OFF_W,
                       oG = oIN + 8u * SM::OFF_G, oF = oIN + 8u * SM::OFF_F, oF2 = oIN + 8u * SM::OFF_F2, oM = oIN + 8u * SM::OFF_M, oK = oIN + 8u * SM::OFF_K;
    constexpr unsigned KROW_B = 8u * SM::KP;
    constexpr unsigned FXT_B = 8u * SM::FXT, UT_B = 8u * SM::UT, oOUT = 8u * SM::O_OUT, OUTBUF_B = 8u * SM::OUTBUF;
    constexpr unsigned oFX = oFL, oFY = oFL + NSUB * FXT_B;
    const LfbStageParams &S = P.S;
    const LfbLayout &L = S.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % S.n_items;
    const int ytile = blockIdx.x / S.n_items;
    const int i0 = (blockIdx.y + S.xt0) * ZT_TX;
    // rows of this tile: range [S.j0, S.j1) first, then (the second edge strip of a slab) [S.jb0, S.jb1)
    const int ytiles_a = (S.j1 - S.j0 + TY - 1) / TY;
    const bool range_b = ytile >= ytiles_a;
    const int j0 = range_b ? S.jb0 + (ytile - ytiles_a) * TY : S.j0 + ytile * TY;
    const int j1 = range_b ? S.jb1 : S.j1;
    const LfbItem item = S.items[item_id];
    const int Nz = L.N[2];
    const bool use_G = !S.first_stage;
    const bool flat_v = FLATY && item.is_map && item.src == 1;      // xi_v on a Flat y axis: forced by the cell-centred v
    const bool item_var = !item.is_map || flat_v;
    const int  src_map = flat_v ? LFB_MAX_VARS - 1 : item.src;      // tensor maps of the forcing source's two snapshots
    const bool relax = WALLS && S.relax;                  // boundary relaxation: only in the general instantiation
    const bool zper = WALLS && L.topo[2] == LFB_PERIODIC; // Periodic z: likewise
    const ZtK K = {P.k133, P.k13, P.k16, P.eps43};
    const unsigned bar0 = sbase + oBAR, barP = bar0 + 8u * ZT_RI;

    if (tid == 0) {
        if (sbase & 127u) __trap();                       // the TMA destinations assume a 128-byte aligned window
        for (int b = 0; b <= ZT_RI; ++b) mbar_init(bar0 + 8u * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= SM::NMAIN) {
        // ============ helper warps: faces on the north / east edge of the tile, and the TMA producer lane ============
        const int hl = tid - SM::NMAIN;
        int sub, tx, ty, kind;                 // kind 1: south face of row TY, 2: west face of column 32, 0: idle, 3: producer
        if (hl < NSUB * 32) { kind = 1; sub = hl / 32; tx = hl % 32; ty = TY; }
        else {
            const int xl = hl - NSUB * 32;
            if (xl < NSUB * TY) { kind = 2; sub = xl / TY; ty = xl % TY; tx = ZT_TX; }
            else { kind = xl == 31 ? 3 : 0; sub = 0; tx = 0; ty = 0; }
        }
        // One loop for the whole warp (bar.sync is warp-aligned): the producer lane issues the TMA loads of the planes
        // ZT_D ahead, the face lanes compute their edge faces, idle lanes only keep the barrier count.
        const int cx = L.xoff + i0 - ZT_HX, cy = L.H[1] + j0 - 3, ix = L.xoff + i0, iy = L.H[1] + j0, z0 = L.H[2];
        const unsigned in_bytes = 8u * (ZT_CW * TY + (FLATY ? 0 : 32 * (TY + 1)) + 2 * 32 * TY) + (use_G ? NSUB * NT_B : 0u) + (item_var ? 2 * NT_B : 0u) + (relax ? NT_B : 0u) +
                                  (IMM ? 8u * SM::KP * (TY + 1) : 0u);
        auto issue_C = [&](int pz, unsigned bar) {            // haloed tracer planes at z index pz (-2 .. Nz+2)
            const unsigned dst = sbase + oC + (unsigned)((pz + 2) & (ZT_RC - 1)) * CSLOT_B;
#pragma unroll
            for (int s = 0; s < NSUB; ++s) tma_load_4d(dst + s * CTP_B, &P.mC, cx, cy, z0 + pz, item.tracer_c + s, bar);
        };
        auto issue_group = [&](int g) {                       // inputs of plane g and tracer plane g + 4
            const unsigned bar = bar0 + 8u * (g & (ZT_RI - 1));
            const bool withC = g + 4 <= Nz + 2;
            mbar_expect_tx(bar, in_bytes + (withC ? NSUB * CTP_B : 0u));
            if (withC) issue_C(g + 4, bar);
            const unsigned is = sbase + (unsigned)(g & (ZT_RI - 1)) * ISLOT_B;
            tma_load_3d(is + oU, &P.mU, cx, iy, z0 + g, bar);
            if (!FLATY) tma_load_3d(is + oV, &P.mV, ix, iy, z0 + g, bar);
            tma_load_3d(is + oW, &P.mW, ix, iy, z0 + g, bar);
            if (use_G) {
#pragma unroll
                for (int s = 0; s < NSUB; ++s) tma_load_4d(is + oG + s * VT_B, &P.mG, ix, iy, z0 + g, item.tracer_c + s, bar);
            }
            if (item_var) {
                tma_load_3d(is + oF, &P.mF[src_map], ix, iy, z0 + g, bar);
                tma_load_3d(is + oF2, &P.mF2[src_map], ix, iy, z0 + g, bar);
            }
            if (relax) tma_load_3d(is + oM, &P.mM, ix, iy, z0 + g, bar);
            if (IMM) tma_load_3d(is + oK, &P.mK, ix, iy, z0 + g, bar);
        };
        // L2 prefetch of the tiles of group g (the shared-memory rings only allow the loads themselves to run one barrier
        // interval ahead; the prefetch runs ZT_PF planes ahead and costs no shared memory)
        auto prefetch_group = [&](int g) {
            if (g + 4 <= Nz + 2) {
#pragma unroll
                for (int s = 0; s < NSUB; ++s) tma_prefetch_4d(&P.mC, cx, cy, z0 + g + 4, item.tracer_c + s);
            }
            if (g >= Nz) return;
            tma_prefetch_3d(&P.mU, cx, iy, z0 + g);
            if (!FLATY) tma_prefetch_3d(&P.mV, ix, iy, z0 + g);
            tma_prefetch_3d(&P.mW, ix, iy, z0 + g);
            if (use_G) {
#pragma unroll
                for (int s = 0; s < NSUB; ++s) tma_prefetch_4d(&P.mG, ix, iy, z0 + g, item.tracer_c + s);
            }
            if (item_var) { tma_prefetch_3d(&P.mF[src_map], ix, iy, z0 + g); tma_prefetch_3d(&P.mF2[src_map], ix, iy, z0 + g); }
        };
        if (kind == 3) {
            mbar_expect_tx(barP, 6u * NSUB * CTP_B);
            for (int pz = -2; pz <= 3; ++pz) issue_C(pz, barP);
            for (int g = 0; g < ZT_D; ++g) issue_group(g);
        }
        __syncwarp();
        const bool face_lane = (kind == 1 && !FLATY) || kind == 2;
        const unsigned aT = sbase + oC + 8u * (sub * SM::CTP + (ty + 3) * ZT_CW + tx + ZT_HX);
        const unsigned aQ = sbase + (kind == 1 ? oV + 8u * (ty * 32 + tx) : oU + 8u * (ty * ZT_CW + tx + ZT_HX));
        const unsigned aF = sbase + (kind == 1 ? oFY + 8u * (sub * SM::VT + ty * 32 + tx)
                                  : ZT_TMA_STORE ? oFX + 8u * (sub * SM::UT + ty * ZT_CW + tx + ZT_HX)
                                                 : oFX + 8u * (sub * SM::CTP + (ty + 3) * ZT_CW + tx + ZT_HX));
        // scheme of this lane's face (lower order next to the walls of a Bounded axis)
        const int h_order = kind == 1 ? zt_face_order(L.goff[1] + j0 + ty + 1, L.NG[1], L.topo[1] == LFB_BOUNDED)
                            : FLATY ? zt_face_order(32 * (j0 + ty) + tx + 1, S.fold_nx, L.topo[0] == LFB_BOUNDED)
                                    : zt_face_order(L.goff[0] + i0 + tx + 1, L.NG[0], L.topo[0] == LFB_BOUNDED);
        // ---- side job: the velocities of the NEXT stage time (S.pf), for the elements the next stage's tiles read ----
        // Work units of a tile plane, one 128-bit access per lane (two doubles), two rows per warp instruction:
        //   units 0 .. 3 TY/2 - 1   rows (2r, 2r + 1) of the tile in u, v, w
        //   unit 3 TY/2             the v row behind the last row of this launch's row range (north face of that row)
        //   unit 3 TY/2 + 1         the u column behind the last column of the grid (east face / periodic image)
        // and one more plane of w (top faces of the last plane).  The units of a tile are dealt round-robin to the helper
        // warps of the blocks (items) of the tile.  The first unit of a warp is software-pipelined by one barrier interval
        // (an HBM round trip is about as long as an interval); further units (few items per tile) follow serially.
        static_assert(TY % 2 == 0, "two rows per side-job unit");
        constexpr int NHW = NSUB + 1, UPA = TY / 2, NUNIT = 3 * UPA + 2;
        const int lane = hl & 31, hw = hl >> 5;
        const bool pf_on = S.pf.on != 0;
        const int unit_step = S.n_items * NHW, unit0 = item_id + S.n_items * hw;
        // the two elements of unit u owned by this lane: axis (-1: none) and offset in plane 0
        auto pf_elem = [&](int u, int &ax, int64_t &off) {
            int i = i0 + 2 * (lane & 15), j;
            bool ok;
            if (u < 3 * UPA)       { ax = u / UPA; j = j0 + 2 * (u % UPA) + (lane >> 4); ok = j < j1 && (ax == 0 ? i <= L.N[0] : i < L.N[0]); }
            else if (u == 3 * UPA) { ax = 1; j = j1; ok = lane < 16 && j0 + TY >= j1 && i < L.N[0]; }
            else                   { ax = 0; i = L.N[0]; j = j0 + lane; ok = lane < TY && j < j1 && i0 + ZT_TX == L.N[0]; }
            if (!ok || u >= NUNIT || !S.pf.dst[ax]) ax = -1;
            off = L.at(i, j, 0);
        };
        auto pf_value = [&](double2 a, double2 b) {
            double2 r;
            r.x = __dadd_rn(__dmul_rn(b.x, S.pf.w2), __dmul_rn(a.x, S.pf.w1));
            r.y = __dadd_rn(__dmul_rn(b.y, S.pf.w2), __dmul_rn(a.y, S.pf.w1));
            if (S.pf.negate) { r.x = -r.x; r.y = -r.y; }
            return r;
        };
        const double2 *pa = nullptr, *pb = nullptr;       // running pointers of the pipelined unit (plane k0)
        double2 *pd = nullptr, *pdl = nullptr;            // ... pdl trails pd by one interval
        if (pf_on) {
            int ax; int64_t off;
            pf_elem(unit0, ax, off);
            if (ax >= 0) {
                pa = reinterpret_cast<const double2 *>(S.pf.s1[ax] + off); pb = reinterpret_cast<const double2 *>(S.pf.s2[ax] + off);
                pd = reinterpret_cast<double2 *>(S.pf.dst[ax] + off);
            }
        }
        const int64_t plane2 = L.plane / 2;                // plane stride in double2 (the pitch is a multiple of 16)
        // units beyond the first of this warp, and the extra w plane: load, combine, store
        auto pf_rest = [&](int k, int u_from, bool w_only) {
            for (int u = u_from; u < NUNIT; u += unit_step) {
                int ax; int64_t off;
                pf_elem(u, ax, off);
                if (ax < 0 || (w_only && ax != 2)) continue;
                off += (int64_t)k * L.plane;
                stg_stream(reinterpret_cast<double2 *>(S.pf.dst[ax] + off),
                           pf_value(ldg_stream(reinterpret_cast<const double2 *>(S.pf.s1[ax] + off)),
                                    ldg_stream(reinterpret_cast<const double2 *>(S.pf.s2[ax] + off))));
            }
        };
        double2 va[ZT_PB], vb[ZT_PB];
#if ZT_TMA_STORE
        // Stores of the finished planes.  The main threads put U_new and G^n of the two planes of interval n (planes 2n,
        // 2n + 1) into out-buffer n & 1 after barrier n and fence them to the async proxy; after barrier n + 1 the
        // producer lane sends the tiles with cp.async.bulk.tensor stores, and waits until the TMA unit has read them
        // before it arrives at barrier n + 2, behind which the buffer is written again.
        const CUtensorMap *mSU = &P.mSU[range_b ? 1 : 0], *mSG = &P.mSG[range_b ? 1 : 0];
        auto store_interval = [&](int n) {
            const unsigned ob = sbase + oOUT + (unsigned)(n & 1) * OUTBUF_B;
#pragma unroll
            for (int p = 0; p < ZT_PB; ++p) {
                const int k = n * ZT_PB + p;
                if (k >= Nz) break;
#pragma unroll
                for (int s = 0; s < NSUB; ++s) {
                    const unsigned tile = ob + (unsigned)((p * NSUB + s) * 2) * NT_B;
                    tma_store_4d(mSU, tile, ix, iy, z0 + k, item.tracer_c + s);
                    if (S.store_G) tma_store_4d(mSG, tile + NT_B, ix, iy, z0 + k, item.tracer_c + s);
                }
            }
            tma_store_commit();
        };
#endif
        if (face_lane) mbar_wait(barP, 0);
        __syncwarp();
        block_barrier();                 // prologue barrier: see the main warps
        for (int k0 = 0; k0 < Nz; k0 += ZT_PB) {
            // the loads of the next barrier interval are issued together at the start of this one, so the consumers
            // can wait for them before this interval's barrier without stalling
#if ZT_TMA_STORE
            if (kind == 3 && k0 >= 2 * ZT_PB) store_interval(k0 / ZT_PB - 2);
#endif
            if (kind == 3) {
#pragma unroll
                for (int p = 0; p < ZT_PB; ++p) if (k0 + ZT_D + p < Nz && (ZT_PB == ZT_D || p == 0)) issue_group(k0 + ZT_D + p);
#if ZT_PF > 0
#pragma unroll
                for (int p = 0; p < ZT_PB; ++p) if (k0 + ZT_PF + p < Nz) prefetch_group(k0 + ZT_PF + p);
#endif
            }
            // side job: store what was loaded during the previous interval, then issue this interval's loads
            if (pf_on && pa) {
#pragma unroll
                for (int p = 0; p < ZT_PB; ++p) {
                    if (k0 > 0) stg_stream(pdl + p * plane2, pf_value(va[p], vb[p]));
                    if (k0 + p < Nz) { va[p] = ldg_stream(pa + p * plane2); vb[p] = ldg_stream(pb + p * plane2); }
                }
                pdl = pd;
                pa += ZT_PB * plane2; pb += ZT_PB * plane2; pd += ZT_PB * plane2;
            }
#pragma unroll
            for (int p = 0; p < ZT_PB; ++p) {
                const int k = k0 + p;
                if (k >= Nz) break;
                if (face_lane) {
                    mbar_wait(bar0 + 8u * (k & (ZT_RI - 1)), (k / ZT_RI) & 1);
                    const unsigned cs = (unsigned)((k + 2) & (ZT_RC - 1)) * CSLOT_B;
                    const double q = lds(aQ + (unsigned)(k & (ZT_RI - 1)) * ISLOT_B);
                    double cf = kind == 1 ? zt_face<ROW_B>(aT + cs, q, K) : zt_face<8>(aT + cs, q, K);
                    if constexpr (IMM) {
                        // south face of row TY / west face of column 32: the code of that cell
                        const int code = (int)lds(sbase + oK + (unsigned)(k & (ZT_RI - 1)) * ISLOT_B + 8u * (ty * SM::KP + tx));
                        const int order = kind == 1 ? (code >> 3) & 7 : code & 7;
                        if (order != 5) cf = kind == 1 ? zt_face_buffer<ROW_B>(aT + cs, q, order) : zt_face_buffer<8>(aT + cs, q, order);
                        sts(aF + (unsigned)(k & (ZT_RI - 1)) * FBUF_B, order == 7 ? 0.0 : q * cf);
                    } else {
                        if (WALLS && h_order != 5) cf = kind == 1 ? zt_face_buffer<ROW_B>(aT + cs, q, h_order) : zt_face_buffer<8>(aT + cs, q, h_order);
                        sts(aF + (unsigned)(k & (ZT_RI - 1)) * FBUF_B, q * cf);
                    }
                }
            }
            __syncwarp();
            if (pf_on && unit0 + unit_step < NUNIT)
                for (int p = 0; p < ZT_PB; ++p) if (k0 + p < Nz) pf_rest(k0 + p, unit0 + unit_step, false);
#if ZT_TMA_STORE
            if (kind == 3) tma_store_wait_read();
#endif
            block_barrier();
        }
#if ZT_TMA_STORE
        {
            const int NI = (Nz + ZT_PB - 1) / ZT_PB;
            if (kind == 3 && NI >= 2) store_interval(NI - 2);
            block_barrier();             // the main threads have written (and fenced) the last interval
            if (kind == 3) { store_interval(NI - 1); tma_store_wait_all(); }
        }
#endif
        if (pf_on) {
            const int kl = (Nz - 1) / ZT_PB * ZT_PB;     // first plane of the last interval
            if (pa) {
#pragma unroll
                for (int p = 0; p < ZT_PB; ++p) if (kl + p < Nz) stg_stream(pdl + p * plane2, pf_value(va[p], vb[p]));
            }
            pf_rest(Nz, unit0, true);                     // top faces of the last plane
        }
        return;
    }

    // ============ main warps: one cell column per thread ============
    // NS tracers per thread.  With ZT_PAIR the cosine and the sine tracer of the item share a thread: they are advected
    // by the same velocities, so the velocity loads, the upwind decisions and addresses, the mbarrier waits and the
    // barrier are executed once for both, the partner value of the coupling term is a register, and two independent
    // FP64 chains per thread cover the FP64 latency with fewer warps.
    constexpr int NS = SM::NS;
    const int t = NS == 2 ? tid : tid % SM::NT, sub0 = NS == 2 ? 0 : tid / SM::NT;
    const int tx = t % ZT_TX, ty = t / ZT_TX;
    // address registers: own cell in the pitch-32 / pitch-40 tiles; the per-tracer tiles sit at constant offsets
#if ZT_OPAQUE >= 2
    // (opaque to the compiler, which otherwise re-derives them from threadIdx in every barrier interval)
    unsigned a32 = sbase + 8u * t, a40 = sbase + 8u * (ty * ZT_CW + tx + ZT_HX);
    unsigned a32s0 = a32 + sub0 * VT_B, a40s0 = a40 + 3 * ROW_B + sub0 * CTP_B;
    asm volatile("" : "+r"(a32), "+r"(a40), "+r"(a32s0), "+r"(a40s0));
#else
    const unsigned a32 = sbase + 8u * t, a40 = sbase + 8u * (ty * ZT_CW + tx + ZT_HX);
    const unsigned a32s0 = a32 + sub0 * VT_B, a40s0 = a40 + 3 * ROW_B + sub0 * CTP_B;
#endif
#define A32S(s) (a32s0 + (unsigned)(s) * VT_B)
#define A40S(s) (a40s0 + (unsigned)(s) * CTP_B)
    // own entry of the x-flux tile of tracer s ([TY][40] with TMA stores, else the tracer-tile geometry)
#define AFX(s) (ZT_TMA_STORE ? a40 + (unsigned)(sub0 + (s)) * UT_B + oFX : A40S(s) + oFX)
    const int partner_off = (NSUB == 2 && NS == 1) ? (sub0 ? -(int)CTP_B : (int)CTP_B) : 0;
    // the partner's cell as an address register of its own (opaque to the compiler, which otherwise re-selects it per plane)
    [[maybe_unused]] unsigned aPart = a40s0 + partner_off;
#if ZT_OPAQUE
    asm volatile("" : "+r"(aPart));
#endif
    // partial tiles at the end of the y range / of x (folded form: the last row of 32 cells may be partial)
    const bool active = j0 + ty < j1 && (FLATY ? 32 * (j0 + ty) + tx < S.fold_nx : i0 + tx < L.N[0]);
    const int64_t plane_b = L.plane * 8;                             // byte stride between planes
    // rows beyond j1 / columns beyond Nx (partial tiles) compute like the others but store into the scratch field
    const int64_t cell0 = L.at(i0 + tx, j0 + ty, 0);
    double *__restrict__ pUn[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) pUn[s] = active ? S.U_new + cell0 + (int64_t)(item.tracer_c + sub0 + s) * L.len : P.scratch + cell0;
    const int64_t g_off = active ? (S.G - S.U_new) : 0;              // pG = pUn + g_off
    // offsets of this column's periodic images in the halo (0 = none); y only when the y halo is not owned by a neighbour rank
    // (32-bit element offsets: a plane has fewer than 2^31 elements, lfb_ztma_supported)
    int xwrap = 0, ywrap = 0;
    if (active) {
        const int i = i0 + tx, j = j0 + ty;
        if (L.fuse_halo[0]) { if (i < L.H[0]) xwrap = L.N[0]; else if (i >= L.N[0] - L.H[0]) xwrap = -L.N[0]; }
        if (L.fuse_halo[1]) {
            if (j < L.H[1]) ywrap = L.N[1] * (int)L.pitch; else if (j >= L.N[1] - L.H[1]) ywrap = -L.N[1] * (int)L.pitch;
        }
    }
    // scheme of the west / south face of this column, and (block-uniform) whether any face of the tile is next to a wall
    // (folded form: the real column of this thread is 32 (j0 + ty) + tx)
    const int ox = FLATY ? zt_face_order(32 * (j0 + ty) + tx + 1, S.fold_nx, L.topo[0] == LFB_BOUNDED)
                         : zt_face_order(L.goff[0] + i0 + tx + 1, L.NG[0], L.topo[0] == LFB_BOUNDED);
    const int oy = FLATY ? 5 : zt_face_order(L.goff[1] + j0 + ty + 1, L.NG[1], L.topo[1] == LFB_BOUNDED);
    const bool tile_walls = FLATY ? (L.topo[0] == LFB_BOUNDED && (j0 == 0 || 32 * (j0 + TY) + 1 >= S.fold_nx - 1))
                                  : ((L.topo[0] == LFB_BOUNDED && (L.goff[0] + i0 < 3 || L.goff[0] + i0 + ZT_TX + 1 >= L.NG[0] - 1)) ||
                                     (L.topo[1] == LFB_BOUNDED && (L.goff[1] + j0 < 3 || L.goff[1] + j0 + TY + 1 >= L.NG[1] - 1)));
    // block-uniform: does any column of this tile have a periodic image?
    const bool tile_wraps = (L.fuse_halo[0] && (i0 < L.H[0] || i0 + ZT_TX > L.N[0] - L.H[0])) ||
                            (L.fuse_halo[1] && (j0 < L.H[1] || j0 + TY > L.N[1] - L.H[1]));
    // forcing = kA * src1 + kB * src2 + (k_own * own + k_oth * partner)   (create_forcing, utils:731-865).  The sources
    // are two shared-memory values: the two faces of the cell for a map (kA = kB carry the 1/2 of the face average),
    // the two bracketing snapshots of the variable for a filtered variable (kA, kB = the weights of the linear
    // interpolation in time of update_input_data!, utils:1036-1058)
    double kA[NS], kB[NS], k_oth[NS];
    const double k_own = -item.c;
    unsigned aS1, aS2;
    {
        const double c = item.c, d = item.d, n2 = c * c + d * d;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int sub = sub0 + s;
            k_oth[s] = NSUB == 1 ? 0.0 : (sub == 0 ? -d : d);
            if (flat_v) {
                // the cell-centred v at the stage time = the linear interpolation of its two bracketing snapshots
                const double coef = S.vsign * ((NSUB == 2 && sub == 1) ? -d / n2 : -c / n2);
                kA[s] = coef * S.var_w1; kB[s] = coef * S.var_w2;
            } else if (item.is_map) {
                kA[s] = kB[s] = 0.5 * ((NSUB == 2 && sub == 1) ? -d / n2 : -c / n2);
            } else {
                kA[s] = sub == 0 ? S.var_w1 : 0.0; kB[s] = sub == 0 ? S.var_w2 : 0.0;
            }
        }
        if (flat_v || !item.is_map) { aS1 = a32 + oF; aS2 = a32 + oF2; }
        else if (item.src == 0)     { aS1 = a40 + oU; aS2 = aS1 + 8; }
        else if (item.src == 1)     { aS1 = a32 + oV; aS2 = aS1 + 8 * 32; }
        else                        { aS1 = a32 + oW; aS2 = aS1 + NT_B; }
    }
    // boundary relaxation (utils:584-699): forcing += (1 / tau_r) mask (g_eq - g), g_eq = kq * src with
    // src = wA src1 + wB src2 the interpolated variable / the cell-centred velocity.  The four coefficients of a tracer
    // sit in shared memory (behind the mbarriers), so that they cost no registers in the march.
    const unsigned aRX0 = sbase + oBAR + 64u + 32u * sub0;
#define ARX(s) (aRX0 + 32u * (unsigned)(s))
    if (relax && t == 0) {
        const double c = item.c, d = item.d, n2 = c * c + d * d;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int sub = sub0 + s;
            double kq;
            if (NSUB == 1)     kq = item.is_map ? -1.0 / (c * c) : 1.0 / c;
            else if (sub == 0) kq = item.is_map ? (d * d - c * c) / (n2 * n2) : c / n2;
            else               kq = item.is_map ? (-2.0 * c * d) / (n2 * n2) : d / n2;
            sts(ARX(s), kq);
            sts(ARX(s) + 8, flat_v ? S.vsign * S.var_w1 : item.is_map ? 0.5 : S.var_w1);
            sts(ARX(s) + 16, flat_v ? S.vsign * S.var_w2 : item.is_map ? 0.5 : S.var_w2);
            sts(ARX(s) + 24, 1.0 / S.relax_timescale);
        }
    }
    const double dtg = P.dtg, dtz = P.dtz;
    const double rdx = P.rd[0], rdy = P.rd[1], rdz = P.rd[2];
    const double nrdx = -rdx, nrdy = -rdy;

    // ---- prologue: the z window of the column -------------------------------------------------------------
    double d0[NS], d1[NS], d2[NS], d3[NS], d4[NS];   // d[k-2..k+2], d[m] = c[m+1] - c[m]
    double Qa[NS], Qb[NS], Qc[NS], Qd[NS];           // Q[k-1..k+2], Q[m] = 13/3 (d[m] - d[m-1])^2 + 4/3 eps
    double ck[NS], ck1[NS], ck3[NS];                 // c[k], c[k+1], c[k+3]
    double Fb[NS], dm3[NS];                          // flux through the bottom face of cell k; Periodic z: d[-3]
    double wb = 0.0;                                 // w at the bottom face of cell k
    mbar_wait(barP, 0);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const unsigned aC = A40S(s) + oC;
        const double cm2 = lds(aC + 0 * CSLOT_B), cm1 = lds(aC + 1 * CSLOT_B);
        Fb[s] = 0.0; dm3[s] = 0.0;
        if (zper) dm3[s] = cm2 - S.U_old[(int64_t)(item.tracer_c + sub0 + s) * L.len + L.at(i0 + tx, j0 + ty, -3)];
        ck[s] = lds(aC + 2 * CSLOT_B); ck1[s] = lds(aC + 3 * CSLOT_B);
        const double ck2 = lds(aC + 4 * CSLOT_B);
        ck3[s] = lds(aC + 5 * CSLOT_B);
        d0[s] = cm1 - cm2; d1[s] = ck[s] - cm1; d2[s] = ck1[s] - ck[s]; d3[s] = ck2 - ck1[s]; d4[s] = ck3[s] - ck2;
        double q;
        q = d1[s] - d0[s]; Qa[s] = fma(K.k133 * q, q, K.eps43);
        q = d2[s] - d1[s]; Qb[s] = fma(K.k133 * q, q, K.eps43);
        q = d3[s] - d2[s]; Qc[s] = fma(K.k133 * q, q, K.eps43);
        q = d4[s] - d3[s]; Qd[s] = fma(K.k133 * q, q, K.eps43);
    }
    // The first in-loop TMA load reuses the ring slot of plane -2: every thread must have read its prologue values,
    // and (main threads pass this barrier only after their wait on barP) the prologue loads must have landed, before
    // the producer lane issues it.
    block_barrier();

    double part[NS][ZT_PB], base[NS][ZT_PB];     // own part of the tendency / of the substep, carried over the barrier
    [[maybe_unused]] int cell_imm[ZT_PB] = {};   // immersed grids: this cell of plane p lies inside the boundary

    // Faces and own tendency part of plane k.  J >= 0: k = 8 m + 2 + J is known modulo the ring period, so every
    // ring offset, flux buffer and mbarrier parity is an immediate and the plane is interior (top face index
    // m = k + 2 in 4 .. Nz - 2: WENO5); J < 0: any plane, everything computed from k.
    auto plane_step = [&](const int k, auto jtag, const int p) {
        constexpr int J = decltype(jtag)::value;
        constexpr bool FAST = J >= 0;
        const int kk = FAST ? 2 + J : k;         // k modulo 8 where it matters
        const unsigned is = (unsigned)(kk & (ZT_RI - 1)) * ISLOT_B;
        const unsigned fb = (unsigned)(kk & (ZT_RI - 1)) * FBUF_B;
        const unsigned oCk = oC + (unsigned)((kk + 2) & (ZT_RC - 1)) * CSLOT_B;
        // velocities and upwind decisions: once for all tracers of the thread
        const double cu = lds(a40 + oU + is), cv = FLATY ? 0.0 : lds(a32 + oV + is);
        const double cw = lds(a32 + oW + NT_B + is);
        const double wb0 = (!FAST && k == 0) ? lds(a32 + oW + is) : wb;
        const ZtUp ux = zt_up<8>(cu), uy = zt_up<ROW_B>(cv);
        const double ue = lds(a40 + oU + is + 8), vn = FLATY ? 0.0 : lds(a32 + oV + is + 8 * 32);
        const double sA = lds(aS1 + is), sB = lds(aS2 + is);
        const bool z_weno = IMM || FAST || zper || (k >= 2 && k <= Nz - 4);      // immersed grids: decided per cell by the code
        const bool left = zt_positive(cw);
        const unsigned ml = z_weno ? __ballot_sync(0xffffffffu, left) : 0u;
        const double dux = ue - cu, dvy = vn - cv, dwz = cw - wb0;
        // immersed grid: schemes of this cell's west / south / bottom / top face and its immersed flag
        int ow = ox, os = oy;
        [[maybe_unused]] int ob = 5, ot = 5;
        bool buffers = WALLS && tile_walls;
        if constexpr (IMM) {
            const int code = (int)lds(sbase + oK + is + 8u * (ty * SM::KP + tx));
            ow = code & 7; os = (code >> 3) & 7; ob = (code >> 6) & 7; ot = (code >> 9) & 7;
            cell_imm[p] = (code >> 12) & 1;
            buffers = true;
        }
        double ckk[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) ckk[s] = ck[s];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const unsigned aCk = A40S(s) + oCk;
            double fx = zt_face_up<8>(aCk, ux, K), fy = FLATY ? 0.0 : zt_face_up<ROW_B>(aCk, uy, K);
            if (buffers) {                                    // buffer schemes next to walls / the immersed boundary
                if (ow != 5) fx = zt_face_buffer<8>(aCk, cu, ow);
                if (!FLATY && os != 5) fy = zt_face_buffer<ROW_B>(aCk, cv, os);
            }
            double Fw = cu * fx, Fs = cv * fy;
            if constexpr (IMM) { if (ow == 7) Fw = 0.0; if (os == 7) Fs = 0.0; }
            sts(AFX(s) + fb, Fw);
            if (!FLATY) sts(A32S(s) + oFY + fb, Fs);
            if (!FAST && k == 0) {
                if (zper) {
                    // Periodic z: the bottom face of cell 0 is a WENO5 face between cells -1 and 0 (c[-1] = c[0] - d[-1])
                    Fb[s] = wb0 * (zt_positive(wb0) ? (ck[s] - d1[s]) + zt_inc(dm3[s], d0[s], d1[s], d2[s], K)
                                                    : ck[s] - zt_inc(d3[s], d2[s], d1[s], d0[s], K));
                } else {
                    Fb[s] = wb0 * ck[s];                     // wall face m = 1: Centered2 on the mirrored halo
                    if constexpr (IMM) { if (ob == 7) Fb[s] = 0.0; }
                }
            }
            double ct;
            bool coded = false;
            if constexpr (IMM) {
                if (ot != 5) { ct = zt_top_face_coded(ot, cw, d1[s], d2[s], d3[s], ck[s], ck1[s]); coded = true; }
            }
            if (coded) {
            } else if (z_weno) {
                // the window is used by a warp-uniform branch per direction (both only in warps that straddle w = 0)
#if ZT_ZSPLIT
                if (ml == 0xffffffffu) ct = ck[s] + zt_inc_q(d0[s], d1[s], d2[s], d3[s], Qa[s], Qb[s], Qc[s], K);
                else {
                    ct = ck1[s] - zt_inc_q(d4[s], d3[s], d2[s], d1[s], Qd[s], Qc[s], Qb[s], K);
                    if (ml != 0u) { const double ctL = ck[s] + zt_inc_q(d0[s], d1[s], d2[s], d3[s], Qa[s], Qb[s], Qc[s], K); ct = left ? ctL : ct; }
                }
#else
                ct = ck1[s];
                if (ml != 0xffffffffu) ct = ck1[s] - zt_inc_q(d4[s], d3[s], d2[s], d1[s], Qd[s], Qc[s], Qb[s], K);
                if (ml != 0u) { const double ctL = ck[s] + zt_inc_q(d0[s], d1[s], d2[s], d3[s], Qa[s], Qb[s], Qc[s], K); ct = left ? ctL : ct; }
#endif
            } else {
                ct = zt_top_face_buffer(k + 2, Nz, cw, d1[s], d2[s], d3[s], ck[s], ck1[s]);
            }
            double Ft = cw * ct;
            if constexpr (IMM) { if (ot == 7) Ft = 0.0; }
            // ---- own part of the tendency: c du - dF per axis in units of face velocity, forcing, G^- part of the substep ----
            double forcing = fma(kA[s], sA, fma(kB[s], sB, k_own * ck[s]));
            if (relax) forcing = fma(lds(a32 + oM + is) * lds(ARX(s) + 24), fma(lds(ARX(s)), fma(lds(ARX(s) + 8), sA, lds(ARX(s) + 16) * sB), -ck[s]), forcing);
            if (NSUB == 2) forcing = fma(k_oth[s], NS == 2 ? ckk[1 - s] : ZT_OPAQUE ? lds(aPart + oCk) : lds(aCk + partner_off), forcing);
            const double az = fma(ck[s], dwz, Fb[s] - Ft);
            part[s][p] = FLATY ? fma(rdx, fma(ck[s], dux, Fw), fma(rdz, az, forcing))
                               : fma(rdx, fma(ck[s], dux, Fw), fma(rdy, fma(ck[s], dvy, Fs), fma(rdz, az, forcing)));
            base[s][p] = S.first_stage ? ck[s] : fma(dtz, lds(A32S(s) + oG + is), ck[s]);
            // ---- advance the window to cell k+1: d[k+3] = c[k+4] - c[k+3] from the ring ----
            {
                Fb[s] = Ft;
                const double c2 = lds(A40S(s) + oC + (unsigned)((kk + 4) & (ZT_RC - 1)) * CSLOT_B);
                const double c4 = lds(A40S(s) + oC + (unsigned)((kk + 6) & (ZT_RC - 1)) * CSLOT_B);
                const double dnew = c4 - ck3[s];
                const double q = dnew - d4[s];
                d0[s] = d1[s]; d1[s] = d2[s]; d2[s] = d3[s]; d3[s] = d4[s]; d4[s] = dnew;
                Qa[s] = Qb[s]; Qb[s] = Qc[s]; Qc[s] = Qd[s]; Qd[s] = fma(K.k133 * q, q, K.eps43);
                ck[s] = ck1[s]; ck1[s] = c2; ck3[s] = c4;
            }
        }
        wb = cw;
    };
    // wait for the TMA group of plane k.  It is called before the barrier of the interval before, so that the code
    // from one barrier to the next wait is a single basic block in which the finish of the planes just synchronised
    // and the loads of the next planes overlap.
    auto wait_plane = [&](const int k, auto jtag) {
        constexpr int J = decltype(jtag)::value;
        const int kk = J >= 0 ? 2 + J : k;
        mbar_wait(bar0 + 8u * (kk & (ZT_RI - 1)), (kk / ZT_RI) & 1);
    };
    // neighbour fluxes of plane k (behind its barrier), RK3 substep, stores
    auto finish_plane = [&](const int k, auto jtag, const int p) {
        constexpr int J = decltype(jtag)::value;
        const int kk = J >= 0 ? 2 + J : k;
        const unsigned fb = (unsigned)(kk & (ZT_RI - 1)) * FBUF_B;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double Fe = lds(AFX(s) + fb + 8);
            double Gn;
            if (FLATY) Gn = fma(nrdx, Fe, part[s][p]);
            else { const double Fn = lds(A32S(s) + oFY + fb + 8 * 32); Gn = fma(nrdx, Fe, fma(nrdy, Fn, part[s][p])); }
            double unew = fma(dtg, Gn, base[s][p]);
            if constexpr (IMM) { if (cell_imm[p]) unew = 0.0; }   // mask_immersed_field! (update_lagrangian_filter_state.jl:22-24)
#if ZT_TMA_STORE
            // out-buffer of this barrier interval: tiles [TY][32] of U_new and G^n per plane and tracer
            {
                const unsigned ob = oOUT + (unsigned)((J >= 0 ? ((2 + J) / ZT_PB) : (k / ZT_PB)) & 1) * OUTBUF_B +
                                    (unsigned)((p * NSUB + sub0 + s) * 2) * NT_B;
                sts(a32 + ob, unew);
                if (S.store_G) sts(a32 + ob + NT_B, Gn);
            }
            const bool images = active && (tile_wraps || (zper && L.fuse_halo[2] && (k < L.H[2] || k >= Nz - L.H[2])));
            double *__restrict__ pu = images ? reinterpret_cast<double *>(reinterpret_cast<char *>(pUn[s]) + (int64_t)k * plane_b) : nullptr;
            if (!images) continue;
#else
            double *__restrict__ pu = pUn[s];
            pu[0] = unew;
            if (S.store_G) pu[g_off] = Gn;
#endif
            if (tile_wraps) {
                // periodic halo images of the new state (the tracer halo fill of update_state!, fused)
                if (xwrap) pu[xwrap] = unew;
                if (ywrap) { pu[ywrap] = unew; if (xwrap) pu[xwrap + ywrap] = unew; }
            }
            if (zper && L.fuse_halo[2] && (k < L.H[2] || k >= Nz - L.H[2])) {
                double *pz = reinterpret_cast<double *>(reinterpret_cast<char *>(pu) + (k < L.H[2] ? Nz : -Nz) * plane_b);
                pz[0] = unew;
                if (tile_wraps) {
                    if (xwrap) pz[xwrap] = unew;
                    if (ywrap) { pz[ywrap] = unew; if (xwrap) pz[xwrap + ywrap] = unew; }
                }
            }
#if !ZT_TMA_STORE
            pUn[s] = reinterpret_cast<double *>(reinterpret_cast<char *>(pu) + plane_b);
#endif
        }
#if ZT_TMA_STORE
        if (p == ZT_PB - 1 || k == Nz - 1) fence_proxy_async();      // the interval's tiles become visible to the TMA unit
#endif
    };
    using Dyn = std::integral_constant<int, -1>;
    // generic planes [ka, kb), ZT_PB per barrier
    auto generic_range = [&](int ka, const int kb) {
#pragma unroll 1
        for (; ka < kb; ka += ZT_PB) {
#pragma unroll
            for (int p = 0; p < ZT_PB; ++p) if (ka + p < Nz) plane_step(ka + p, Dyn(), p);
#pragma unroll
            for (int p = 0; p < ZT_PB; ++p) if (ka + ZT_PB + p < Nz) wait_plane(ka + ZT_PB + p, Dyn());
            block_barrier();
#pragma unroll
            for (int p = 0; p < ZT_PB; ++p) if (ka + p < Nz) finish_plane(ka + p, Dyn(), p);
        }
    };
    static_assert(8 % ZT_PB == 0 && 2 % ZT_PB == 0, "ZT_PB must be 1 or 2");
    int k = 2;
#pragma unroll
    for (int p = 0; p < ZT_PB; ++p) wait_plane(p, Dyn());
    generic_range(0, 2);
    // interior planes in chunks of the ring period: k = 2, 10, 18, ... while the whole chunk is interior
#pragma unroll 1
    for (; k + 7 <= Nz - 4; k += 8) {
#define ZT_PLANE(J) plane_step(k + J, std::integral_constant<int, J>(), J % ZT_PB)
#define ZT_FINISH(J) finish_plane(k + J, std::integral_constant<int, J>(), J % ZT_PB)
#define ZT_WAIT(J) wait_plane(k + J, std::integral_constant<int, J>())
#if ZT_PB == 2
        ZT_PLANE(0); ZT_PLANE(1); ZT_WAIT(2); ZT_WAIT(3); block_barrier(); ZT_FINISH(0); ZT_FINISH(1);
        ZT_PLANE(2); ZT_PLANE(3); ZT_WAIT(4); ZT_WAIT(5); block_barrier(); ZT_FINISH(2); ZT_FINISH(3);
        ZT_PLANE(4); ZT_PLANE(5); ZT_WAIT(6); ZT_WAIT(7); block_barrier(); ZT_FINISH(4); ZT_FINISH(5);
        ZT_PLANE(6); ZT_PLANE(7); ZT_WAIT(8); ZT_WAIT(9); block_barrier(); ZT_FINISH(6); ZT_FINISH(7);
#else
        ZT_PLANE(0); ZT_WAIT(1); block_barrier(); ZT_FINISH(0); ZT_PLANE(1); ZT_WAIT(2); block_barrier(); ZT_FINISH(1);
        ZT_PLANE(2); ZT_WAIT(3); block_barrier(); ZT_FINISH(2); ZT_PLANE(3); ZT_WAIT(4); block_barrier(); ZT_FINISH(3);
        ZT_PLANE(4); ZT_WAIT(5); block_barrier(); ZT_FINISH(4); ZT_PLANE(5); ZT_WAIT(6); block_barrier(); ZT_FINISH(5);
        ZT_PLANE(6); ZT_WAIT(7); block_barrier(); ZT_FINISH(6); ZT_PLANE(7); ZT_WAIT(8); block_barrier(); ZT_FINISH(7);
#endif
#undef ZT_PLANE
#undef ZT_FINISH
#undef ZT_WAIT
    }
    generic_range(k, Nz);
#if ZT_TMA_STORE
    block_barrier();                     // the producer lane stores the last interval behind this barrier
#endif
#undef A32S
#undef A40S
#undef AFX
#undef ARX
}

#ifndef ZT_TY
#define ZT_TY 8
#endif

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*zt_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static zt_encode_fn zt_encoder()
{
    static zt_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (zt_encode_fn)p;
    }
    return fn;
}

// tensor map over one field (rank 3) or a stack of n fields (rank 4) in the padded layout L, box bx x by x bz (x 1).
// fold_rows > 0: the folded view of a grid with a Flat y axis -- rows of 32 cells whose haloed tiles (40 cells) overlap
// in memory (row stride 32 < inner extent 40; profiles/microbench/tma_overlap.cu), fold_rows rows per plane; rows outside
// are zero-filled by the TMA unit.
static int zt_make_map(CUtensorMap *m, const LfbLayout &L, const void *base, int n_stack, int bx, int by, int bz = 1, int fold_rows = 0,
                       int clip_cols = 0, int clip_rows = 0)
{
    zt_encode_fn enc = zt_encoder();
    if (!enc) return -1;
    const cuuint32_t rank = n_stack > 0 ? 4 : 3;
    cuuint64_t dim[4] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)L.planes, (cuuint64_t)(n_stack > 0 ? n_stack : 1)};
    cuuint64_t str[3] = {(cuuint64_t)L.pitch * 8, (cuuint64_t)L.plane * 8, (cuuint64_t)L.len * 8};
    if (fold_rows > 0) { dim[0] = ZT_CW; dim[1] = (cuuint64_t)fold_rows; }
    // store maps: the tensor ends at column clip_cols / row clip_rows, what lies beyond is not written
    if (clip_cols > 0) dim[0] = (cuuint64_t)clip_cols;
    if (clip_rows > 0) dim[1] = (cuuint64_t)clip_rows;
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<void *>(base), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, ZT_L2_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

// Can the z-march kernel run the grid folded (Flat y)?  Periodic or Bounded x; the pitch-40 tile of the last row of 32
// cells must stay inside the plane.
static bool zt_foldable(const LfbLayout &L)
{
    return L.topo[1] == LFB_FLAT && L.topo[0] != LFB_FLAT && L.topo[2] != LFB_FLAT && L.H[0] >= 3 && L.xoff == LFB_XPAD &&
           L.plane % 32 == 0 && (int64_t)32 * ((L.N[0] + 31) / 32) + 24 <= L.plane;
}

extern "C" int lfb_ztma_supported(const LfbStageParams *P)
{
    const LfbLayout &L = P->L;
    if (P->advect != LFB_ADVECTION_WENO5 || P->dzc || (P->relax && !P->mask)) return 0;     // immersed grids: whole cells only
    const bool fold = zt_foldable(L);
    if (!fold && (L.topo[0] == LFB_FLAT || L.topo[1] == LFB_FLAT || L.topo[2] == LFB_FLAT)) return 0;
    if (L.topo[2] == LFB_PERIODIC && L.N[2] < 2 * L.H[2]) return 0;
    if (L.plane >= ((int64_t)1 << 30)) return 0;                   // 32-bit offsets of the periodic images within a plane
    if (!(P->vel_present[0] || P->vel_present[1] || P->vel_present[2])) return 0;     // absent components: a zero field
    if (L.N[0] < 8 || (!fold && L.N[1] < 4) || L.N[2] < 8) return 0;
    if (L.H[0] < 3 || (!fold && L.H[1] < 3) || L.H[2] < 3) return 0;
    if (!zt_encoder()) return 0;
    return fold ? 2 : 1;
}

extern "C" int lfb_ztma_tile_rows(void) { return ZT_TY; }

template <int NSUB, int TY, int WALLS, bool FLATY>
static cudaError_t zt_launch(const LfbStageParams *S, cudaStream_t stream)
{
    using SM = ZtSmem<NSUB, TY, WALLS == 2>;
    // the attribute is per device (a process may hold handles on several GPUs)
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(lfb_stage_ztma<NSUB, TY, WALLS, FLATY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    LfbZtParams P;
    memset(&P, 0, sizeof(P));
    P.S = *S;
    LfbStageParams &SS = P.S;
    int fold_rows = 0;
    if (FLATY) {
        // the folded view: rows of 32 cells, row pitch 32, every pointer moved so that interior column 0 sits at
        // column ZT_HX of the view (element (i, j, k) of the view = real cell (32 j + i, 0, k))
        LfbLayout &F = SS.L;
        const LfbLayout &R = S->L;
        fold_rows = (R.N[0] + 31) / 32;
        SS.fold = 1; SS.fold_nx = R.N[0];
        F.N[0] = 32; F.N[1] = fold_rows; F.NG[0] = 32; F.NG[1] = fold_rows; F.goff[0] = F.goff[1] = 0;
        F.H[1] = 0; F.topo[1] = LFB_PERIODIC;
        F.xoff = ZT_HX; F.pitch = 32; F.rows = (int)(R.plane / 32); F.stride[1] = 32;
        F.fuse_halo[0] = F.fuse_halo[1] = 0;
        const int64_t shift = R.xoff - ZT_HX;
        auto mv = [&](const double *p) { return p ? p + shift : p; };
        SS.U_old = mv(S->U_old); SS.U_new = const_cast<double *>(mv(S->U_new)); SS.G = const_cast<double *>(mv(S->G));
        SS.scratch = const_cast<double *>(mv(S->scratch)); SS.mask = mv(S->mask); SS.codes = mv(S->codes);
        for (int a = 0; a < 3; ++a) {
            SS.vel[a] = mv(S->vel[a]);
            SS.pf.s1[a] = mv(S->pf.s1[a]); SS.pf.s2[a] = mv(S->pf.s2[a]); SS.pf.dst[a] = const_cast<double *>(mv(S->pf.dst[a]));
        }
        SS.pf.dst[1] = nullptr;                                   // v is not advected with on a Flat y axis
        for (int q = 0; q < LFB_MAX_VARS; ++q) { SS.var[q] = mv(S->var[q]); SS.var2[q] = mv(S->var2[q]); }
        SS.j0 = 0; SS.j1 = fold_rows; SS.jb0 = SS.jb1 = 0;
    }
    const LfbLayout &L = SS.L;
    int n_t = 0;
    for (int q = 0; q < S->n_items; ++q) n_t = S->items[q].tracer_c + NSUB > n_t ? S->items[q].tracer_c + NSUB : n_t;
    int bad = 0;
    bad |= zt_make_map(&P.mC, L, SS.U_old, n_t, ZT_CW, TY + 6, 1, fold_rows);
    bad |= zt_make_map(&P.mG, L, SS.G, n_t, 32, TY, 1, fold_rows);
    bad |= zt_make_map(&P.mU, L, SS.vel[0], 0, ZT_CW, TY, 1, fold_rows);
    if (!FLATY) bad |= zt_make_map(&P.mV, L, SS.vel[1], 0, 32, TY + 1);
    bad |= zt_make_map(&P.mW, L, SS.vel[2], 0, 32, TY, 2, fold_rows);
    if (!SS.var_fused) { SS.var_w1 = 1.0; SS.var_w2 = 0.0; }         // inputs already interpolated: one source
    for (int q = 0; q < LFB_MAX_VARS; ++q)
        if (SS.var[q]) {
            bad |= zt_make_map(&P.mF[q], L, SS.var[q], 0, 32, TY, 1, fold_rows);
            bad |= zt_make_map(&P.mF2[q], L, SS.var_fused ? SS.var2[q] : SS.var[q], 0, 32, TY, 1, fold_rows);
        }
    if (S->relax) bad |= zt_make_map(&P.mM, L, SS.mask, 0, 32, TY, 1, fold_rows);
    if (WALLS == 2) bad |= zt_make_map(&P.mK, L, SS.codes, 0, SM::KP, TY + 1, 1, fold_rows);
#if ZT_TMA_STORE
    for (int r = 0; r < 2; ++r) {
        const int jend = r == 0 ? SS.j1 : SS.jb1;
        if (r == 1 && SS.jb1 <= SS.jb0) break;
        const int cols = L.xoff + L.N[0], rows = L.H[1] + jend;
        bad |= zt_make_map(&P.mSU[r], L, SS.U_new, n_t, 32, TY, 1, 0, cols, rows);
        bad |= zt_make_map(&P.mSG[r], L, SS.G, n_t, 32, TY, 1, 0, cols, rows);
    }
#endif
    if (bad) return cudaErrorInvalidValue;
    P.k133 = 13.0 / 3.0; P.k13 = 1.0 / 3.0; P.k16 = 1.0 / 6.0; P.eps43 = LFB_WENO_EPS * (4.0 / 3.0);
    for (int a = 0; a < 3; ++a) P.rd[a] = S->area[a] / S->volume;
    P.dtg = S->dt * S->gamma; P.dtz = S->dt * S->zeta;
    P.scratch = SS.scratch;
    const int ytiles = (SS.j1 - SS.j0 + TY - 1) / TY + (SS.jb1 > SS.jb0 ? (SS.jb1 - SS.jb0 + TY - 1) / TY : 0);
    dim3 grid(S->n_items * ytiles, (!FLATY && S->xt_n > 0) ? S->xt_n : (L.N[0] + ZT_TX - 1) / ZT_TX, 1);
    if (FLATY) SS.xt0 = 0;
    lfb_stage_ztma<NSUB, TY, WALLS, FLATY><<<grid, SM::THREADS, SM::BYTES, stream>>>(P);
    return cudaGetLastError();
}

extern "C" cudaError_t lfb_launch_stage_ztma(const LfbStageParams *P, cudaStream_t stream)
{
    if (P->j1 <= P->j0) return cudaSuccess;
    const bool walls = (P->L.topo[0] == LFB_BOUNDED && !P->plain_cols) || (P->L.topo[1] == LFB_BOUNDED && !P->plain_rows) ||
                       P->L.topo[2] == LFB_PERIODIC || P->relax;                     // the general instantiation
    if (P->codes) {                                    // immersed grid: general instantiation with face codes
        if (P->fold) return P->single_exp ? zt_launch<1, ZT_TY, 2, true>(P, stream) : zt_launch<2, ZT_TY, 2, true>(P, stream);
        return P->single_exp ? zt_launch<1, ZT_TY, 2, false>(P, stream) : zt_launch<2, ZT_TY, 2, false>(P, stream);
    }
    if (P->fold) {
        if (walls) return P->single_exp ? zt_launch<1, ZT_TY, 1, true>(P, stream) : zt_launch<2, ZT_TY, 1, true>(P, stream);
        return P->single_exp ? zt_launch<1, ZT_TY, 0, true>(P, stream) : zt_launch<2, ZT_TY, 0, true>(P, stream);
    }
    if (walls) return P->single_exp ? zt_launch<1, ZT_TY, 1, false>(P, stream) : zt_launch<2, ZT_TY, 1, false>(P, stream);
    return P->single_exp ? zt_launch<1, ZT_TY, 0, false>(P, stream) : zt_launch<2, ZT_TY, 0, false>(P, stream);
}
