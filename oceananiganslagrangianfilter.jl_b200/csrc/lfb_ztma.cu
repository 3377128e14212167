// lfb_ztma.cu -- TMA-staged z-marching RK3 stage kernel (production path on 3-D (P|slab, P|slab, B) grids).
//
// Same fused work as lfb_stage_generic / lfb_stage_zmarch (WENO5-Z flux divergence, c div(U), forcing, RK3
// substep, tendency caching, periodic halo images; reference call sites: compute_Gc! / tracer_tendency,
// lagrangian_filter_tendency_kernel_functions.jl:36-61, lagrangian_filter_rk3_substep.jl:31-61,
// cache_lagrangian_filter_tendencies.jl:8-31, update_lagrangian_filter_state.jl:33-34),
// organised around the two limits this stencil has on B200: the FP64 pipe (64 lanes per SM) and the warp
// schedulers' issue slots (an FP64 instruction occupies the pipe for two cycles, so every other instruction
// competes with it one-for-two).  Relative to lfb_stage_zmarch:
//   * every input tile is brought into shared memory by TMA (cp.async.bulk.tensor + mbarrier): the haloed
//     40 x (TY+6) plane of each tracer, the u / v / w face-velocity tiles, G^- and the forcing variable.  One
//     elected lane issues them two planes ahead; no thread computes a global load address, stages a halo
//     element or publishes a velocity any more.  The tracer planes live in an 8-deep ring, so the plane that
//     feeds the z window (k+4) and the plane the in-plane stencils read (k) are the same TMA load;
//   * upwinding is done on addresses, not values: the five cells of a face stencil are loaded in upwind order
//     (mirrored shared-memory addressing), so no FP64 select or negation is executed; along z the register
//     window is used by a warp-uniform branch per direction;
//   * the arithmetic per face is shorter: smoothness indicators scaled by 4/3 (one multiply less each), the
//     blend written as x1 - (A0 y0 / 2 + A2 y2 / 3) / den with y = second differences already at hand, a cubic
//     reciprocal refinement folded into the last FMA, fluxes and divergences in units of face velocity
//     (areas / volume applied once per axis).
// One block = the tracers of one (field, coefficient-set) item on a 32 x TY tile, marching over all Nz planes;
// a thread owns one cell column of one tracer and computes its west, south and top faces; the faces on the
// east / north tile edge are computed by helper warps, fluxes are exchanged through shared memory with one
// bar.sync per plane (that barrier also orders the reuse of the TMA ring slots).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "lfb_zm_common.cuh"

#define ZT_TX 32
#define ZT_CW 40          // haloed tile width: 4 + 32 + 4 columns (TMA needs a 16-byte aligned inner start coordinate, so
                          // the 3-wide halo is widened to 4)
#define ZT_HX 4
#define ZT_UW 34          // u tile width (west + east faces, padded to a 16-byte multiple)
#define ZT_RC 8           // ring of haloed tracer planes: k .. k+4 live, 2 in flight, 1 being released
#define ZT_RI 4           // ring of per-plane input tiles
#define ZT_D  2           // planes the producer runs ahead

struct LfbZtParams {
    LfbStageParams S;
    alignas(64) CUtensorMap mC;                  // U_old stack  (x, y, z, tracer), box 40 x (TY+6)
    alignas(64) CUtensorMap mG;                  // G stack      (x, y, z, tracer), box 32 x TY
    alignas(64) CUtensorMap mU;                  // u            (x, y, z),         box 34 x TY
    alignas(64) CUtensorMap mV;                  // v                               box 32 x (TY+1)
    alignas(64) CUtensorMap mW;                  // w                               box 32 x TY
    alignas(64) CUtensorMap mF[LFB_MAX_VARS];    // original variables              box 32 x TY
    double k133, k13, k16, eps43, rd[3];         // 13/3, 1/3, 1/6, 4/3 eps, area / volume per axis
};

template <int NSUB, int TY>
struct ZtSmem {
    static constexpr int CH = TY + 6;
    static constexpr int CT = CH * ZT_CW;                        // doubles of one haloed tracer tile = the TMA box
    static constexpr int CTP = (CT + 15) / 16 * 16;              // padded to 128 bytes
    static constexpr int NT = ZT_TX * TY;
    static constexpr int UT = (TY * ZT_UW + 15) / 16 * 16;       // [TY][34]
    static constexpr int VT = (TY + 1) * 32;                     // [TY+1][32]
    static constexpr int OFF_U = 0, OFF_V = UT, OFF_W = OFF_V + VT, OFF_G = OFF_W + NT, OFF_F = OFF_G + NSUB * NT;
    static constexpr int ISLOT = OFF_F + NT;                     // doubles per input slot
    static constexpr int CSLOT = NSUB * CTP;
    static constexpr int FBUF = NSUB * (UT + VT);                // Fx[NSUB][TY][34], Fy[NSUB][TY+1][32]
    static constexpr int O_C = 0;
    static constexpr int O_IN = O_C + ZT_RC * CSLOT;
    static constexpr int O_W0 = O_IN + ZT_RI * ISLOT;            // w at the bottom faces of plane 0
    static constexpr int O_FL = O_W0 + NT;
    static constexpr int O_BAR = O_FL + 2 * FBUF;                // ZT_RI group barriers + the prologue barrier
    static constexpr int DOUBLES = O_BAR + 8;
    static constexpr int BYTES = DOUBLES * 8 + 128;              // + slack to align the base to 128 bytes
    static constexpr int NMAIN = NSUB * NT;
    static constexpr int THREADS = NMAIN + NSUB * 32 + 32;
    static_assert(NSUB * TY < 32, "east-edge faces and the producer lane share one warp");
    static_assert(NT % 16 == 0 && VT % 16 == 0, "TMA destinations must stay 128-byte aligned");
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

// ---- arithmetic ---------------------------------------------------------------------------------------------------
struct ZtK { double k133, k13, k16, eps43; };

// WENO5-Z blend.  b_s = 4/3 beta_s, y0 = s0 - s1, y2 = s1 - s2 (second differences of the differences),
// x1 = dn/3 + do/6 the increment of the central candidate.  With A_s = (G_s + tau^2) prod_{j != s} G_j,
// G_s = (b_s + 4/3 eps)^2:   face - n = x1 - (A0 y0 / 2 + A2 y2 / 3) / (3 A0 + 6 A1 + A2)
// (identical to sum alpha_s q_s / sum alpha_s of SURVEY.md App. A.5; the common scale of beta, eps, tau cancels).
__device__ __forceinline__ double zt_blend(double b0, double b1, double b2, double y0, double y2, double x1, const ZtK &K)
{
    const double tau = b0 - b2;
    const double tt = tau * tau;
    const double g0 = b0 + K.eps43, g1 = b1 + K.eps43, g2 = b2 + K.eps43;
    const double G0 = g0 * g0, G1 = g1 * g1, G2 = g2 * g2;
    const double A0 = (G0 + tt) * (G1 * G2);
    const double A1 = (G1 + tt) * (G0 * G2);
    const double A2 = (G2 + tt) * (G0 * G1);
    const double den = fma(3.0, A0, fma(6.0, A1, A2));
    const double num = fma(0.5, A0 * y0, K.k13 * (A2 * y2));
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));       // ~20 bits; one cubic step -> ~60 bits
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
    const double b0 = fma(t0, t0, (K.k133 * s0) * s0);
    const double b1 = fma(t1, t1, (K.k133 * s1) * s1);
    const double b2 = fma(t2, t2, (K.k133 * s2) * s2);
    return zt_blend(b0, b1, b2, s0 - s1, s1 - s2, fma(K.k13, dn, K.k16 * do_), K);
}

// ... with the curvature terms Q_s = 13/3 s_s^2 supplied (z direction, register window)
__device__ __forceinline__ double zt_inc_q(double dp, double do_, double dn, double de, double Q2, double Q1, double Q0, const ZtK &K)
{
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double b0 = fma(t0, t0, Q0);
    const double b1 = fma(t1, t1, Q1);
    const double b2 = fma(t2, t2, Q2);
    const double y0 = fma(-2.0, dn, de) + do_, y2 = fma(-2.0, do_, dn) + dp;
    return zt_blend(b0, b1, b2, y0, y2, fma(K.k13, dn, K.k16 * do_), K);
}

// Upwind face value at the low face of the cell at shared byte address `a` along the axis with byte stride S:
// the stencil is read in upwind order (p, o, n | e, f), mirrored when the face velocity is not positive.
template <int S>
__device__ __forceinline__ double zt_face(unsigned a, double q, const ZtK &K)
{
    const bool left = q > 0.0;
    const int step = left ? S : -S;
    const unsigned an = left ? a - S : a;
    const double n = lds(an), e = lds(an + step), o = lds(an - step);
    const double f = lds(an + 2 * step), p = lds(an - 2 * step);
    return n + zt_inc(o - p, n - o, e - n, f - e, K);
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

template <int NSUB, int TY>
__global__ void __launch_bounds__(ZtSmem<NSUB, TY>::THREADS, 1) lfb_stage_ztma(const __grid_constant__ LfbZtParams P)
{
    using SM = ZtSmem<NSUB, TY>;
    extern __shared__ double zt_smem_raw[];
    const unsigned sbase = ((unsigned)__cvta_generic_to_shared(zt_smem_raw) + 127u) & ~127u;
    constexpr int ROW_B = 8 * ZT_CW;
    constexpr unsigned CSLOT_B = 8u * SM::CSLOT, ISLOT_B = 8u * SM::ISLOT, FBUF_B = 8u * SM::FBUF, CTP_B = 8u * SM::CTP;
    constexpr unsigned oC = 8u * SM::O_C, oIN = 8u * SM::O_IN, oW0 = 8u * SM::O_W0, oFL = 8u * SM::O_FL, oBAR = 8u * SM::O_BAR;
    constexpr unsigned oU = oIN + 8u * SM::OFF_U, oV = oIN + 8u * SM::OFF_V, oW = oIN + 8u * SM::OFF_W,
                       oG = oIN + 8u * SM::OFF_G, oF = oIN + 8u * SM::OFF_F;
    constexpr unsigned oFX = oFL, oFY = oFL + 8u * NSUB * SM::UT;
    constexpr unsigned CBOX_B = 8u * SM::CT, NT_B = 8u * SM::NT;
    const LfbStageParams &S = P.S;
    const LfbLayout &L = S.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % S.n_items;
    const int ytile = blockIdx.x / S.n_items;
    const int i0 = blockIdx.y * ZT_TX;
    const int j0 = S.j0 + ytile * TY;
    const LfbItem item = S.items[item_id];
    const int Nz = L.N[2];
    const bool use_G = !S.first_stage;
    const bool item_var = !item.is_map;
    const ZtK K = {P.k133, P.k13, P.k16, P.eps43};
    const unsigned bar0 = sbase + oBAR, barP = bar0 + 8u * ZT_RI;

    if (tid == 0) {
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
        // One loop for the whole warp (bar.sync is warp-aligned): the producer lane issues the TMA loads of plane
        // k + ZT_D, the face lanes compute their edge face of plane k, idle lanes only keep the barrier count.
        const int cx = L.xoff + i0 - ZT_HX, cy = L.H[1] + j0 - 3, ix = L.xoff + i0, iy = L.H[1] + j0, z0 = L.H[2];
        const unsigned in_bytes = 8u * (ZT_UW * TY + 32 * (TY + 1) + 32 * TY) + (use_G ? NSUB * NT_B : 0u) + (item_var ? NT_B : 0u);
        auto issue_C = [&](int pz, unsigned bar) {            // haloed tracer planes at z index pz (-2 .. Nz+2)
            const unsigned dst = sbase + oC + (unsigned)((pz + 2) & (ZT_RC - 1)) * CSLOT_B;
#pragma unroll
            for (int s = 0; s < NSUB; ++s) tma_load_4d(dst + s * CTP_B, &P.mC, cx, cy, z0 + pz, item.tracer_c + s, bar);
        };
        auto issue_group = [&](int g) {                       // inputs of plane g and tracer plane g + 4
            const unsigned bar = bar0 + 8u * (g & (ZT_RI - 1));
            const bool withC = g + 4 <= Nz + 2;
            mbar_expect_tx(bar, in_bytes + (withC ? NSUB * CBOX_B : 0u));
            if (withC) issue_C(g + 4, bar);
            const unsigned is = sbase + (unsigned)(g & (ZT_RI - 1)) * ISLOT_B;
            tma_load_3d(is + oU, &P.mU, ix, iy, z0 + g, bar);
            tma_load_3d(is + oV, &P.mV, ix, iy, z0 + g, bar);
            tma_load_3d(is + oW, &P.mW, ix, iy, z0 + g + 1, bar);
            if (use_G) {
#pragma unroll
                for (int s = 0; s < NSUB; ++s) tma_load_4d(is + oG + s * NT_B, &P.mG, ix, iy, z0 + g, item.tracer_c + s, bar);
            }
            if (item_var) tma_load_3d(is + oF, &P.mF[item.src], ix, iy, z0 + g, bar);
        };
        if (kind == 3) {
            mbar_expect_tx(barP, 6u * NSUB * CBOX_B + NT_B);
            for (int pz = -2; pz <= 3; ++pz) issue_C(pz, barP);
            tma_load_3d(sbase + oW0, &P.mW, ix, iy, z0, barP);
            for (int g = 0; g < ZT_D; ++g) issue_group(g);
        }
        __syncwarp();
        const bool face_lane = kind == 1 || kind == 2;
        const unsigned aT = sbase + oC + 8u * (sub * SM::CTP + (ty + 3) * ZT_CW + tx + ZT_HX);
        const unsigned aQ = sbase + (kind == 1 ? oV + 8u * (ty * 32 + tx) : oU + 8u * (ty * ZT_UW + tx));
        const unsigned aF = sbase + (kind == 1 ? oFY + 8u * (sub * SM::VT + ty * 32 + tx) : oFX + 8u * (sub * SM::UT + ty * ZT_UW + tx));
        if (face_lane) mbar_wait(barP, 0);
        __syncwarp();
        block_barrier();                 // prologue barrier: see the main warps
        for (int k = 0; k < Nz; ++k) {
            if (kind == 3) {
                if (k + ZT_D < Nz) issue_group(k + ZT_D);
            } else if (face_lane) {
                mbar_wait(bar0 + 8u * (k & (ZT_RI - 1)), (k / ZT_RI) & 1);
                const unsigned cs = (unsigned)((k + 2) & (ZT_RC - 1)) * CSLOT_B;
                const double q = lds(aQ + (unsigned)(k & (ZT_RI - 1)) * ISLOT_B);
                const double cf = kind == 1 ? zt_face<ROW_B>(aT + cs, q, K) : zt_face<8>(aT + cs, q, K);
                sts(aF + (unsigned)(k & 1) * FBUF_B, q * cf);
            }
            __syncwarp();
            block_barrier();
        }
        return;
    }

    // ============ main warps: one cell column of one tracer per thread ============
    const int t = tid % SM::NT, sub = tid / SM::NT;
    const int tx = t % ZT_TX, ty = t / ZT_TX;
    const unsigned aC = sbase + oC + 8u * (sub * SM::CTP + (ty + 3) * ZT_CW + tx + ZT_HX);     // own cell in a tracer tile
    const unsigned a32 = sbase + 8u * t;                                                    // pitch-32 tiles: v, w, G, f, Fy
    const unsigned a34 = sbase + 8u * (ty * ZT_UW + tx);                                    // pitch-34 tiles: u, Fx
    const int partner_off = NSUB == 2 ? (sub ? -(int)CTP_B : (int)CTP_B) : 0;
    const bool active = j0 + ty < S.j1;                              // partial tiles at the end of the y range
    const int64_t off0 = L.at(i0 + tx, j0 + ty, 0) + (int64_t)(item.tracer_c + sub) * L.len;
    double *__restrict__ pUn = S.U_new + off0;
    double *__restrict__ pG = S.G + off0;
    const int64_t plane = L.plane;
    const bool need_var = item_var && sub == 0;
    // offsets of this column's periodic images in the halo (0 = none); y only when the y halo is not owned by a neighbour rank
    int64_t xwrap = 0, ywrap = 0;
    {
        const int i = i0 + tx, j = j0 + ty;
        if (L.fuse_halo[0]) { if (i < L.H[0]) xwrap = L.N[0]; else if (i >= L.N[0] - L.H[0]) xwrap = -L.N[0]; }
        if (L.fuse_halo[1]) {
            if (j < L.H[1]) ywrap = (int64_t)L.N[1] * L.pitch; else if (j >= L.N[1] - L.H[1]) ywrap = -(int64_t)L.N[1] * L.pitch;
        }
    }
    // forcing = k_src * src + (k_own * own + k_oth * partner)   (create_forcing, utils:731-865)
    double k_src, k_own, k_oth;
    {
        const double c = item.c, d = item.d, n2 = c * c + d * d;
        k_own = -c;
        k_oth = NSUB == 1 ? 0.0 : (sub == 0 ? -d : d);
        if (item.is_map) k_src = (NSUB == 2 && sub == 1) ? -d / n2 : -c / n2;
        else             k_src = sub == 0 ? 1.0 : 0.0;
    }
    const int map_axis = item.is_map ? item.src : -1;
    const double dtg = S.dt * S.gamma, dtz = S.dt * S.zeta;
    const double rdx = P.rd[0], rdy = P.rd[1], rdz = P.rd[2];

    // ---- prologue: the z window of the column -------------------------------------------------------------
    double d0, d1, d2, d3, d4;                   // d[k-2..k+2], d[m] = c[m+1] - c[m]
    double Qa, Qb, Qc, Qd;                       // Q[k-1..k+2], Q[m] = 13/3 (d[m] - d[m-1])^2
    double ck, ck1;                              // c[k], c[k+1]
    double Fb, wb;
    mbar_wait(barP, 0);
    {
        const double cm2 = lds(aC + 0 * CSLOT_B), cm1 = lds(aC + 1 * CSLOT_B);
        ck = lds(aC + 2 * CSLOT_B); ck1 = lds(aC + 3 * CSLOT_B);
        const double ck2 = lds(aC + 4 * CSLOT_B), ck3 = lds(aC + 5 * CSLOT_B);
        d0 = cm1 - cm2; d1 = ck - cm1; d2 = ck1 - ck; d3 = ck2 - ck1; d4 = ck3 - ck2;
        double s;
        s = d1 - d0; Qa = (K.k133 * s) * s;
        s = d2 - d1; Qb = (K.k133 * s) * s;
        s = d3 - d2; Qc = (K.k133 * s) * s;
        s = d4 - d3; Qd = (K.k133 * s) * s;
        wb = lds(a32 + oW0);
        Fb = wb * ck;                                    // wall face m = 1: Centered2 on the mirrored halo
    }
    // The first in-loop TMA load reuses the ring slot of plane -2: every thread must have read its prologue values,
    // and (main threads pass this barrier only after their wait on barP) the prologue loads must have landed, before
    // the producer lane issues it.
    block_barrier();

#pragma unroll 2
    for (int k = 0; k < Nz; ++k) {
        const unsigned is = (unsigned)(k & (ZT_RI - 1)) * ISLOT_B;
        const unsigned cs = (unsigned)((k + 2) & (ZT_RC - 1)) * CSLOT_B;
        const unsigned fb = (unsigned)(k & 1) * FBUF_B;
        mbar_wait(bar0 + 8u * (k & (ZT_RI - 1)), (k / ZT_RI) & 1);
        // ---- faces of plane k ----------------------------------------------------------------------------
        const double cu = lds(a34 + oU + is), cv = lds(a32 + oV + is), cw = lds(a32 + oW + is);
        double other = 0.0;
        if (NSUB == 2) other = lds(aC + cs + partner_off);
        const double Fw = cu * zt_face<8>(aC + cs, cu, K);
        const double Fs = cv * zt_face<ROW_B>(aC + cs, cv, K);
        sts(a34 + oFX + 8u * sub * SM::UT + fb, Fw);
        sts(a32 + oFY + 8u * sub * SM::VT + fb, Fs);
        double ct;
        if (k >= 2 && k <= Nz - 4) {                     // top face of cell k: 1-based face index m = k + 2 in 4 .. Nz - 2
            const bool left = cw > 0.0;
            const unsigned ml = __ballot_sync(0xffffffffu, left);
            double incL = 0.0, incR = 0.0;
            if (ml != 0u) incL = zt_inc_q(d0, d1, d2, d3, Qa, Qb, Qc, K);
            if (ml != 0xffffffffu) incR = zt_inc_q(d4, d3, d2, d1, Qd, Qc, Qb, K);
            ct = left ? ck + incL : ck1 - incR;
        } else {
            ct = zt_top_face_buffer(k + 2, Nz, cw, d1, d2, d3, ck, ck1);
        }
        const double Ft = cw * ct;
        block_barrier();

        // ---- tendency, RK3 substep, advance the window -------------------------------------------------
        {
            const double Fe = lds(a34 + oFX + 8u * sub * SM::UT + fb + 8), Fn = lds(a32 + oFY + 8u * sub * SM::VT + fb + 8 * 32);
            const double ue = lds(a34 + oU + is + 8), vn = lds(a32 + oV + is + 8 * 32);
            const double ax = fma(ck, ue - cu, Fw - Fe);          // c du - dF per axis, in units of face velocity
            const double ay = fma(ck, vn - cv, Fs - Fn);
            const double az = fma(ck, cw - wb, Fb - Ft);
            double src;
            if (map_axis >= 0) src = map_axis == 0 ? 0.5 * (cu + ue) : (map_axis == 1 ? 0.5 * (cv + vn) : 0.5 * (wb + cw));
            else src = need_var ? lds(a32 + oF + is) : 0.0;
            const double forcing = fma(k_src, src, fma(k_own, ck, k_oth * other));
            const double Gn = fma(rdx, ax, fma(rdy, ay, fma(rdz, az, forcing)));
            double unew;
            if (S.first_stage) unew = fma(dtg, Gn, ck);
            else               unew = fma(dtg, Gn, fma(dtz, lds(a32 + oG + 8u * sub * SM::NT + is), ck));
            if (active) {
                pUn[0] = unew;
                // periodic halo images of the new state (the tracer halo fill of update_state!, fused)
                if (xwrap) pUn[xwrap] = unew;
                if (ywrap) { pUn[ywrap] = unew; if (xwrap) pUn[xwrap + ywrap] = unew; }
                if (S.store_G) pG[0] = Gn;
            }
            // advance the window to cell k+1: d[k+3] = c[k+4] - c[k+3] from the ring
            Fb = Ft; wb = cw;
            const double c3 = lds(aC + (unsigned)((k + 5) & (ZT_RC - 1)) * CSLOT_B);
            const double c4 = lds(aC + (unsigned)((k + 6) & (ZT_RC - 1)) * CSLOT_B);
            const double dnew = c4 - c3;
            const double s = dnew - d4;
            d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = dnew;
            Qa = Qb; Qb = Qc; Qc = Qd; Qd = (K.k133 * s) * s;
            ck = ck1;
            ck1 = lds(aC + (unsigned)((k + 4) & (ZT_RC - 1)) * CSLOT_B);  // c[k+2]
        }
        pUn += plane; pG += plane;
    }
}

#ifndef ZT_TY
#define ZT_TY 6
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

// tensor map over one field (rank 3) or a stack of n fields (rank 4) in the padded layout L, box bx x by x 1 (x 1)
static int zt_make_map(CUtensorMap *m, const LfbLayout &L, const void *base, int n_stack, int bx, int by)
{
    zt_encode_fn enc = zt_encoder();
    if (!enc) return -1;
    const cuuint32_t rank = n_stack > 0 ? 4 : 3;
    cuuint64_t dim[4] = {(cuuint64_t)L.pitch, (cuuint64_t)L.rows, (cuuint64_t)L.planes, (cuuint64_t)(n_stack > 0 ? n_stack : 1)};
    cuuint64_t str[3] = {(cuuint64_t)L.pitch * 8, (cuuint64_t)L.plane * 8, (cuuint64_t)L.len * 8};
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<void *>(base), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

extern "C" int lfb_ztma_supported(const LfbStageParams *P)
{
    const LfbLayout &L = P->L;
    if (!P->advect || P->relax) return 0;
    if (L.topo[0] != LFB_PERIODIC || L.topo[2] != LFB_BOUNDED) return 0;
    if (L.topo[1] != LFB_PERIODIC && !(L.connected_lo[1] && L.connected_hi[1])) return 0;
    if (!(P->vel_present[0] && P->vel_present[1] && P->vel_present[2])) return 0;
    if (L.N[0] % ZT_TX || L.N[1] < 4 || L.N[2] < 8) return 0;
    if (L.H[0] < 3 || L.H[1] < 3 || L.H[2] < 3) return 0;
    if (!zt_encoder()) return 0;
    return 1;
}

extern "C" int lfb_ztma_tile_rows(void) { return ZT_TY; }

template <int NSUB, int TY>
static cudaError_t zt_launch(const LfbStageParams *S, cudaStream_t stream)
{
    using SM = ZtSmem<NSUB, TY>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lfb_stage_ztma<NSUB, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    LfbZtParams P;
    memset(&P, 0, sizeof(P));
    P.S = *S;
    const LfbLayout &L = S->L;
    int n_t = 0;
    for (int q = 0; q < S->n_items; ++q) n_t = S->items[q].tracer_c + NSUB > n_t ? S->items[q].tracer_c + NSUB : n_t;
    int bad = 0;
    bad |= zt_make_map(&P.mC, L, S->U_old, n_t, ZT_CW, TY + 6);
    bad |= zt_make_map(&P.mG, L, S->G, n_t, 32, TY);
    bad |= zt_make_map(&P.mU, L, S->vel[0], 0, ZT_UW, TY);
    bad |= zt_make_map(&P.mV, L, S->vel[1], 0, 32, TY + 1);
    bad |= zt_make_map(&P.mW, L, S->vel[2], 0, 32, TY);
    for (int q = 0; q < LFB_MAX_VARS; ++q)
        if (S->var[q]) bad |= zt_make_map(&P.mF[q], L, S->var[q], 0, 32, TY);
    if (bad) return cudaErrorInvalidValue;
    P.k133 = 13.0 / 3.0; P.k13 = 1.0 / 3.0; P.k16 = 1.0 / 6.0; P.eps43 = LFB_WENO_EPS * (4.0 / 3.0);
    for (int a = 0; a < 3; ++a) P.rd[a] = S->area[a] / S->volume;
    const int ytiles = (S->j1 - S->j0 + TY - 1) / TY;
    dim3 grid(S->n_items * ytiles, L.N[0] / ZT_TX, 1);
    lfb_stage_ztma<NSUB, TY><<<grid, SM::THREADS, SM::BYTES, stream>>>(P);
    return cudaGetLastError();
}

extern "C" cudaError_t lfb_launch_stage_ztma(const LfbStageParams *P, cudaStream_t stream)
{
    if (P->j1 <= P->j0) return cudaSuccess;
    return P->single_exp ? zt_launch<1, ZT_TY>(P, stream) : zt_launch<2, ZT_TY>(P, stream);
}
