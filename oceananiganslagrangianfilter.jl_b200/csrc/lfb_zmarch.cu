// lfb_zmarch.cu -- the production RK3 stage kernel for 3-D grids that are Periodic (or slab-connected)
// in x and y and Bounded in z: shared-memory x-y tiles marching along z.
//
// Same fused work as lfb_stage_generic (WENO5-Z flux divergence, c div(U), forcing, RK3 substep,
// tendency caching; reference call sites listed there), reorganised around the FP64 pipe, which bounds
// this kernel on B200 ("roofline"):
//   * every face value is computed ONCE: a thread owns one cell column of one tracer and computes its
//     west, south and top face; the east / north fluxes come from the neighbouring threads through
//     shared memory, the bottom flux is carried in a register along the march;
//   * the extra faces on the east / north edge of a tile are packed into three helper warps instead
//     of diverging lanes of the main warps;
//   * along z the first differences d and the curvature terms Q = 13/4 (d_m - d_{m-1})^2 of the
//     smoothness indicators live in a register window and are computed once per cell (both upwind
//     orientations and both neighbouring faces reuse them);
//   * one division per face (common-denominator form of the Z weights, lfb_math.cuh), no negations
//     or sign flips on the FP64 pipe (the increment is odd in the differences).
// One block handles the two tracers (cosine, sine) of one (field, coefficient-set) item on a 32 x 8
// tile.  x-y neighbours are read from a double-buffered tile with a 3-wide halo; the halo of plane k+1
// is brought in with cp.async while plane k is computed, the fluxes are double-buffered too, so one
// __syncthreads per plane suffices.  Global loads for plane k+1 are issued before the arithmetic of
// plane k.  Blocks of the same tile (different items) are adjacent in launch order so the velocity
// planes they share are served by L2, and y-neighbouring tiles run in the same wave so their halo rows
// are L2 hits.
#include <cuda_runtime.h>

#include "lfb_internal.h"
#include "lfb_math.cuh"

#define ZM_TX 32
#define ZM_TY 8
#define ZM_MAIN (ZM_TX * ZM_TY)
#define ZM_CW (ZM_TX + 6)
#define ZM_CH (ZM_TY + 6)

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

// Z weights and candidate increments from the smoothness indicators; returns (face value - n).
__device__ __forceinline__ double weno5z_from_betas(double b0, double b1, double b2, double dp, double do_, double dn,
                                                    double de)
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
    return num / den;
}

// WENO5-Z increment with the curvature terms Q_s = 13/4 s_s^2 supplied (z direction)
__device__ __forceinline__ double weno5z_inc_q(double dp, double do_, double dn, double de, double Q2, double Q1, double Q0)
{
    const double t0 = fma(-3.0, dn, de), t1 = dn + do_, t2 = fma(3.0, do_, -dp);
    const double b0 = fma(0.75 * t0, t0, Q0);
    const double b1 = fma(0.75 * t1, t1, Q1);
    const double b2 = fma(0.75 * t2, t2, Q2);
    return weno5z_from_betas(b0, b1, b2, dp, do_, dn, de);
}

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

// Upwind WENO5-Z face value from a line of the shared tile: `a` is the 32-bit address of the cell on the
// high side of the face (cell m), `s` the byte stride along the axis, cm its value, q the face velocity.
__device__ __forceinline__ double face_from_tile(unsigned a, int s, double cm, double q)
{
    const bool left = q > 0.0;
    const double m1 = lds(a - s), m2 = lds(a - 2 * s), p1 = lds(a + s);
    const double far = lds(left ? a - 3 * s : a + 2 * s);
    const double n = left ? m1 : cm;
    const double e = left ? cm : m1;
    const double o = left ? m2 : p1;
    const double f = left ? p1 : m2;
    return n + weno5z_increment(o - far, n - o, e - n, f - e);
}

// shared memory carve-up (doubles); NSUB tracers per block
template <int NSUB>
struct ZmSmem {
    static constexpr int C_SUB = ZM_CH * ZM_CW;                 // one tracer plane with halo
    static constexpr int C_BUF = NSUB * C_SUB;
    static constexpr int FX_SUB = ZM_TY * (ZM_TX + 1);
    static constexpr int FY_SUB = (ZM_TY + 1) * ZM_TX;
    static constexpr int OFF_C = 0;                             // C[2][NSUB][CH][CW]
    static constexpr int OFF_FX = OFF_C + 2 * C_BUF;            // Fx[2][NSUB][TY][TX+1]
    static constexpr int OFF_FY = OFF_FX + 2 * NSUB * FX_SUB;   // Fy[2][NSUB][TY+1][TX]
    static constexpr int OFF_QX = OFF_FY + 2 * NSUB * FY_SUB;   // Qx[2][TY][TX+1]
    static constexpr int OFF_QY = OFF_QX + 2 * FX_SUB;          // Qy[2][TY+1][TX]
    static constexpr int TOTAL = OFF_QY + 2 * FY_SUB;
};

template <int NSUB>
__global__ void __launch_bounds__(NSUB *ZM_MAIN + NSUB * 32 + 32, 1) lfb_stage_zmarch(const __grid_constant__ LfbStageParams P)
{
    using SM = ZmSmem<NSUB>;
    extern __shared__ __align__(16) double zm_smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(zm_smem);
    const LfbLayout &L = P.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % P.n_items;
    const int ytile = blockIdx.x / P.n_items;
    const int i0 = blockIdx.y * ZM_TX;
    const int j0 = P.j0 + ytile * ZM_TY;
    const LfbItem item = P.items[item_id];
    const int Nz = L.N[2];
    const int64_t plane = L.plane;

    // ---- roles ----------------------------------------------------------------------------------
    const bool is_main = tid < NSUB * ZM_MAIN;
    int sub, tx, ty;
    int hkind = 0;                               // helpers: 1 = extra y face (north edge), 2 = extra x face (east edge)
    if (is_main) {
        tx = tid % ZM_TX; ty = (tid / ZM_TX) % ZM_TY; sub = tid / ZM_MAIN;
    } else {
        const int hl = tid - NSUB * ZM_MAIN;
        if (hl < NSUB * 32) { hkind = 1; sub = hl / 32; tx = hl % 32; ty = ZM_TY; }
        else {
            const int xl = hl - NSUB * 32;
            if (xl < NSUB * ZM_TY) { hkind = 2; sub = xl / ZM_TY; ty = xl % ZM_TY; tx = ZM_TX; }
            else { hkind = 0; sub = 0; tx = 0; ty = 0; }
        }
    }
    const bool do_x = is_main || hkind == 2, do_y = is_main || hkind == 1;
    // per-thread shared addresses (bytes)
    const unsigned aC = sbase + 8u * (SM::OFF_C + sub * SM::C_SUB + (ty + 3) * ZM_CW + tx + 3);
    const unsigned aFx = sbase + 8u * (SM::OFF_FX + sub * SM::FX_SUB + ty * (ZM_TX + 1) + tx);
    const unsigned aFy = sbase + 8u * (SM::OFF_FY + sub * SM::FY_SUB + ty * ZM_TX + tx);
    const unsigned aQx = sbase + 8u * (SM::OFF_QX + ty * (ZM_TX + 1) + tx);
    const unsigned aQy = sbase + 8u * (SM::OFF_QY + ty * ZM_TX + tx);
    constexpr unsigned C_BUF_B = 8u * SM::C_BUF, FX_BUF_B = 8u * NSUB * SM::FX_SUB, FY_BUF_B = 8u * NSUB * SM::FY_SUB;
    constexpr unsigned QX_BUF_B = 8u * SM::FX_SUB, QY_BUF_B = 8u * SM::FY_SUB;
    const int partner_off = (NSUB == 2) ? (sub ? -8 * SM::C_SUB : 8 * SM::C_SUB) : 0;

    // element offsets of this thread's column: `off` into layout-shaped arrays, `toff` into the tracer stacks
    int64_t off = L.at(i0 + tx, j0 + ty, 0);
    int64_t toff = (int64_t)(item.tracer_c + sub) * L.len + off;
    const double *__restrict__ Uold = P.U_old;
    const double *__restrict__ velu = P.vel[0], *__restrict__ velv = P.vel[1], *__restrict__ velw = P.vel[2];
    const double *__restrict__ varp = item.is_map ? nullptr : P.var[item.src];
    const bool need_var = is_main && !item.is_map && sub == 0;
    const bool use_G = !P.first_stage;

    // halo duties of the tile (main threads only): up to two elements per thread
    unsigned h_s0 = 0, h_s1 = 0;                 // shared byte address in buffer 0 (0 = none)
    int64_t h_g0 = 0, h_g1 = 0;                  // element offset in the tracer stack, plane 0
    if (is_main) {
        const int t = tid % ZM_MAIN;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = t + r * ZM_MAIN;
            int row = -1, cc = 0;
            if (e < 3 * ZM_CW) { row = e / ZM_CW; cc = e % ZM_CW; }
            else if (e < 6 * ZM_CW) { row = ZM_TY + 3 + (e - 3 * ZM_CW) / ZM_CW; cc = (e - 3 * ZM_CW) % ZM_CW; }
            else if (e < 6 * ZM_CW + 3 * ZM_TY) { row = 3 + (e - 6 * ZM_CW) / 3; cc = (e - 6 * ZM_CW) % 3; }
            else if (e < 6 * ZM_CW + 6 * ZM_TY) { row = 3 + (e - 6 * ZM_CW - 3 * ZM_TY) / 3; cc = ZM_TX + 3 + (e - 6 * ZM_CW - 3 * ZM_TY) % 3; }
            if (row >= 0) {
                const unsigned sa = sbase + 8u * (SM::OFF_C + sub * SM::C_SUB + row * ZM_CW + cc);
                const int64_t ga = (int64_t)(item.tracer_c + sub) * L.len + L.at(i0 - 3 + cc, j0 - 3 + row, 0);
                if (r == 0) { h_s0 = sa; h_g0 = ga; } else { h_s1 = sa; h_g1 = ga; }
            }
        }
    }

    // ---- prologue: z window of the column, plane-0 tile, inputs of plane 0 --------------------------
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;     // d[k-2..k+2], d[m] = c[m+1] - c[m]
    double Qa = 0, Qb = 0, Qc = 0, Qd = 0;             // Q[k-1..k+2]
    double ck = 0, ck1 = 0, ck2 = 0, ck3 = 0, ck4 = 0; // c[k..k+4]
    double Fb = 0, wb = 0;
    double cu = 0, cv = 0, cw = 0, cG = 0, cf = 0;     // inputs of the current plane
    if (is_main) {
        const double cm2 = Uold[toff - 2 * plane], cm1 = Uold[toff - plane];
        ck = Uold[toff]; ck1 = Uold[toff + plane]; ck2 = Uold[toff + 2 * plane];
        ck3 = Uold[toff + 3 * plane]; ck4 = Uold[toff + 4 * plane];
        d0 = cm1 - cm2; d1 = ck - cm1; d2 = ck1 - ck; d3 = ck2 - ck1; d4 = ck3 - ck2;
        double s;
        s = d1 - d0; Qa = 3.25 * s * s;
        s = d2 - d1; Qb = 3.25 * s * s;
        s = d3 - d2; Qc = 3.25 * s * s;
        s = d4 - d3; Qd = 3.25 * s * s;
        wb = velw[off];
        Fb = (P.area[2] * wb) * ck;                     // wall face m = 1: Centered2 on the mirrored halo
        sts(aC, ck);
        if (h_s0) sts(h_s0, Uold[h_g0]);
        if (h_s1) sts(h_s1, Uold[h_g1]);
        cw = velw[off + plane];
        if (use_G) cG = P.G[toff];
        if (need_var) cf = varp[off];
    }
    if (do_x) cu = velu[off];
    if (do_y) cv = velv[off];
    __syncthreads();

    for (int k = 0; k < Nz; ++k) {
        const unsigned b = k & 1;
        const bool more = k + 1 < Nz;
        // ---- next plane: tile interior + halo, raw inputs (consumed after the barrier / next iteration) ----
        double nu = 0, nv = 0, nw = 0, nG = 0, nf = 0;
        if (more) {
            if (is_main) {
                sts(aC + (b ^ 1) * C_BUF_B, ck1);
                if (h_s0) cp_async8(h_s0 + (b ^ 1) * C_BUF_B, Uold + h_g0 + plane);
                if (h_s1) cp_async8(h_s1 + (b ^ 1) * C_BUF_B, Uold + h_g1 + plane);
                nw = velw[off + 2 * plane];
                if (use_G) nG = P.G[toff + plane];
                if (need_var) nf = varp[off + plane];
            }
            if (do_x) nu = velu[off + plane];
            if (do_y) nv = velv[off + plane];
        }
        // ---- faces of plane k ------------------------------------------------------------------
        const unsigned aCb = aC + b * C_BUF_B;
        double Fw = 0, Fs = 0, Ft = 0, other = 0;
        if (is_main && NSUB == 2) other = lds(aCb + partner_off);
        const double cmid = is_main ? ck : lds(aCb);
        if (do_x) {                                      // west face of (i, j)  [helpers: east edge of the tile]
            Fw = (P.area[0] * cu) * face_from_tile(aCb, 8, cmid, cu);
            sts(aFx + b * FX_BUF_B, Fw);
            if (sub == 0) sts(aQx + b * QX_BUF_B, cu);
        }
        if (do_y) {                                      // south face of (i, j) [helpers: north edge of the tile]
            Fs = (P.area[1] * cv) * face_from_tile(aCb, 8 * ZM_CW, cmid, cv);
            sts(aFy + b * FY_BUF_B, Fs);
            if (sub == 0) sts(aQy + b * QY_BUF_B, cv);
        }
        if (is_main) {                                   // top face of cell k: 1-based face index m = k + 2
            const int m = k + 2;
            const bool left = cw > 0.0;
            double ct;
            if (m >= 4 && m <= Nz - 2) {
                // the increment is odd in the differences: right-biased = minus the mirrored evaluation
                const double inc = weno5z_inc_q(left ? d0 : d4, left ? d1 : d3, d2, left ? d3 : d1,
                                                left ? Qa : Qd, left ? Qb : Qc, left ? Qc : Qb);
                ct = left ? ck + inc : ck1 - inc;
            } else if (m == Nz + 1) {
                ct = ck;                                  // top wall
            } else if (m == 3 || m == Nz - 1) {
                const double inc = weno3z_inc(left ? d1 : d3, d2);
                ct = left ? ck + inc : ck1 - inc;
            } else {
                ct = (ck + ck1) / 2.0;                    // m = 2, Nz: Centered2
            }
            Ft = (P.area[2] * cw) * ct;
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- tendency, RK3 substep, advance the window ---------------------------------------------
        if (is_main) {
            const double Fe = lds(aFx + b * FX_BUF_B + 8), Fn = lds(aFy + b * FY_BUF_B + 8 * ZM_TX);
            const double ue = lds(aQx + b * QX_BUF_B + 8), vn = lds(aQy + b * QY_BUF_B + 8 * ZM_TX);
            const double inv_vol = 1.0 / P.volume;
            double div_flux = (Fe - Fw);
            div_flux = div_flux + (Fn - Fs);
            div_flux = div_flux + (Ft - Fb);
            div_flux *= inv_vol;
            double div_u = (P.area[0] * ue - P.area[0] * cu);
            div_u = div_u + (P.area[1] * vn - P.area[1] * cv);
            div_u = div_u + (P.area[2] * cw - P.area[2] * wb);
            div_u *= inv_vol;
            const double c = item.c, d = item.d;
            double gC, gS;
            if (NSUB == 2) { gC = sub ? other : ck; gS = sub ? ck : other; }
            else { gC = ck; gS = 0.0; }
            double src = cf;
            if (item.is_map) src = item.src == 0 ? (cu + ue) / 2.0 : (item.src == 1 ? (cv + vn) / 2.0 : (wb + cw) / 2.0);
            const double n2 = c * c + d * d;
            double forcing;
            if (NSUB == 1)      forcing = item.is_map ? (-c / n2 * src) + (-c * gC) : (-c * gC) + src;
            else if (sub == 0)  forcing = item.is_map ? (-c / n2 * src) + (-c * gC - d * gS) : (-c * gC - d * gS) + src;
            else                forcing = item.is_map ? (-d / n2 * src) + (-c * gS + d * gC) : (-c * gS + d * gC);
            const double Gn = (-div_flux + ck * div_u) + forcing;
            double unew;
            if (P.first_stage) unew = ck + P.dt * P.gamma * Gn;
            else               unew = ck + P.dt * (P.gamma * Gn + P.zeta * cG);
            P.U_new[toff] = unew;
            if (P.store_G) P.G[toff] = Gn;
            // advance the window to cell k+1
            Fb = Ft; wb = cw;
            double c5 = 0;
            if (k + 5 <= Nz + 2) c5 = Uold[toff + 5 * plane];
            const double dnew = ck4 - ck3;                // d[k+3]
            const double s = dnew - d4;
            d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = dnew;
            Qa = Qb; Qb = Qc; Qc = Qd; Qd = 3.25 * s * s;
            ck = ck1; ck1 = ck2; ck2 = ck3; ck3 = ck4; ck4 = c5;
        }
        cu = nu; cv = nv; cw = nw; cG = nG; cf = nf;
        off += plane; toff += plane;
        h_g0 += plane; h_g1 += plane;
    }
}

extern "C" int lfb_zmarch_supported(const LfbStageParams *P)
{
    const LfbLayout &L = P->L;
    if (!P->advect || P->relax) return 0;
    if (L.topo[0] != LFB_PERIODIC || L.topo[2] != LFB_BOUNDED) return 0;
    if (L.topo[1] != LFB_PERIODIC && !(L.connected_lo[1] && L.connected_hi[1])) return 0;
    if (!(P->vel_present[0] && P->vel_present[1] && P->vel_present[2])) return 0;
    if (L.N[0] % ZM_TX || (P->j1 - P->j0) % ZM_TY || L.N[2] < 8) return 0;
    if (L.H[0] < 3 || L.H[1] < 3 || L.H[2] < 3) return 0;
    return 1;
}

extern "C" cudaError_t lfb_launch_stage_zmarch(const LfbStageParams *P, cudaStream_t stream)
{
    const LfbLayout &L = P->L;
    if (P->j1 <= P->j0) return cudaSuccess;
    const int ytiles = (P->j1 - P->j0) / ZM_TY;
    dim3 grid(P->n_items * ytiles, L.N[0] / ZM_TX, 1);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(lfb_stage_zmarch<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * ZmSmem<1>::TOTAL));
        cudaFuncSetAttribute(lfb_stage_zmarch<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * ZmSmem<2>::TOTAL));
        configured = true;
    }
    if (P->single_exp) lfb_stage_zmarch<1><<<grid, 1 * ZM_MAIN + 32 + 32, 8 * ZmSmem<1>::TOTAL, stream>>>(*P);
    else               lfb_stage_zmarch<2><<<grid, 2 * ZM_MAIN + 64 + 32, 8 * ZmSmem<2>::TOTAL, stream>>>(*P);
    return cudaGetLastError();
}
