// lfb_zmarch.cu -- the production RK3 stage kernel for 3-D grids that are Periodic (or slab-connected)
// in x and y and Bounded in z: shared-memory x-y tiles marching along z.
//
// Same fused work as lfb_stage_generic (time interpolation of the inputs, WENO5-Z flux divergence,
// c div(U), forcing, RK3 substep, tendency caching; reference call sites listed there), reorganised
// around the FP64 pipe, which bounds this kernel on B200 (DESIGN.md section "roofline"):
//   * every face value is computed ONCE: a thread owns one cell column of one tracer and computes
//     its west, south and top face; the east / north fluxes come from the neighbouring threads
//     through shared memory, the bottom flux is carried in a register along the march;
//   * the extra faces on the east / north edge of a tile are packed into three helper warps
//     instead of diverging lanes of the main warps;
//   * along z the first differences d and the curvature terms Q = 13/4 (d_m - d_{m-1})^2 of the
//     smoothness indicators live in a register window and are computed once per cell (both upwind
//     orientations and both neighbouring faces reuse them);
//   * one division per face (common-denominator form of the Z weights, lfb_math.cuh).
// One block handles the two tracers (cosine, sine) of one (field, coefficient-set) item on a
// 32 x 8 tile; x-y neighbours are read from a double-buffered tile with a 3-wide halo that is filled
// with cp.async one plane ahead; global loads for plane k+1 are issued before the arithmetic of
// plane k retires.  Blocks of the same tile (different items) are adjacent in launch order so the
// velocity planes they share are served by L2.
#include <cuda_runtime.h>

#include "lfb_internal.h"
#include "lfb_math.cuh"

#define ZM_TX 32
#define ZM_TY 8
#define ZM_MAIN (ZM_TX * ZM_TY)
#define ZM_CW (ZM_TX + 6)
#define ZM_CH (ZM_TY + 6)

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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

// upwind WENO5-Z face value from a line of values in shared memory: `c` points at the cell on the
// high side of the face (cell m), `s` is the stride along the axis, cm = c[0] if already in a register
__device__ __forceinline__ double face_from_line(const double *c, int s, double cm, double q)
{
    const bool left = q > 0.0;
    const double a = c[-s], b = c[-2 * s], e1 = c[s];
    const double far = c[left ? -3 * s : 2 * s];
    const double n = left ? a : cm;
    const double e = left ? cm : a;
    const double o = left ? b : e1;
    const double f = left ? e1 : b;
    return n + weno5z_increment(o - far, n - o, e - n, f - e);
}

struct ZmShared {
    double C[2][2][ZM_CH][ZM_CW];          // [buffer][sub][row][col] tracer plane with halo
    double Fx[2][ZM_TY][ZM_TX + 1];        // x-face fluxes  [sub][row][face]
    double Fy[2][ZM_TY + 1][ZM_TX];        // y-face fluxes  [sub][face row][col]
    double Qx[ZM_TY][ZM_TX + 1];           // interpolated u at the x faces
    double Qy[ZM_TY + 1][ZM_TX];           // interpolated v at the y faces
};

template <int NSUB>
__global__ void __launch_bounds__(NSUB *ZM_MAIN + NSUB * 32 + 32, 1) lfb_stage_zmarch(const __grid_constant__ LfbStageParams P)
{
    __shared__ ZmShared S;
    const LfbLayout &L = P.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % P.n_items;
    const int ytile = blockIdx.x / P.n_items;
    const int i0 = blockIdx.y * ZM_TX;
    const int j0 = P.j0 + ytile * ZM_TY;
    const LfbItem item = P.items[item_id];
    const int Nz = L.N[2];
    const int64_t plane = L.plane;
    const double w1 = P.w1, w2 = P.w2;
    const bool interp = P.interp != 0;
    const double vs = P.vel_sign;
    const double Ax = P.area[0], Ay = P.area[1], Az = P.area[2];
    const double inv_vol = 1.0 / P.volume;

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
    const int i = i0 + tx, j = j0 + ty;
    const int64_t col = L.at(i, j, 0);           // plane-0 index of this thread's column
    const double *U = P.U_old + (int64_t)(item.tracer_c + sub) * L.len;

    // halo duties of the tile (main threads only): up to two elements per thread
    int h_s[2] = {-1, -1};
    int64_t h_g[2] = {0, 0};
    if (is_main) {
        const int t = tid % ZM_MAIN;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = t + r * ZM_MAIN;
            int row, cc;
            if (e < 3 * ZM_CW) { row = e / ZM_CW; cc = e % ZM_CW; }
            else if (e < 6 * ZM_CW) { row = ZM_TY + 3 + (e - 3 * ZM_CW) / ZM_CW; cc = (e - 3 * ZM_CW) % ZM_CW; }
            else if (e < 6 * ZM_CW + 3 * ZM_TY) { row = 3 + (e - 6 * ZM_CW) / 3; cc = (e - 6 * ZM_CW) % 3; }
            else if (e < 6 * ZM_CW + 6 * ZM_TY) { row = 3 + (e - 6 * ZM_CW - 3 * ZM_TY) / 3; cc = ZM_TX + 3 + (e - 6 * ZM_CW - 3 * ZM_TY) % 3; }
            else { row = -1; cc = 0; }
            if (row >= 0) {
                h_s[r] = row * ZM_CW + cc;
                h_g[r] = L.at(i0 - 3 + cc, j0 - 3 + row, 0);
            }
        }
    }

    // ---- prologue: z window of the column -------------------------------------------------------
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;     // d[k-2..k+2], d[m] = c[m+1] - c[m]
    double Qa = 0, Qb = 0, Qc = 0, Qd = 0;             // Q[k-1..k+2]
    double ck = 0, ck1 = 0, ck2 = 0, clast = 0, cnext = 0;      // c[k], c[k+1], c[k+2], c[k+3], c[k+4]
    double Fb = 0, wb = 0;
    if (is_main) {
        const double cm2 = U[col - 2 * plane], cm1 = U[col - plane];
        ck = U[col]; ck1 = U[col + plane];
        ck2 = U[col + 2 * plane];
        clast = U[col + 3 * plane];
        cnext = U[col + 4 * plane];
        d0 = cm1 - cm2; d1 = ck - cm1; d2 = ck1 - ck; d3 = ck2 - ck1; d4 = clast - ck2;
        double s;
        s = d1 - d0; Qa = 3.25 * s * s;
        s = d2 - d1; Qb = 3.25 * s * s;
        s = d3 - d2; Qc = 3.25 * s * s;
        s = d4 - d3; Qd = 3.25 * s * s;
        double w = P.vel1[2][col];
        if (interp) w = P.vel2[2][col] * w2 + w * w1;
        wb = vs < 0 ? -w : w;
        Fb = (Az * wb) * ck;                            // wall face m = 1: Centered2 on the mirrored halo
        S.C[0][sub][ty + 3][tx + 3] = ck;
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (h_s[r] >= 0) (&S.C[0][sub][0][0])[h_s[r]] = U[h_g[r]];
    }
    // raw inputs of plane 0
    double ru1 = 0, ru2 = 0, rv1 = 0, rv2 = 0, rw1 = 0, rw2 = 0, rG = 0, rf1 = 0, rf2 = 0;
    const int64_t gix = (int64_t)(item.tracer_c + sub) * L.len + col;
    const bool need_var = is_main && !item.is_map && sub == 0;
    {
        const int64_t q = col;
        if (is_main || hkind == 2) { ru1 = P.vel1[0][q]; if (interp) ru2 = P.vel2[0][q]; }
        if (is_main || hkind == 1) { rv1 = P.vel1[1][q]; if (interp) rv2 = P.vel2[1][q]; }
        if (is_main) {
            rw1 = P.vel1[2][q + plane]; if (interp) rw2 = P.vel2[2][q + plane];
            if (!P.first_stage) rG = P.G[gix];
            if (need_var) { rf1 = P.var1[item.src][q]; if (interp) rf2 = P.var2[item.src][q]; }
        }
    }
    __syncthreads();

    for (int k = 0; k < Nz; ++k) {
        const int b = k & 1, nb = b ^ 1;
        const int64_t kp = (int64_t)k * plane;
        // ---- phase B: face values of plane k ---------------------------------------------------
        // halo of plane k+1 -> other tile buffer, asynchronously (no registers)
        if (is_main && k + 1 < Nz) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (h_s[r] >= 0) cp_async8(&(&S.C[nb][sub][0][0])[h_s[r]], U + h_g[r] + kp + plane);
        }
        double uw = interp ? ru2 * w2 + ru1 * w1 : ru1;
        double vsouth = interp ? rv2 * w2 + rv1 * w1 : rv1;
        double wt = interp ? rw2 * w2 + rw1 * w1 : rw1;
        if (vs < 0) { uw = -uw; vsouth = -vsouth; wt = -wt; }
        double Fw = 0, Fs = 0, Ft = 0;
        const double *cc = &S.C[b][sub][ty + 3][tx + 3];
        if (is_main || hkind == 2) {                     // west face of (i, j)  [helpers: east edge of the tile]
            const double cf = face_from_line(cc, 1, is_main ? ck : cc[0], uw);
            Fw = (Ax * uw) * cf;
            S.Fx[sub][ty][tx] = Fw;
            if (sub == 0) S.Qx[ty][tx] = uw;
        }
        if (is_main || hkind == 1) {                     // south face of (i, j) [helpers: north edge of the tile]
            const double cf = face_from_line(cc, ZM_CW, is_main ? ck : cc[0], vsouth);
            Fs = (Ay * vsouth) * cf;
            S.Fy[sub][ty][tx] = Fs;
            if (sub == 0) S.Qy[ty][tx] = vsouth;
        }
        if (is_main) {                                   // top face of cell k: 1-based face index m = k + 2
            const int m = k + 2;
            double ct;
            if (m >= 4 && m <= Nz - 2) {
                const bool left = wt > 0.0;
                const double dp = left ? d0 : -d4, do_ = left ? d1 : -d3, dn = left ? d2 : -d2, de = left ? d3 : -d1;
                const double Q2 = left ? Qa : Qd, Q1 = left ? Qb : Qc, Q0 = left ? Qc : Qb;
                ct = (left ? ck : ck1) + weno5z_inc_q(dp, do_, dn, de, Q2, Q1, Q0);
            } else if (m == Nz + 1) {
                ct = ck;                                  // top wall
            } else if (m == 3 || m == Nz - 1) {
                const bool left = wt > 0.0;
                ct = left ? ck + weno3z_inc(d1, d2) : ck1 + weno3z_inc(-d3, -d2);
            } else {
                ct = (ck + ck1) / 2.0;                    // m = 2, Nz: Centered2
            }
            Ft = (Az * wt) * ct;
        }
        __syncthreads();

        // ---- phase C: tendency, RK3 substep, advance the window --------------------------------
        const double G_prev = rG;
        const double f_src = need_var ? (interp ? rf2 * w2 + rf1 * w1 : rf1) : 0.0;
        if (k + 1 < Nz) {                                 // raw inputs of plane k+1 (consumed after the next barrier)
            const int64_t q = col + kp + plane;
            if (is_main || hkind == 2) { ru1 = P.vel1[0][q]; if (interp) ru2 = P.vel2[0][q]; }
            if (is_main || hkind == 1) { rv1 = P.vel1[1][q]; if (interp) rv2 = P.vel2[1][q]; }
            if (is_main) {
                rw1 = P.vel1[2][q + plane]; if (interp) rw2 = P.vel2[2][q + plane];
                if (!P.first_stage) rG = P.G[gix + kp + plane];
                if (need_var) { rf1 = P.var1[item.src][q]; if (interp) rf2 = P.var2[item.src][q]; }
            }
        }
        if (is_main) {
            const double Fe = S.Fx[sub][ty][tx + 1], Fn = S.Fy[sub][ty + 1][tx];
            const double ue = S.Qx[ty][tx + 1], vn = S.Qy[ty + 1][tx];
            double div_flux = (Fe - Fw);
            div_flux = div_flux + (Fn - Fs);
            div_flux = div_flux + (Ft - Fb);
            div_flux *= inv_vol;
            double div_u = (Ax * ue - Ax * uw);
            div_u = div_u + (Ay * vn - Ay * vsouth);
            div_u = div_u + (Az * wt - Az * wb);
            div_u *= inv_vol;
            const double c = item.c, d = item.d;
            double gC, gS;
            if (NSUB == 2) {
                const double other = S.C[b][sub ^ 1][ty + 3][tx + 3];
                gC = sub ? other : ck; gS = sub ? ck : other;
            } else { gC = ck; gS = 0.0; }
            double src = f_src;
            if (item.is_map) src = item.src == 0 ? (uw + ue) / 2.0 : (item.src == 1 ? (vsouth + vn) / 2.0 : (wb + wt) / 2.0);
            const double n2 = c * c + d * d;
            double forcing;
            if (NSUB == 1)      forcing = item.is_map ? (-c / n2 * src) + (-c * gC) : (-c * gC) + src;
            else if (sub == 0)  forcing = item.is_map ? (-c / n2 * src) + (-c * gC - d * gS) : (-c * gC - d * gS) + src;
            else                forcing = item.is_map ? (-d / n2 * src) + (-c * gS + d * gC) : (-c * gS + d * gC);
            const double Gn = (-div_flux + ck * div_u) + forcing;
            const int64_t o = gix + kp;
            double unew;
            if (P.first_stage) unew = ck + P.dt * P.gamma * Gn;
            else               unew = ck + P.dt * (P.gamma * Gn + P.zeta * G_prev);
            P.U_new[o] = unew;
            if (P.store_G) P.G[o] = Gn;
            // advance the window to cell k+1
            Fb = Ft; wb = wt;
            const double dnew = cnext - clast;            // d[k+3]
            const double s = dnew - d4;
            d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = dnew;
            Qa = Qb; Qb = Qc; Qc = Qd; Qd = 3.25 * s * s;
            ck = ck1; ck1 = ck2; ck2 = clast; clast = cnext;
            if (k + 5 <= Nz + 2) cnext = U[col + kp + 5 * plane];
            if (k + 1 < Nz) S.C[nb][sub][ty + 3][tx + 3] = ck;
        }
        cp_async_wait_all();
        __syncthreads();
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
    if (P->single_exp) lfb_stage_zmarch<1><<<grid, 1 * ZM_MAIN + 32 + 32, 0, stream>>>(*P);
    else               lfb_stage_zmarch<2><<<grid, 2 * ZM_MAIN + 64 + 32, 0, stream>>>(*P);
    return cudaGetLastError();
}
