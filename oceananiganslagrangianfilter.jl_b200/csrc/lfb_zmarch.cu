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

#include "lfb_zm_common.cuh"

// shared memory carve-up (doubles).  Every per-cell array (tracer tile, x/y fluxes, x/y face velocities)
// uses the geometry of the haloed tile (ZM_CH rows of ZM_CW), so one per-thread offset addresses them
// all; everything that is double-buffered along the march lives in one per-buffer block, so a single
// uniform offset (k & 1) * BUF selects the buffer.
template <int NSUB, int TY>
struct ZmSmem {
    static constexpr int CH = TY + 6;
    static constexpr int TILE = CH * ZM_CW;                     // one haloed plane
    static constexpr int NMAIN = NSUB * ZM_TX * TY;
    static constexpr int OFF_C = 0;                             // C [NSUB][TILE]  tracer values
    static constexpr int OFF_FX = OFF_C + NSUB * TILE;          // Fx[NSUB][TILE]  west-face fluxes
    static constexpr int OFF_FY = OFF_FX + NSUB * TILE;         // Fy[NSUB][TILE]  south-face fluxes
    static constexpr int OFF_QX = OFF_FY + NSUB * TILE;         // Qx[TILE]        u at west faces
    static constexpr int OFF_QY = OFF_QX + TILE;                // Qy[TILE]        v at south faces
    static constexpr int OFF_IN = OFF_QY + TILE;                // In[6][NMAIN]    per-thread inputs u,v,w,G-,f,c[k+5]
    static constexpr int BUF = (OFF_IN + 6 * NMAIN + 15) / 16 * 16;
    static constexpr int BYTES = 2 * BUF * 8;
    static constexpr int THREADS = NMAIN + NSUB * 32 + 32;
};

template <int NSUB, int TY>
__global__ void __launch_bounds__(ZmSmem<NSUB, TY>::THREADS, 1) lfb_stage_zmarch(const __grid_constant__ LfbStageParams P)
{
    using SM = ZmSmem<NSUB, TY>;
    extern __shared__ __align__(16) double zm_smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(zm_smem);
    constexpr unsigned BUF_B = 8u * SM::BUF;
    constexpr int ROW_B = 8 * ZM_CW;                                  // byte stride between tile rows
    constexpr unsigned oC = 8u * SM::OFF_C, oFX = 8u * SM::OFF_FX, oFY = 8u * SM::OFF_FY, oQX = 8u * SM::OFF_QX,
                       oQY = 8u * SM::OFF_QY;
    const LfbLayout &L = P.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % P.n_items;
    const int ytile = blockIdx.x / P.n_items;
    const int i0 = blockIdx.y * ZM_TX;
    const int j0 = P.j0 + ytile * TY;
    const LfbItem item = P.items[item_id];
    const int Nz = L.N[2];
    const int64_t plane = L.plane;
    const int jmax = L.N[1] + L.H[1];                                 // last row that exists in the arrays

    if (tid >= SM::NMAIN) {
        // ================= helper warps: the extra faces on the north / east edge of the tile =================
        const int hl = tid - SM::NMAIN;
        int sub, tx, ty, kind;                   // kind 1: y face at row j0+TY, 2: x face at column i0+TX, 0: idle lane
        if (hl < NSUB * 32) { kind = 1; sub = hl / 32; tx = hl % 32; ty = TY; }
        else {
            const int xl = hl - NSUB * 32;
            if (xl < NSUB * TY) { kind = 2; sub = xl / TY; ty = xl % TY; tx = ZM_TX; }
            else { kind = 0; sub = 0; tx = 0; ty = 0; }
        }
        const unsigned aT = sbase + 8u * (sub * SM::TILE + (ty + 3) * ZM_CW + tx + 3);
        const unsigned aF = aT + (kind == 1 ? oFY : oFX);
        const unsigned aQ = sbase + 8u * ((ty + 3) * ZM_CW + tx + 3) + (kind == 1 ? oQY : oQX);
        const double *__restrict__ vel = (kind == 1 ? P.vel[1] : P.vel[0]) + L.at(i0 + tx, min(j0 + ty, jmax), 0);
        const double area = kind == 1 ? P.area[1] : P.area[0];
        double q = kind ? vel[0] : 0.0;
        block_barrier();                                            // prologue barrier of the main warps
        for (int k = 0; k < Nz; ++k) {
            const unsigned boff = (k & 1) * BUF_B;
            if (kind) {
                const double cm = lds(aT + oC + boff);
                const double cf = kind == 1 ? zm_face_from_tile<ROW_B>(aT + oC + boff, cm, q)
                                            : zm_face_from_tile<8>(aT + oC + boff, cm, q);
                sts(aF + boff, (area * q) * cf);
                if (sub == 0) sts(aQ + boff, q);
                vel += plane;
                if (k + 1 < Nz) q = vel[0];
            }
            block_barrier();
        }
        return;
    }

    // ================= main warps: one cell column of one tracer per thread =================
    const int tx = tid % ZM_TX, ty = (tid / ZM_TX) % TY, sub = tid / (ZM_TX * TY);
    const unsigned aT = sbase + 8u * (sub * SM::TILE + (ty + 3) * ZM_CW + tx + 3);       // tracer-indexed arrays
    const unsigned aV = sbase + 8u * ((ty + 3) * ZM_CW + tx + 3);                        // velocity arrays
    const unsigned aIn = sbase + 8u * (SM::OFF_IN + tid);
    constexpr unsigned IN_F = 8u * SM::NMAIN;                        // stride between the input fields of a thread
    const int partner_off = (NSUB == 2) ? (sub ? -8 * SM::TILE : 8 * SM::TILE) : 0;
    const bool active = j0 + ty < P.j1;                              // partial tiles at the end of the y range

    // global pointers of this thread's column (advanced by one plane per iteration)
    const int64_t off0 = L.at(i0 + tx, min(j0 + ty, jmax), 0);
    const int64_t tbase = (int64_t)(item.tracer_c + sub) * L.len;
    const double *__restrict__ pU = P.U_old + tbase + off0;
    double *__restrict__ pUn = P.U_new + tbase + off0;
    double *__restrict__ pG = P.G + tbase + off0;
    const double *__restrict__ pu = P.vel[0] + off0, *__restrict__ pv = P.vel[1] + off0, *__restrict__ pw = P.vel[2] + off0;
    const bool need_var = !item.is_map && sub == 0;
    const double *__restrict__ pf = need_var ? P.var[item.src] + off0 : nullptr;
    const bool use_G = !P.first_stage;
    // offsets of this column's periodic images in the halo (0 = none); y only when the y halo is not owned by a
    // neighbour rank
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
    const double inv_vol = 1.0 / P.volume;
    const double dtg = P.dt * P.gamma, dtz = P.dt * P.zeta;

    // halo duties of the tile: up to two elements per thread
    unsigned h_s0 = 0, h_s1 = 0;                 // shared byte address in buffer 0 (0 = none)
    const double *h_g0 = nullptr, *h_g1 = nullptr;
    {
        constexpr int NT = ZM_TX * TY;
        const int t = tid % NT;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = t + r * NT;
            int row = -1, cc = 0;
            if (e < 3 * ZM_CW) { row = e / ZM_CW; cc = e % ZM_CW; }
            else if (e < 6 * ZM_CW) { row = TY + 3 + (e - 3 * ZM_CW) / ZM_CW; cc = (e - 3 * ZM_CW) % ZM_CW; }
            else if (e < 6 * ZM_CW + 3 * TY) { row = 3 + (e - 6 * ZM_CW) / 3; cc = (e - 6 * ZM_CW) % 3; }
            else if (e < 6 * ZM_CW + 6 * TY) { row = 3 + (e - 6 * ZM_CW - 3 * TY) / 3; cc = ZM_TX + 3 + (e - 6 * ZM_CW - 3 * TY) % 3; }
            if (row >= 0) {
                const unsigned sa = sbase + oC + 8u * (sub * SM::TILE + row * ZM_CW + cc);
                const double *ga = P.U_old + tbase + L.at(i0 - 3 + cc, min(j0 - 3 + row, jmax), 0);
                if (r == 0) { h_s0 = sa; h_g0 = ga; } else { h_s1 = sa; h_g1 = ga; }
            }
        }
    }
    static_assert(6 * ZM_CW + 6 * TY <= 2 * ZM_TX * TY, "tile halo does not fit two duties per thread");

    // ---- prologue: z window of the column, plane-0 tile and inputs ------------------------------------
    double d0, d1, d2, d3, d4;                   // d[k-2..k+2], d[m] = c[m+1] - c[m]
    double Qa, Qb, Qc, Qd;                       // Q[k-1..k+2], Q[m] = 13/4 (d[m] - d[m-1])^2
    double ck, ck1, ck2, ck3, ck4;               // c[k..k+4]
    double Fb, wb;
    {
        const double cm2 = pU[-2 * plane], cm1 = pU[-plane];
        ck = pU[0]; ck1 = pU[plane]; ck2 = pU[2 * plane]; ck3 = pU[3 * plane]; ck4 = pU[4 * plane];
        d0 = cm1 - cm2; d1 = ck - cm1; d2 = ck1 - ck; d3 = ck2 - ck1; d4 = ck3 - ck2;
        double s;
        s = d1 - d0; Qa = 3.25 * s * s;
        s = d2 - d1; Qb = 3.25 * s * s;
        s = d3 - d2; Qc = 3.25 * s * s;
        s = d4 - d3; Qd = 3.25 * s * s;
        wb = pw[0];
        Fb = (P.area[2] * wb) * ck;                     // wall face m = 1: Centered2 on the mirrored halo
        sts(aT + oC, ck);
        if (h_s0) cp_async8(h_s0, h_g0);
        if (h_s1) cp_async8(h_s1, h_g1);
        cp_async8(aIn, pu);
        cp_async8(aIn + IN_F, pv);
        cp_async8(aIn + 2 * IN_F, pw + plane);
        if (use_G) cp_async8(aIn + 3 * IN_F, pG);
        if (need_var) cp_async8(aIn + 4 * IN_F, pf);
        cp_async8(aIn + 5 * IN_F, pU + 5 * plane);
        cp_async_wait_all();
    }
    block_barrier();

#ifdef ZM_UNROLL
#pragma unroll ZM_UNROLL
#endif
    for (int k = 0; k < Nz; ++k) {
        const unsigned boff = (k & 1) * BUF_B, noff = BUF_B - boff;
        // ---- plane k+1 -> the other buffer: own value, halo and per-thread inputs, all asynchronous ----
        if (k + 1 < Nz) {
            sts(aT + oC + noff, ck1);
            if (h_s0) cp_async8(h_s0 + noff, h_g0 + plane);
            if (h_s1) cp_async8(h_s1 + noff, h_g1 + plane);
            cp_async8(aIn + noff, pu + plane);
            cp_async8(aIn + noff + IN_F, pv + plane);
            cp_async8(aIn + noff + 2 * IN_F, pw + 2 * plane);
            if (use_G) cp_async8(aIn + noff + 3 * IN_F, pG + plane);
            if (need_var) cp_async8(aIn + noff + 4 * IN_F, pf + plane);
            if (k + 6 <= Nz + 2) cp_async8(aIn + noff + 5 * IN_F, pU + 6 * plane);
        }
        // ---- faces of plane k ------------------------------------------------------------------
        const unsigned aCb = aT + oC + boff;
        const double cu = lds(aIn + boff), cv = lds(aIn + boff + IN_F), cw = lds(aIn + boff + 2 * IN_F);
        double other = 0.0;
        if (NSUB == 2) other = lds(aCb + partner_off);
        const bool left_x = cu > 0.0, left_y = cv > 0.0;
        const ZmLine lx = zm_load_line<8>(aCb, left_x);
        const ZmLine ly = zm_load_line<ROW_B>(aCb, left_y);
        const double Fw = (P.area[0] * cu) * zm_face(lx, ck, left_x);
        const double Fs = (P.area[1] * cv) * zm_face(ly, ck, left_y);
        sts(aT + oFX + boff, Fw);
        sts(aT + oFY + boff, Fs);
        if (sub == 0) { sts(aV + oQX + boff, cu); sts(aV + oQY + boff, cv); }
        double ct;
        {                                                // top face of cell k: 1-based face index m = k + 2
            const int m = k + 2;
            const bool left = cw > 0.0;
            if (m >= 4 && m <= Nz - 2) {
                const double inc = zm_inc_q(left ? d0 : d4, left ? d1 : d3, d2, left ? d3 : d1,
                                            left ? Qa : Qd, left ? Qb : Qc, left ? Qc : Qb);
                ct = left ? ck + inc : ck1 - inc;
            } else if (m == Nz + 1) {
                ct = ck;                                  // top wall
            } else if (m == 3 || m == Nz - 1) {
                const double inc = weno3z_inc(left ? d1 : d3, d2);
                ct = left ? ck + inc : ck1 - inc;
            } else {
                ct = 0.5 * (ck + ck1);                    // m = 2, Nz: Centered2
            }
        }
        const double Ft = (P.area[2] * cw) * ct;
        cp_async_wait_all();
        block_barrier();

        // ---- tendency, RK3 substep, advance the window ---------------------------------------------
        {
            const double Fe = lds(aT + oFX + boff + 8), Fn = lds(aT + oFY + boff + ROW_B);
            const double ue = lds(aV + oQX + boff + 8), vn = lds(aV + oQY + boff + ROW_B);
            double div_flux = (Fe - Fw);
            div_flux = div_flux + (Fn - Fs);
            div_flux = div_flux + (Ft - Fb);
            div_flux *= inv_vol;
            double div_u = (P.area[0] * ue - P.area[0] * cu);
            div_u = div_u + (P.area[1] * vn - P.area[1] * cv);
            div_u = div_u + (P.area[2] * cw - P.area[2] * wb);
            div_u *= inv_vol;
            double src;
            if (map_axis >= 0) src = map_axis == 0 ? 0.5 * (cu + ue) : (map_axis == 1 ? 0.5 * (cv + vn) : 0.5 * (wb + cw));
            else src = need_var ? lds(aIn + boff + 4 * IN_F) : 0.0;
            const double forcing = fma(k_src, src, fma(k_own, ck, k_oth * other));
            const double Gn = fma(ck, div_u, -div_flux) + forcing;
            double unew;
            if (P.first_stage) unew = fma(dtg, Gn, ck);
            else               unew = fma(dtg, Gn, fma(dtz, lds(aIn + boff + 3 * IN_F), ck));
            if (active) {
                pUn[0] = unew;
                // periodic halo images of the new state (the tracer halo fill of update_state!, fused)
                if (xwrap) pUn[xwrap] = unew;
                if (ywrap) { pUn[ywrap] = unew; if (xwrap) pUn[xwrap + ywrap] = unew; }
                if (P.store_G) pG[0] = Gn;
            }
            // advance the window to cell k+1
            Fb = Ft; wb = cw;
            const double dnew = ck4 - ck3;                // d[k+3]
            const double s = dnew - d4;
            d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = dnew;
            Qa = Qb; Qb = Qc; Qc = Qd; Qd = 3.25 * s * s;
            ck = ck1; ck1 = ck2; ck2 = ck3; ck3 = ck4;
            ck4 = lds(aIn + boff + 5 * IN_F);             // c[k+5], staged one plane ago
        }
        pU += plane; pUn += plane; pG += plane; pu += plane; pv += plane; pw += plane;
        if (need_var) pf += plane;
        h_g0 += plane; h_g1 += plane;
    }
}

#ifndef ZM_TY
#define ZM_TY 6          // tile height: 6 rows -> 480 threads per block, 136 registers per thread available
#endif

extern "C" int lfb_zmarch_supported(const LfbStageParams *P)
{
    const LfbLayout &L = P->L;
    if (!P->advect || P->relax) return 0;
    if (L.topo[0] != LFB_PERIODIC || L.topo[2] != LFB_BOUNDED) return 0;
    if (L.topo[1] != LFB_PERIODIC && !(L.connected_lo[1] && L.connected_hi[1])) return 0;
    if (!(P->vel_present[0] && P->vel_present[1] && P->vel_present[2])) return 0;
    if (L.N[0] % ZM_TX || L.N[1] < 4 || L.N[2] < 8) return 0;
    if (L.H[0] < 3 || L.H[1] < 3 || L.H[2] < 3) return 0;
    return 1;
}

template <int NSUB, int TY>
static cudaError_t zm_launch(const LfbStageParams *P, cudaStream_t stream)
{
    using SM = ZmSmem<NSUB, TY>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lfb_stage_zmarch<NSUB, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int ytiles = (P->j1 - P->j0 + TY - 1) / TY;
    dim3 grid(P->n_items * ytiles, P->L.N[0] / ZM_TX, 1);
    lfb_stage_zmarch<NSUB, TY><<<grid, SM::THREADS, SM::BYTES, stream>>>(*P);
    return cudaGetLastError();
}

extern "C" cudaError_t lfb_launch_stage_zmarch(const LfbStageParams *P, cudaStream_t stream)
{
    if (P->j1 <= P->j0) return cudaSuccess;
    return P->single_exp ? zm_launch<1, ZM_TY>(P, stream) : zm_launch<2, ZM_TY>(P, stream);
}
