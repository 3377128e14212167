// lfb_zsplit.cu -- role-split variant of the z-marching RK3 stage kernel.
//
// Same arithmetic, tile geometry and data flow as lfb_stage_zmarch (see lfb_zmarch.cu), but the three face
// directions of a cell are computed by three different threads so that the latency of the long FP64
// dependency chains of WENO5-Z is hidden by twice as many resident warps:
//   * z warpgroups   (one thread per cell column): register z window, top face, tendency, RK3 substep,
//                     stores, tile upkeep (own value + halo of the next plane);
//   * x warpgroups   west face of every cell from the shared tile -> flux + face velocity to shared memory;
//   * y warpgroups   south face likewise;
//   * one extra warpgroup for the faces on the east / north edge of the tile.
// Warpgroups re-balance the register file with setmaxnreg (the x / y threads need no window), one bar.sync
// per plane: the face warps run ahead into plane k+1 while the z warps finish the update of plane k.
// Tile: 32 x 4 cells x NSUB tracers; 7 warpgroups (896 threads) for NSUB = 2, 4 (512 threads) for NSUB = 1.
#include "lfb_zm_common.cuh"

#define ZS_TY 4
#define ZS_WG 128                       // threads per warpgroup = cells of one tracer tile (32 x 4)

template <int NSUB>
struct ZsSmem {
    static constexpr int CH = ZS_TY + 6;
    static constexpr int TILE = CH * ZM_CW;
    static constexpr int NZ = NSUB * ZS_WG;                      // z-role threads
    static constexpr int NFACE = 2 * NSUB * ZS_WG + ZS_WG;       // x, y and extra-face threads
    static constexpr int THREADS = NZ + NFACE;
    static constexpr int OFF_C = 0;                              // C [NSUB][TILE]
    static constexpr int OFF_FX = OFF_C + NSUB * TILE;           // Fx[NSUB][TILE]
    static constexpr int OFF_FY = OFF_FX + NSUB * TILE;          // Fy[NSUB][TILE]
    static constexpr int OFF_QX = OFF_FY + NSUB * TILE;          // Qx[TILE]
    static constexpr int OFF_QY = OFF_QX + TILE;                 // Qy[TILE]
    static constexpr int OFF_IN = OFF_QY + TILE;                 // In[4][NZ]: w, G-, f, c[k+5] of the z threads
    static constexpr int OFF_VF = OFF_IN + 4 * NZ;               // Vf[NFACE]: face velocity of the face threads
    static constexpr int BUF = (OFF_VF + NFACE + 15) / 16 * 16;
    static constexpr int BYTES = 2 * BUF * 8;
    // register split: launch gives floor(65536 / THREADS / 8) * 8 per thread; face threads shrink to RF, z threads grow
    static constexpr int R0 = (65536 / THREADS) / 8 * 8;
    static constexpr int RF = NSUB == 2 ? 56 : 64;
    static constexpr int RZ = ((R0 * THREADS - RF * NFACE) / NZ) / 8 * 8 > 232 ? 232 : ((R0 * THREADS - RF * NFACE) / NZ) / 8 * 8;
};

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

template <int NSUB>
__global__ void __launch_bounds__(ZsSmem<NSUB>::THREADS, 1) lfb_stage_zsplit(const __grid_constant__ LfbStageParams P)
{
    using SM = ZsSmem<NSUB>;
    extern __shared__ __align__(16) double zs_smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(zs_smem);
    constexpr unsigned BUF_B = 8u * SM::BUF;
    constexpr int ROW_B = 8 * ZM_CW;
    constexpr unsigned oC = 8u * SM::OFF_C, oFX = 8u * SM::OFF_FX, oFY = 8u * SM::OFF_FY, oQX = 8u * SM::OFF_QX,
                       oQY = 8u * SM::OFF_QY;
    const LfbLayout &L = P.L;
    const int tid = threadIdx.x;
    const int item_id = blockIdx.x % P.n_items;
    const int ytile = blockIdx.x / P.n_items;
    const int i0 = blockIdx.y * ZM_TX;
    const int j0 = P.j0 + ytile * ZS_TY;
    const LfbItem item = P.items[item_id];
    const int Nz = L.N[2];
    const int64_t plane = L.plane;
    const int jmax = L.N[1] + L.H[1];

    if (tid >= SM::NZ) {
        // ================= face warpgroups: one in-plane face per thread =================
        reg_dec<SM::RF>();
        const int ft = tid - SM::NZ;
        int sub, tx, ty, kind;                   // kind 1: x face (west of cell tx), 2: y face (south of cell ty), 0: idle
        if (ft < NSUB * ZS_WG) { kind = 1; sub = ft / ZS_WG; tx = ft % ZM_TX; ty = (ft % ZS_WG) / ZM_TX; }
        else if (ft < 2 * NSUB * ZS_WG) { const int q = ft - NSUB * ZS_WG; kind = 2; sub = q / ZS_WG; tx = q % ZM_TX; ty = (q % ZS_WG) / ZM_TX; }
        else {
            const int q = ft - 2 * NSUB * ZS_WG;
            if (q < NSUB * 32) { kind = 2; sub = q / 32; tx = q % 32; ty = ZS_TY; }                 // north edge
            else if (q < NSUB * 32 + NSUB * ZS_TY) { const int r = q - NSUB * 32; kind = 1; sub = r / ZS_TY; ty = r % ZS_TY; tx = ZM_TX; }   // east edge
            else { kind = 0; sub = 0; tx = 0; ty = 0; }
        }
        const unsigned aT = sbase + 8u * (sub * SM::TILE + (ty + 3) * ZM_CW + tx + 3);
        const unsigned aF = aT + (kind == 2 ? oFY : oFX);
        const unsigned aQ = sbase + 8u * ((ty + 3) * ZM_CW + tx + 3) + (kind == 2 ? oQY : oQX);
        const unsigned aVf = sbase + 8u * (SM::OFF_VF + ft);
        const double *__restrict__ vel = (kind == 2 ? P.vel[1] : P.vel[0]) + L.at(i0 + tx, min(j0 + ty, jmax), 0);
        const double area = kind == 2 ? P.area[1] : P.area[0];
        if (kind) cp_async8(aVf, vel);
        cp_async_wait_all();
        block_barrier();                                            // prologue
        for (int k = 0; k < Nz; ++k) {
            const unsigned boff = (k & 1) * BUF_B, noff = BUF_B - boff;
            if (kind) {
                vel += plane;
                if (k + 1 < Nz) cp_async8(aVf + noff, vel);
                const double q = lds(aVf + boff);
                const double cm = lds(aT + oC + boff);
                const bool left = q > 0.0;
                double cf;
                if (kind == 2) { const ZmLine l = zm_load_line<ROW_B>(aT + oC + boff, left); cf = zm_face(l, cm, left); }
                else           { const ZmLine l = zm_load_line<8>(aT + oC + boff, left);     cf = zm_face(l, cm, left); }
                sts(aF + boff, (area * q) * cf);
                if (sub == 0) sts(aQ + boff, q);
            }
            cp_async_wait_all();
            block_barrier();
        }
        return;
    }

    // ================= z warpgroups: one cell column of one tracer per thread =================
    reg_inc<SM::RZ>();
    const int tx = tid % ZM_TX, ty = (tid / ZM_TX) % ZS_TY, sub = tid / ZS_WG;
    const unsigned aT = sbase + 8u * (sub * SM::TILE + (ty + 3) * ZM_CW + tx + 3);
    const unsigned aV = sbase + 8u * ((ty + 3) * ZM_CW + tx + 3);
    const unsigned aIn = sbase + 8u * (SM::OFF_IN + tid);
    constexpr unsigned IN_F = 8u * SM::NZ;
    const int partner_off = (NSUB == 2) ? (sub ? -8 * SM::TILE : 8 * SM::TILE) : 0;
    const bool active = j0 + ty < P.j1;

    const int64_t off0 = L.at(i0 + tx, min(j0 + ty, jmax), 0);
    const int64_t tbase = (int64_t)(item.tracer_c + sub) * L.len;
    const double *__restrict__ pU = P.U_old + tbase + off0;
    double *__restrict__ pUn = P.U_new + tbase + off0;
    double *__restrict__ pG = P.G + tbase + off0;
    const double *__restrict__ pw = P.vel[2] + off0;
    const bool need_var = !item.is_map && sub == 0;
    const double *__restrict__ pf = need_var ? P.var[item.src] + off0 : nullptr;
    const bool use_G = !P.first_stage;
    int64_t xwrap = 0, ywrap = 0;
    {
        const int i = i0 + tx, j = j0 + ty;
        if (L.fuse_halo[0]) { if (i < L.H[0]) xwrap = L.N[0]; else if (i >= L.N[0] - L.H[0]) xwrap = -L.N[0]; }
        if (L.fuse_halo[1]) {
            if (j < L.H[1]) ywrap = (int64_t)L.N[1] * L.pitch; else if (j >= L.N[1] - L.H[1]) ywrap = -(int64_t)L.N[1] * L.pitch;
        }
    }
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

    // halo duties of the tile: (38 x 10 - 32 x 4) = 252 elements per tracer, two per z thread
    unsigned h_s0 = 0, h_s1 = 0;
    const double *h_g0 = nullptr, *h_g1 = nullptr;
    {
        const int t = tid % ZS_WG;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int e = t + r * ZS_WG;
            int row = -1, cc = 0;
            if (e < 3 * ZM_CW) { row = e / ZM_CW; cc = e % ZM_CW; }
            else if (e < 6 * ZM_CW) { row = ZS_TY + 3 + (e - 3 * ZM_CW) / ZM_CW; cc = (e - 3 * ZM_CW) % ZM_CW; }
            else if (e < 6 * ZM_CW + 3 * ZS_TY) { row = 3 + (e - 6 * ZM_CW) / 3; cc = (e - 6 * ZM_CW) % 3; }
            else if (e < 6 * ZM_CW + 6 * ZS_TY) { row = 3 + (e - 6 * ZM_CW - 3 * ZS_TY) / 3; cc = ZM_TX + 3 + (e - 6 * ZM_CW - 3 * ZS_TY) % 3; }
            if (row >= 0) {
                const unsigned sa = sbase + oC + 8u * (sub * SM::TILE + row * ZM_CW + cc);
                const double *ga = P.U_old + tbase + L.at(i0 - 3 + cc, min(j0 - 3 + row, jmax), 0);
                if (r == 0) { h_s0 = sa; h_g0 = ga; } else { h_s1 = sa; h_g1 = ga; }
            }
        }
    }
    static_assert(6 * ZM_CW + 6 * ZS_TY <= 2 * ZS_WG, "tile halo does not fit two duties per thread");

    double d0, d1, d2, d3, d4, Qa, Qb, Qc, Qd, ck, ck1, ck2, ck3, ck4, Fb, wb;
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
        Fb = (P.area[2] * wb) * ck;
        sts(aT + oC, ck);
        if (h_s0) cp_async8(h_s0, h_g0);
        if (h_s1) cp_async8(h_s1, h_g1);
        cp_async8(aIn, pw + plane);
        if (use_G) cp_async8(aIn + IN_F, pG);
        if (need_var) cp_async8(aIn + 2 * IN_F, pf);
        cp_async8(aIn + 3 * IN_F, pU + 5 * plane);
        cp_async_wait_all();
    }
    block_barrier();

    for (int k = 0; k < Nz; ++k) {
        const unsigned boff = (k & 1) * BUF_B, noff = BUF_B - boff;
        if (k + 1 < Nz) {
            sts(aT + oC + noff, ck1);
            if (h_s0) cp_async8(h_s0 + noff, h_g0 + plane);
            if (h_s1) cp_async8(h_s1 + noff, h_g1 + plane);
            cp_async8(aIn + noff, pw + 2 * plane);
            if (use_G) cp_async8(aIn + noff + IN_F, pG + plane);
            if (need_var) cp_async8(aIn + noff + 2 * IN_F, pf + plane);
            if (k + 6 <= Nz + 2) cp_async8(aIn + noff + 3 * IN_F, pU + 6 * plane);
        }
        const double cw = lds(aIn + boff);
        double other = 0.0;
        if (NSUB == 2) other = lds(aT + oC + boff + partner_off);
        double ct;
        {
            const int m = k + 2;
            const bool left = cw > 0.0;
            if (m >= 4 && m <= Nz - 2) {
                const double inc = zm_inc_q(left ? d0 : d4, left ? d1 : d3, d2, left ? d3 : d1,
                                            left ? Qa : Qd, left ? Qb : Qc, left ? Qc : Qb);
                ct = left ? ck + inc : ck1 - inc;
            } else if (m == Nz + 1) {
                ct = ck;
            } else if (m == 3 || m == Nz - 1) {
                const double inc = weno3z_inc(left ? d1 : d3, d2);
                ct = left ? ck + inc : ck1 - inc;
            } else {
                ct = 0.5 * (ck + ck1);
            }
        }
        const double Ft = (P.area[2] * cw) * ct;
        cp_async_wait_all();
        block_barrier();
        {
            const double Fw = lds(aT + oFX + boff), Fe = lds(aT + oFX + boff + 8);
            const double Fs = lds(aT + oFY + boff), Fn = lds(aT + oFY + boff + ROW_B);
            const double cu = lds(aV + oQX + boff), ue = lds(aV + oQX + boff + 8);
            const double cv = lds(aV + oQY + boff), vn = lds(aV + oQY + boff + ROW_B);
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
            else src = need_var ? lds(aIn + boff + 2 * IN_F) : 0.0;
            const double forcing = fma(k_src, src, fma(k_own, ck, k_oth * other));
            const double Gn = fma(ck, div_u, -div_flux) + forcing;
            double unew;
            if (P.first_stage) unew = fma(dtg, Gn, ck);
            else               unew = fma(dtg, Gn, fma(dtz, lds(aIn + boff + IN_F), ck));
            if (active) {
                pUn[0] = unew;
                if (xwrap) pUn[xwrap] = unew;
                if (ywrap) { pUn[ywrap] = unew; if (xwrap) pUn[xwrap + ywrap] = unew; }
                if (P.store_G) pG[0] = Gn;
            }
            Fb = Ft; wb = cw;
            const double dnew = ck4 - ck3;
            const double s = dnew - d4;
            d0 = d1; d1 = d2; d2 = d3; d3 = d4; d4 = dnew;
            Qa = Qb; Qb = Qc; Qc = Qd; Qd = 3.25 * s * s;
            ck = ck1; ck1 = ck2; ck2 = ck3; ck3 = ck4;
            ck4 = lds(aIn + boff + 3 * IN_F);
        }
        pU += plane; pUn += plane; pG += plane; pw += plane;
        if (need_var) pf += plane;
        h_g0 += plane; h_g1 += plane;
    }
}

template <int NSUB>
static cudaError_t zs_launch(const LfbStageParams *P, cudaStream_t stream)
{
    using SM = ZsSmem<NSUB>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(lfb_stage_zsplit<NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int ytiles = (P->j1 - P->j0 + ZS_TY - 1) / ZS_TY;
    dim3 grid(P->n_items * ytiles, P->L.N[0] / ZM_TX, 1);
    lfb_stage_zsplit<NSUB><<<grid, SM::THREADS, SM::BYTES, stream>>>(*P);
    return cudaGetLastError();
}

extern "C" cudaError_t lfb_launch_stage_zsplit(const LfbStageParams *P, cudaStream_t stream)
{
    if (P->j1 <= P->j0) return cudaSuccess;
    return P->single_exp ? zs_launch<1>(P, stream) : zs_launch<2>(P, stream);
}
