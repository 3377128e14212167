// lfb_generic.cu -- topology-agnostic kernels of liblfb200: the generic fused RK3 stage kernel
// (one thread per cell, all tracers), halo fills, layout import/export, initialisation, output
// combinations and the forward+backward sum.  Written for sm_100a; FP64 throughout.
#include "lfb_internal.h"
#include "lfb_math.cuh"

// ------------------------------------------------------------------------------------------------
// Fused RK3 stage, generic flavour.
//
// Replaces, per RK3 stage and for ALL tracers in one launch, the reference's
//   update_input_data!           (utils:1036-1058; linear time interpolation of u,v,w and the
//                                 original variables, done here on the fly from two resident snapshots)
//   compute_Gc! / tracer_tendency (compute_lagrangian_filter_tendencies.jl:63-91,
//                                 lagrangian_filter_tendency_kernel_functions.jl:36-61,
//                                 lagrangian_filtering_advection_operators.jl:19-21)
//   forcing closures             (utils:487-565, 584-699)
//   _rk3_substep_field!          (lagrangian_filter_rk3_substep.jl:31-61)
//   _cache_field_tendencies!     (cache_lagrangian_filter_tendencies.jl:8-31; G is updated in place)
// U is ping-ponged (the stencil reads neighbours of U_old).  Each thread owns one cell, loads the six
// face velocities once and loops over the (field, coefficient set) items.
// ------------------------------------------------------------------------------------------------
template <class T, bool STRICT>
__global__ void __launch_bounds__(128) lfb_stage_generic(const __grid_constant__ LfbStageParams P)
{
    const LfbLayout &L = P.L;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = P.j0 + blockIdx.y;
    const int k = blockIdx.z;
    if (i >= L.N[0] || j >= P.j1) return;
    const int ijk[3] = {i, j, k};
    const int64_t idx = L.at(i, j, k);

    // face areas and cell volume: regular, or Oceananigans' partial-cell metrics (PartialCellBottom): dz at an x / y
    // face = min of the two adjacent cell heights
    T a_lo[3], a_hi[3], vol(P.volume);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) a_lo[ax] = a_hi[ax] = T(P.area[ax]);
    if (P.dzc) {
        const double *__restrict__ hz = P.dzc + idx;
        if (L.topo[0] != LFB_FLAT) { a_lo[0] = T(P.spacing[1]) * T(fmin(hz[-1], hz[0])); a_hi[0] = T(P.spacing[1]) * T(fmin(hz[0], hz[1])); }
        if (L.topo[1] != LFB_FLAT) {
            a_lo[1] = T(P.spacing[0]) * T(fmin(hz[-L.stride[1]], hz[0])); a_hi[1] = T(P.spacing[0]) * T(fmin(hz[0], hz[L.stride[1]]));
        }
        vol = T(P.area[2]) * T(hz[0]);
    }
    // face velocities (low / high face of this cell per axis) and centred velocities for the maps
    T q_lo[3], q_hi[3], q_c[3];
    T div_u(0.0);
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        q_lo[ax] = q_hi[ax] = q_c[ax] = T(0.0);
        if (!P.vel_present[ax]) continue;
        const T lo(P.vel[ax][idx]);
        if (L.topo[ax] == LFB_FLAT) { q_c[ax] = lo; continue; }
        const T hi(P.vel[ax][idx + L.stride[ax]]);
        q_lo[ax] = lo; q_hi[ax] = hi;
        q_c[ax] = (lo + hi) / T(2.0);
        if (P.advect) div_u = div_u + (a_hi[ax] * hi - a_lo[ax] * lo);
    }
    div_u = div_u / vol;

    // periodic images of this cell in the halo: per axis an offset (0 = none); all 2^n - 1 combinations are written
    int64_t wrap[3] = {0, 0, 0};
    int wrap_mask = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (!L.fuse_halo[ax]) continue;
        if (ijk[ax] < L.H[ax]) wrap[ax] = (int64_t)L.N[ax] * L.stride[ax];
        else if (ijk[ax] >= L.N[ax] - L.H[ax]) wrap[ax] = -(int64_t)L.N[ax] * L.stride[ax];
        if (wrap[ax]) wrap_mask |= 1 << ax;
    }
    const bool nwrap = wrap_mask != 0;
    // ImmersedBoundaryGrid: flags of this cell and (through the strides) its neighbours
    const double *__restrict__ imm = P.imm ? P.imm + idx : nullptr;
    const bool cell_immersed = imm && imm[0] != 0.0;

    const int n_sub = P.single_exp ? 1 : 2;
    for (int it = 0; it < P.n_items; ++it) {
        const LfbItem item = P.items[it];
        const T c(item.c), d(item.d);
        const T n2 = c * c + d * d;
        const double *Uc = P.U_old + (int64_t)item.tracer_c * L.len;
        const double *Us = Uc + L.len;
        const T gC(Uc[idx]);
        const T gS = P.single_exp ? T(0.0) : T(Us[idx]);
        T src;
        if (item.is_map) src = q_c[item.src];
        else src = T(P.var[item.src][idx]);

        for (int sub = 0; sub < n_sub; ++sub) {
            const double *U = sub ? Us : Uc;
            const T cval = sub ? gS : gC;
            // advection: -div(U c) + c div(U)
            T div_flux(0.0);
            if (P.advect) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    if (L.topo[ax] == LFB_FLAT || !P.vel_present[ax]) continue;
                    const int64_t s = L.stride[ax];
                    const int m = L.goff[ax] + ijk[ax] + 1;
                    T c_lo = face_value<T, STRICT>(U + idx, s, m, L.NG[ax], L.topo[ax], q_lo[ax], P.advect, imm);
                    T c_hi = face_value<T, STRICT>(U + idx + s, s, m + 1, L.NG[ax], L.topo[ax], q_hi[ax], P.advect, imm ? imm + s : nullptr);
                    T F_lo = a_lo[ax] * q_lo[ax] * c_lo;
                    T F_hi = a_hi[ax] * q_hi[ax] * c_hi;
                    // conditional flux: nothing is advected through a face that touches an immersed cell
                    if (imm && (imm[-s] != 0.0 || imm[0] != 0.0)) F_lo = T(0.0);
                    if (imm && (imm[0] != 0.0 || imm[s] != 0.0)) F_hi = T(0.0);
                    div_flux = div_flux + (F_hi - F_lo);
                }
                div_flux = div_flux / vol;
            }
            // forcing (create_forcing, utils:731-865)
            T forcing;
            if (P.single_exp) {
                forcing = item.is_map ? (-c / n2 * src) + (-c * gC) : (-c * gC) + src;
            } else if (sub == 0) {
                forcing = item.is_map ? (-c / n2 * src) + (-c * gC - d * gS) : (-c * gC - d * gS) + src;
            } else {
                forcing = item.is_map ? (-d / n2 * src) + (-c * gS + d * gC) : (-c * gS + d * gC);
            }
            if (P.relax) {
                T geq;
                if (P.single_exp)   geq = item.is_map ? (T(-1.0) / (c * c)) * src : src / c;
                else if (sub == 0)  geq = item.is_map ? src * (d * d - c * c) / (n2 * n2) : src * c / n2;
                else                geq = item.is_map ? src * (T(-2.0) * c * d) / (n2 * n2) : src * d / n2;
                forcing = forcing + T(-1.0) / T(P.relax_timescale) * (cval - geq) * T(P.mask[idx]);
            }
            const T Gn = (-div_flux + cval * div_u) + forcing;

            const int64_t tix = (int64_t)(item.tracer_c + sub) * L.len + idx;
            if (P.update) {
                T unew;
                if (P.first_stage) unew = cval + T(P.dt) * T(P.gamma) * Gn;
                else               unew = cval + T(P.dt) * (T(P.gamma) * Gn + T(P.zeta) * T(P.G[tix]));
                if (cell_immersed) unew = T(0.0);          // mask_immersed_field! (update_lagrangian_filter_state.jl:22-24)
                P.U_new[tix] = val(unew);
                // periodic halo images of the new state (tracer halo fill of update_state!, fused)
                if (nwrap) {
                    double *base = P.U_new + tix;
                    for (int w = 1; w < 8; ++w) {
                        if ((w & wrap_mask) != w) continue;
                        int64_t o = 0;
                        if (w & 1) o += wrap[0];
                        if (w & 2) o += wrap[1];
                        if (w & 4) o += wrap[2];
                        base[o] = val(unew);
                    }
                }
            }
            if (P.store_G) P.G[tix] = val(Gn);
        }
    }
}

extern "C" cudaError_t lfb_launch_stage_generic(const LfbStageParams *P, int strict, cudaStream_t stream)
{
    const LfbLayout &L = P->L;
    if (P->j1 <= P->j0) return cudaSuccess;
    dim3 block(128, 1, 1);
    dim3 grid((L.N[0] + block.x - 1) / block.x, P->j1 - P->j0, L.N[2]);
    if (strict) lfb_stage_generic<sd, true><<<grid, block, 0, stream>>>(*P);
    else        lfb_stage_generic<double, false><<<grid, block, 0, stream>>>(*P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// update_input_data! (utils:1036-1058): fields at the stage time from the two bracketing snapshots,
// psi2 * n~ + psi1 * (1 - n~), individually rounded like the reference's broadcast; velocities are
// negated on the backward pass (utils:169-173).  One launch for all input fields.
// ------------------------------------------------------------------------------------------------
__global__ void lfb_interp_inputs(const __grid_constant__ LfbInterpParams P)
{
    const int f = blockIdx.y;
    const double *__restrict__ s1 = P.s1[f], *__restrict__ s2 = P.s2[f];
    double *__restrict__ dst = P.dst[f];
    const bool neg = P.negate[f] != 0;
    int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    for (; q < P.len; q += stride) {           // len is a multiple of 16: two elements per thread, 16-byte accesses
        const double2 a = *reinterpret_cast<const double2 *>(s1 + q);
        double2 r = a;
        if (P.interp) {
            const double2 b = *reinterpret_cast<const double2 *>(s2 + q);
            r.x = __dadd_rn(__dmul_rn(b.x, P.w2), __dmul_rn(a.x, P.w1));
            r.y = __dadd_rn(__dmul_rn(b.y, P.w2), __dmul_rn(a.y, P.w1));
        }
        if (neg) { r.x = -r.x; r.y = -r.y; }
        *reinterpret_cast<double2 *>(dst + q) = r;
    }
}

extern "C" cudaError_t lfb_launch_interp_inputs(const LfbInterpParams *P, cudaStream_t stream)
{
    if (P->n_fields == 0) return cudaSuccess;
    dim3 block(256, 1, 1), grid(148 * 8, P->n_fields, 1);
    lfb_interp_inputs<<<grid, block, 0, stream>>>(*P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Halo fill along one axis for a batch of fields (tracer halo fill of update_state!,
// update_lagrangian_filter_state.jl:33-34).  Runs over the full parent extent of
// the two other axes, like Oceananigans' fill kernels.
//   mode 0: periodic copy of H cells from the opposite side
//   mode 1: zero-flux mirror of the first halo cell only (mode 2 / 3: only on the low / high side -- the outer edge
//           of the first / last y-slab of a decomposed Bounded axis)
// n_ax is the interior extent along the axis (N, or N+1 for Face fields on Bounded axes).
// ------------------------------------------------------------------------------------------------
__global__ void lfb_fill_halo_axis(double *base, int64_t field_stride, LfbLayout L, int ax, int mode, int n_ax)
{
    const int o1 = (ax + 1) % 3, o2 = (ax + 2) % 3;
    const int e1 = L.N[o1] + 2 * L.H[o1], e2 = L.N[o2] + 2 * L.H[o2];
    const int q1 = blockIdx.x * blockDim.x + threadIdx.x;
    const int q2 = blockIdx.y;
    if (q1 >= e1 || q2 >= e2) return;
    double *f = base + (int64_t)blockIdx.z * field_stride;
    int c[3];
    c[o1] = q1 - L.H[o1];
    c[o2] = q2 - L.H[o2];
    const int H = L.H[ax];
    if (mode == 0) {
        for (int h = 0; h < H; ++h) {
            c[ax] = -H + h;      int64_t dl = L.at(c[0], c[1], c[2]);
            c[ax] = n_ax - H + h; int64_t sl = L.at(c[0], c[1], c[2]);
            f[dl] = f[sl];
            c[ax] = n_ax + h;    int64_t dr = L.at(c[0], c[1], c[2]);
            c[ax] = h;           int64_t sr = L.at(c[0], c[1], c[2]);
            f[dr] = f[sr];
        }
    } else {
        if (mode != 3) {
            c[ax] = -1;       int64_t dl = L.at(c[0], c[1], c[2]);
            c[ax] = 0;        int64_t sl = L.at(c[0], c[1], c[2]);
            f[dl] = f[sl];
        }
        if (mode != 2) {
            c[ax] = n_ax;     int64_t dr = L.at(c[0], c[1], c[2]);
            c[ax] = n_ax - 1; int64_t sr = L.at(c[0], c[1], c[2]);
            f[dr] = f[sr];
        }
    }
}

extern "C" cudaError_t lfb_launch_fill_halo_axis(double *base, int64_t field_stride, int n_fields, const LfbLayout *L,
                                                 int ax, int mode, cudaStream_t stream)
{
    if (L->H[ax] == 0 || n_fields == 0) return cudaSuccess;
    const int o1 = (ax + 1) % 3, o2 = (ax + 2) % 3;
    const int e1 = L->N[o1] + 2 * L->H[o1], e2 = L->N[o2] + 2 * L->H[o2];
    dim3 block(128, 1, 1), grid((e1 + 127) / 128, e2, n_fields);
    lfb_fill_halo_axis<<<grid, block, 0, stream>>>(base, field_stride, *L, ax, mode, L->N[ax]);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Layout conversion between Oceananigans' parent arrays (dense, extents e = n + 2H) and the padded
// device layout.  TIn/TOut allow Float32 files / writers (the reference's JLD2 data is Float32).
// ------------------------------------------------------------------------------------------------
template <class TIn>
__global__ void lfb_import_parent(double *dst, const TIn *src, LfbLayout L, int ex, int ey, int ez)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x, pj = blockIdx.y, pk = blockIdx.z;
    if (pi >= ex) return;
    dst[L.at(pi - L.H[0], pj - L.H[1], pk - L.H[2])] = (double)src[(int64_t)pi + (int64_t)ex * (pj + (int64_t)ey * pk)];
}

template <class TOut>
__global__ void lfb_export_parent(TOut *dst, const double *src, LfbLayout L, int ex, int ey, int ez)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x, pj = blockIdx.y, pk = blockIdx.z;
    if (pi >= ex) return;
    dst[(int64_t)pi + (int64_t)ex * (pj + (int64_t)ey * pk)] = (TOut)src[L.at(pi - L.H[0], pj - L.H[1], pk - L.H[2])];
}

extern "C" cudaError_t lfb_launch_import(double *dst, const void *src, int is_f32, const LfbLayout *L,
                                         int ex, int ey, int ez, cudaStream_t stream)
{
    dim3 block(128, 1, 1), grid((ex + 127) / 128, ey, ez);
    if (is_f32) lfb_import_parent<float><<<grid, block, 0, stream>>>(dst, (const float *)src, *L, ex, ey, ez);
    else        lfb_import_parent<double><<<grid, block, 0, stream>>>(dst, (const double *)src, *L, ex, ey, ez);
    return cudaGetLastError();
}

extern "C" cudaError_t lfb_launch_export(void *dst, int is_f32, const double *src, const LfbLayout *L,
                                         int ex, int ey, int ez, cudaStream_t stream)
{
    dim3 block(128, 1, 1), grid((ex + 127) / 128, ey, ez);
    if (is_f32) lfb_export_parent<float><<<grid, block, 0, stream>>>((float *)dst, src, *L, ex, ey, ez);
    else        lfb_export_parent<double><<<grid, block, 0, stream>>>((double *)dst, src, *L, ex, ey, ez);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Cell-centred velocity: Field(@at (Center, Center, Center) vel) on the interior (utils:1119,1129).
// ------------------------------------------------------------------------------------------------
__global__ void lfb_centre_velocity(double *dst, const double *vel, LfbLayout L, int ax, double sign)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.N[0]) return;
    const int64_t idx = L.at(i, j, k);
    double v = (L.topo[ax] == LFB_FLAT) ? vel[idx] : __ddiv_rn(__dadd_rn(vel[idx], vel[idx + L.stride[ax]]), 2.0);
    dst[idx] = sign < 0 ? -v : v;
}

extern "C" cudaError_t lfb_launch_centre_velocity(double *dst, const double *vel, const LfbLayout *L, int ax, double sign,
                                                  cudaStream_t stream)
{
    dim3 block(128, 1, 1), grid((L->N[0] + 127) / 128, L->N[1], L->N[2]);
    lfb_centre_velocity<<<grid, block, 0, stream>>>(dst, vel, *L, ax, sign);
    return cudaGetLastError();
}

// Face codes of an immersed grid for the TMA z-march kernel (LfbStageParams::codes): per cell the scheme of its west,
// south, bottom and top face -- the same chain selection as the generic kernel (lfb_face_kind, WENO(order=5)) -- the domain
// walls, the faces no flux crosses (they touch an immersed cell) and the cell's immersed flag.  Cells one column / row
// beyond the interior are coded too (east / north faces of the last tile column / row).
__global__ void lfb_build_face_codes(double *codes, const double *imm, LfbLayout L)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i > L.N[0] + 1) return;
    const int ijk[3] = {i, j, k};
    const int64_t idx = L.at(i, j, k);
    auto face = [&](int ax, int up) -> int {          // low face of the cell (up = 0) or of the cell above it along ax
        if (L.topo[ax] == LFB_FLAT) return 5;
        const int64_t s = L.stride[ax];
        const int m = L.goff[ax] + ijk[ax] + 1 + up;
        const double *f = imm + idx + up * s;
        if (L.topo[ax] == LFB_BOUNDED && (m < 1 || m > L.NG[ax] + 1)) return 5;        // beyond the domain: never used
        if (f[-s] != 0.0 || f[0] != 0.0) return 7;
        if (L.topo[ax] == LFB_BOUNDED && m == 1) return 1;
        if (L.topo[ax] == LFB_BOUNDED && m == L.NG[ax] + 1) return 0;
        const int kind = lfb_face_kind(LFB_ADVECTION_WENO5, m, L.NG[ax], L.topo[ax], f, s);
        return kind == LFB_K_W5 ? 5 : kind == LFB_K_W3 ? 3 : 2;
    };
    const int code = face(0, 0) + 8 * face(1, 0) + 64 * face(2, 0) + 512 * face(2, 1) + 4096 * (imm[idx] != 0.0 ? 1 : 0);
    codes[idx] = (double)code;
}

extern "C" cudaError_t lfb_launch_build_face_codes(double *codes, const double *imm, const LfbLayout *L, cudaStream_t stream)
{
    dim3 block(128, 1, 1), grid((L->N[0] + 2 + 127) / 128, L->N[1] + (L->topo[1] == LFB_FLAT ? 0 : 2), L->N[2]);
    lfb_build_face_codes<<<grid, block, 0, stream>>>(codes, imm, *L);
    return cudaGetLastError();
}

// mask_immersed_field! over a stack of fields (whole padded arrays, halos too): zero where the flag is set
__global__ void lfb_mask_immersed(double *U, const double *imm, int64_t len, int n_fields)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; q < len; q += stride)
        if (imm[q] != 0.0)
            for (int f = 0; f < n_fields; ++f) U[(int64_t)f * len + q] = 0.0;
}

extern "C" cudaError_t lfb_launch_mask_immersed(double *U, const double *imm, int64_t len, int n_fields, cudaStream_t stream)
{
    int blocks = (int)((len + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    lfb_mask_immersed<<<blocks, 256, 0, stream>>>(U, imm, len, n_fields);
    return cudaGetLastError();
}

// dst = factor * src over the whole array (parent(field) .= factor * parent(src), utils:1096-1131,1239-1247)
__global__ void lfb_scale(double *dst, const double *src, double factor, int64_t n)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; q < n; q += stride) dst[q] = __dmul_rn(factor, src[q]);
}

extern "C" cudaError_t lfb_launch_scale(double *dst, const double *src, double factor, int64_t n, cudaStream_t stream)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    lfb_scale<<<blocks, 256, 0, stream>>>(dst, src, factor, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Output combinations (create_output_fields, utils:898-1012), interior only:
//   out = sum_i  wc[i] * U[tc[i]] + ws[i] * U[tc[i] + 1]
// with (wc, ws) = (a, b) for filtered variables and maps, (-a c + b d, -a d - b c) for mean velocities.
// Evaluated left to right in unfused FP64 like the reference's AbstractOperation tree.
// ------------------------------------------------------------------------------------------------
__global__ void lfb_output_combine(double *out, const __grid_constant__ LfbOutputParams P)
{
    const LfbLayout &L = P.L;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.N[0]) return;
    const int64_t idx = L.at(i, j, k);
    double tot = 0.0;
    for (int t = 0; t < P.n_terms; ++t) {
        const double *Uc = P.U + (int64_t)P.tracer_c[t] * L.len;
        double term = __dmul_rn(P.wc[t], Uc[idx]);
        if (!P.single_exp) term = __dadd_rn(term, __dmul_rn(P.ws[t], Uc[L.len + idx]));
        tot = (t == 0) ? term : __dadd_rn(tot, term);
    }
    out[idx] = tot;
}

extern "C" cudaError_t lfb_launch_output(double *out, const LfbOutputParams *P, cudaStream_t stream)
{
    const LfbLayout &L = P->L;
    dim3 block(128, 1, 1), grid((L.N[0] + 127) / 128, L.N[1], L.N[2]);
    lfb_output_combine<<<grid, block, 0, stream>>>(out, *P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Forward + backward combination (sum_forward_backward_contributions!, post:135-164) for one field
// at one output time: out = fwd + sign * (bwd_hi * w + bwd_lo * (1 - w)).
// ------------------------------------------------------------------------------------------------
__global__ void lfb_combine_kernel(double *out, const double *fwd, const double *lo, const double *hi,
                                   double w, double sign, int64_t n)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; q < n; q += stride) {
        double b = (w == 0.0) ? lo[q] : __dadd_rn(__dmul_rn(hi[q], w), __dmul_rn(lo[q], 1.0 - w));
        out[q] = sign < 0 ? __dsub_rn(fwd[q], b) : __dadd_rn(fwd[q], b);
    }
}

extern "C" cudaError_t lfb_launch_combine(double *out, const double *fwd, const double *lo, const double *hi,
                                          double w, double sign, int64_t n, cudaStream_t stream)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    lfb_combine_kernel<<<blocks, 256, 0, stream>>>(out, fwd, lo, hi, w, sign, n);
    return cudaGetLastError();
}

// ... on device-resident pass outputs, each Float32 (as the reference's JLD2Writer rounds them) or Float64
template <class A, class B, class C2>
__global__ void lfb_combine_mixed_kernel(double *out, const A *fwd, const B *lo, const C2 *hi, double w, double sign, int64_t n)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; q < n; q += stride) {
        const double b = (w == 0.0) ? (double)lo[q] : __dadd_rn(__dmul_rn((double)hi[q], w), __dmul_rn((double)lo[q], 1.0 - w));
        out[q] = sign < 0 ? __dsub_rn((double)fwd[q], b) : __dadd_rn((double)fwd[q], b);
    }
}

extern "C" cudaError_t lfb_launch_combine_mixed(double *out, const void *fwd, int fwd_f32, const void *lo, int lo_f32, const void *hi,
                                                int hi_f32, double w, double sign, int64_t n, cudaStream_t stream)
{
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (fwd_f32 != lo_f32 || fwd_f32 != hi_f32) return cudaErrorInvalidValue;     // both passes are written with one writer
    if (fwd_f32) lfb_combine_mixed_kernel<<<blocks, 256, 0, stream>>>(out, (const float *)fwd, (const float *)lo, (const float *)hi, w, sign, n);
    else         lfb_combine_mixed_kernel<<<blocks, 256, 0, stream>>>(out, (const double *)fwd, (const double *)lo, (const double *)hi, w, sign, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// y-halo pack / unpack for the slab exchange: H rows next to each slab edge for a batch of fields,
// full x extent (with x halos) and full z extent, into / out of a dense buffer
//   buf[((field * planes_z + pk) * H + h) * ex + pi]
// side 0: low edge (rows 0..H-1 -> neighbour's high halo), side 1: high edge.
// ------------------------------------------------------------------------------------------------
__global__ void lfb_pack_y(double *buf, const double *base, int64_t field_stride, LfbLayout L, int side, int unpack)
{
    const int ex = L.N[0] + 2 * L.H[0];
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= ex) return;
    const int ez = L.N[2] + 2 * L.H[2];
    const int H = L.H[1];
    const int pk = blockIdx.y % ez, h = blockIdx.y / ez;
    const int fld = blockIdx.z;
    const int64_t b = (((int64_t)fld * ez + pk) * H + h) * ex + pi;
    int j;
    if (!unpack) j = side == 0 ? h : L.N[1] - H + h;          // interior rows at the edge
    else         j = side == 0 ? -H + h : L.N[1] + h;         // halo rows
    double *f = const_cast<double *>(base) + (int64_t)fld * field_stride;
    const int64_t idx = L.at(pi - L.H[0], j, pk - L.H[2]);
    if (unpack) f[idx] = buf[b];
    else        buf[b] = f[idx];
}

extern "C" cudaError_t lfb_launch_pack_y(double *buf, double *base, int64_t field_stride, int n_fields, const LfbLayout *L,
                                         int side, int unpack, cudaStream_t stream)
{
    const int ex = L->N[0] + 2 * L->H[0], ez = L->N[2] + 2 * L->H[2];
    dim3 block(128, 1, 1), grid((ex + 127) / 128, ez * L->H[1], n_fields);
    lfb_pack_y<<<grid, block, 0, stream>>>(buf, base, field_stride, *L, side, unpack);
    return cudaGetLastError();
}
