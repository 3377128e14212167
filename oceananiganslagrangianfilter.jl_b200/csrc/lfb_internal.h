// lfb_internal.h -- shared host/device definitions of liblfb200 (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lfb200.h"

#define LFB_MAX_ITEMS 64   // (variables + maps) x coefficient sets handled by one stage launch
#define LFB_XPAD 16        // doubles in front of interior column 0: 128-byte aligned interior rows

// Device memory layout shared by every field of a handle (centre- and face-located alike):
// element (i,j,k), 0-based interior indices that may run into halos, lives at
//     xoff + i + pitch * ((Hy + j) + rows * (Hz + k))
// pitch is a multiple of 16 doubles and xoff = 16 (0 on a Flat x axis), so interior rows start on
// 128-byte boundaries and a warp reading 32 consecutive cells touches exactly two cache lines.
// rows/planes carry one spare entry for Face fields on Bounded axes (w: Nz+1 levels).
struct LfbLayout {
    int     N[3];        // local interior cells
    int     H[3];        // halo widths
    int     topo[3];     // global topology
    int     xoff;        // column of interior i = 0
    int     pitch;       // row pitch in elements
    int     rows;        // Ny + 2Hy + 1
    int     planes;      // Nz + 2Hz + 1
    int64_t plane;       // pitch * rows
    int64_t len;         // pitch * rows * planes
    int64_t stride[3];   // 1, pitch, plane
    // distributed y-slabs
    int     NG[3];       // global interior cells
    int     goff[3];     // global index of local cell 0
    int     connected_lo[3], connected_hi[3];   // halo on that side is owned by a neighbour rank
    int     fuse_halo[3];                       // stage kernels write the periodic halo images of this axis themselves

    __host__ __device__ inline int64_t at(int i, int j, int k) const {
        return (int64_t)(xoff + i) + (int64_t)pitch * ((int64_t)(H[1] + j) + (int64_t)rows * (int64_t)(H[2] + k));
    }
};

// one (field, coefficient set) work item: a cosine tracer and (unless single exponential) its sine partner
struct LfbItem {
    int    is_map;      // 0: filtered variable forced by the original data; 1: xi map forced by a velocity
    int    src;         // variable index, or velocity axis for maps
    int    tracer_c;    // tracer slot of the _C tracer; the _S tracer is tracer_c + 1
    int    pad;
    double c, d;        // filter coefficients of this set (d = 0 for the single exponential)
};

struct LfbStageParams {
    LfbLayout L;
    // tracer storage: U[buf][tracer], G[tracer]: contiguous blocks of L.len doubles
    const double *U_old;
    double       *U_new;
    double       *G;          // read as G^- (if use_prev) and overwritten with G^n (if store_G)
    int     n_items;
    int     single_exp;
    int     advect;           // LFB_ADVECTION_* (0 = none)
    int     relax;
    const double *imm;        // immersed-cell flags (0 / 1), or null: no immersed boundary
    const double *dzc;        // PartialCellBottom: cell heights dz(i,j,k), or null: regular cells
    // immersed boundary on the TMA z-march kernel: per cell the schemes of its west / south / bottom / top face and its
    // immersed flag, packed by lfb_build_face_codes (code = w + 8 s + 64 b + 512 t + 4096 immersed; face code 5 / 3 / 2 =
    // WENO5 / WENO3 / Centered2, 1 / 0 = lower / upper domain wall, 7 = no flux: the face touches an immersed cell)
    const double *codes;
    LfbItem items[LFB_MAX_ITEMS];
    // inputs at the stage time: u, v, w and the original variables, already interpolated in time (and
    // negated on the backward pass) by lfb_interp_inputs, or aliased to a snapshot slot when the stage
    // time sits on a snapshot
    const double *vel[3];
    const double *var[LFB_MAX_VARS];
    const double *mask;
    int     vel_present[3];
    // geometry
    double  area[3], volume, spacing[3];
    double  relax_timescale;
    // RK3 substep: U_new = U_old + dt * (gamma * G + zeta * G^-)
    double  dt, gamma, zeta;
    int     first_stage;      // zeta = nothing: U += (dt * gamma) * G
    int     store_G;          // write G^n for the next stage
    int     update;           // 0: tendencies only (lfb_get_tendency)
    // sub-range of the local grid this launch covers (boundary strips / interior when overlapping
    // the halo exchange): cells j in [j0, j1)
    int     j0, j1;
    double *scratch;          // one spare field (store target of inactive rows in partial tiles of the z-march kernels)
    // var_fused: the original variables are NOT interpolated in time by lfb_interp_inputs; var[q] / var2[q] are the two
    // bracketing snapshots and the stage kernel forms var_w1 * var[q] + var_w2 * var2[q] itself (TMA z-march kernel)
    int     var_fused;
    const double *var2[LFB_MAX_VARS];
    double  var_w1, var_w2;
    // TMA z-march kernel on a grid with a Flat y axis (the reference's 2-D examples): the x line of a plane is viewed as
    // rows of 32 cells (row r = cells [32 r, 32 r + 32), its haloed tile = cells [32 r - 4, 32 r + 36) through a tensor
    // map whose row stride is smaller than its inner extent), y faces are not computed.  fold_nx = the real Nx.  A v
    // component named in velocity_names then only forces its xi_v map (cell-centred on the Flat axis): its two
    // bracketing snapshots travel in var[LFB_MAX_VARS - 1] / var2[...], vsign = -1 on the backward pass.
    int     fold, fold_nx;
    double  vsign;
    // rows of this launch lie at least one tile row away from the walls of a Bounded y axis (and x, z need no general
    // treatment): the TMA z-march kernel may use its all-periodic instantiation although topo[1] is Bounded
    int     plain_rows;
    // x tiles (of 32 columns) this launch covers: [xt0, xt0 + xt_n); xt_n = 0: all.  plain_cols: they lie away from the
    // walls of a Bounded x axis, so the all-periodic instantiation may be used although topo[0] is Bounded
    int     xt0, xt_n, plain_cols;
    // second row range [jb0, jb1) of the same launch (the two edge strips of a slab in ONE grid); empty if jb1 <= jb0
    int     jb0, jb1;
    // side job of the TMA z-march kernel: while it computes this stage, its helper warps interpolate the velocities of
    // the NEXT stage time from the two bracketing snapshots (update_input_data!, utils:1036-1058) for the rows /
    // columns / planes the next stage's tiles will read: dst = +-(s2 * w2 + s1 * w1).  Replaces the separate
    // lfb_interp_inputs launch for every stage but the first after a (re)start.
    struct {
        int           on, negate;
        const double *s1[3], *s2[3];
        double       *dst[3];            // null: component absent / aliased to a snapshot
        double        w1, w2;
    } pf;
};

// Output combinations (create_output_fields, utils:898-1012): out = sum_i wc[i] U[tc[i]] + ws[i] U[tc[i]+1]
struct LfbOutputParams {
    LfbLayout L;
    const double *U;
    int    n_terms, single_exp;
    int    tracer_c[LFB_MAX_COEFFS];
    double wc[LFB_MAX_COEFFS], ws[LFB_MAX_COEFFS];
};

// update_input_data! (utils:1036-1058): cur = s2 * w2 + s1 * w1 for a batch of fields, optionally negated
struct LfbInterpParams {
    int           n_fields;
    int64_t       len;
    const double *s1[3 + LFB_MAX_VARS], *s2[3 + LFB_MAX_VARS];
    double       *dst[3 + LFB_MAX_VARS];
    int           negate[3 + LFB_MAX_VARS];
    int           interp;      // 0: copy s1 (only used when a negation is needed)
    double        w1, w2;
};
