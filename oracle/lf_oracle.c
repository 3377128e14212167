/*
 * lf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the filter-PDE hot path of
 * loisbaker/OceananigansLagrangianFilter.jl v0.1.3 (reference at /root/reference),
 * written un-fused and in the reference's own order of operations so that it can
 * act as the checker for the CUDA path in liblfb200.so.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file; the product path never does.
 *
 * The arithmetic itself lives in the un-vendored dependency Oceananigans.jl
 * (compat "0.110", reference Project.toml:25), so each function cites the reference
 * call site it stands in for and SURVEY.md Appendix A, which was pinned against the
 * reference's golden file test/data/reference_offline_output.jld2 (exact after the
 * Float32 rounding the reference's writer applies).  tests/test_oracle_golden.py
 * re-checks that pin on every run.
 *
 * Array contract: every field is an Oceananigans `parent(field)` array: FP64,
 * column-major, x fastest, halos included.  A field located on a Face of a Bounded
 * axis has N+1 interior entries on that axis (e.g. w: Nz+1 levels); Flat axes have
 * size 1 and halo 0.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LFO_MAXP 8      /* max coefficient sets (N/2)            */
#define LFO_MAXV 8      /* max variables to filter               */
#define TOPO_PERIODIC 0
#define TOPO_BOUNDED  1
#define TOPO_FLAT     2

typedef struct {
    int    N[3];          /* interior cells Nx,Ny,Nz                                 */
    int    H[3];          /* halo widths (0 on Flat axes)                            */
    int    topo[3];       /* TOPO_* per axis                                         */
    double delta[3];      /* regular spacings (1 on Flat axes, as Oceananigans)      */
    int    n_vars;        /* number of variables to filter                           */
    int    vel_present[3];/* u,v,w named in velocity_names?                          */
    int    n_pairs;       /* N_coeffs (1 for the single exponential too)             */
    int    single_exp;    /* 1 <=> N_coeffs == 0.5 (only a1,c1; no S tracers)        */
    int    with_maps;     /* map_to_mean || compute_mean_velocities                  */
    int    advect;        /* 1 = WENO(order=5) default scheme, 0 = advection=nothing */
    double a[LFO_MAXP], b[LFO_MAXP], c[LFO_MAXP], d[LFO_MAXP];
    /* optional boundary relaxation (utils:584-699): mask pre-evaluated at cell centres */
    int    relax;         /* boundary_relaxation                                     */
    double relax_timescale;
    /* numerics options beyond the reference's defaults */
    int    timestepper;   /* 0 = RungeKutta3 (default, lagrangian_filter.jl:85), 1 = QuasiAdamsBashforth2 */
    double chi;           /* AB2 parameter (Oceananigans default 0.1)                */
    int    immersed;      /* 1: an immersed-cell flag field was set (lfo_set_immersed) */
    int    weights_f32;   /* experiment only: WENO5-Z nonlinear weights evaluated in Float32 */
} lfo_problem;

/* advection scheme ids carried by lfo_problem.advect (OfflineFilterConfig.advection,
 * OfflineLagrangianFilter.jl:139,246): 0 = nothing, 1 = WENO(order=5) (the default) */
#define ADV_NONE 0
#define ADV_WENO5 1
#define ADV_CENTERED2 2
#define ADV_CENTERED4 3
#define ADV_UPWIND1 4
#define ADV_UPWIND3 5
#define ADV_UPWIND5 6
#define ADV_WENO3 7

typedef struct {
    double *p;            /* parent array                                            */
    int     n[3];         /* interior size per axis (N or N+1 for bounded faces)     */
    int     s[3];         /* parent extents n+2H                                     */
    long    len;
} lfo_field;

typedef struct {
    lfo_problem P;
    int     n_maps, n_fields, n_C, n_T;
    int     map_axis[3];          /* velocity axis of map m                          */
    lfo_field *U, *G, *Gm;        /* tracers, tendencies G^n, G^-                    */
    lfo_field vel[3];             /* current (time-interpolated) u,v,w               */
    lfo_field var[LFO_MAXV];      /* current original variables (aux fields)         */
    lfo_field mask;               /* relaxation mask at centres (optional)           */
    lfo_field imm;                /* 1.0 in immersed cells (ImmersedBoundaryGrid), else 0; halos too */
    lfo_field dzc;                /* PartialCellBottom: cell heights dz(i,j,k) at centres (optional) */
    int     partial;
    double  last_dt;              /* AB2: the previous step's dt (Euler step when it changes) */
    /* input series */
    int     n_snap;
    double *snap_t;               /* times as seen by the pass (after shift/reversal) */
    double **snap_vel[3];         /* [axis][snap] parent arrays (already negated if backward) */
    double **snap_var[LFO_MAXV];
    double  time;
    long    iteration;
} lfo_state;

/* ---------------------------------------------------------------- helpers */

static void field_alloc(lfo_field *f, const lfo_problem *P, int face_axis)
{
    f->len = 1;
    for (int ax = 0; ax < 3; ++ax) {
        int n = P->N[ax];
        if (ax == face_axis && P->topo[ax] == TOPO_BOUNDED) n += 1;
        f->n[ax] = n;
        f->s[ax] = n + 2 * P->H[ax];
        f->len *= f->s[ax];
    }
    f->p = (double *)calloc((size_t)f->len, sizeof(double));
}

static inline long fidx(const lfo_field *f, const lfo_problem *P, int i, int j, int k)
{   /* i,j,k are 0-based interior indices (may run into halos) */
    return (long)(i + P->H[0]) + (long)f->s[0] * ((long)(j + P->H[1]) + (long)f->s[1] * (long)(k + P->H[2]));
}

/* a1: set_offline_BW2_filter_params, reference src/Utils/lagrangian_filter_utils.jl:251-283 */
int lfo_bw2_params(int N, double freq_c, double *a, double *b, double *c, double *d)
{
    freq_c = fabs(freq_c);
    if (N == 1) { a[0] = freq_c / 2; c[0] = freq_c; b[0] = 0; d[0] = 0; return 0; } /* N_coeffs = 0.5 */
    if (N < 0 || (N % 2)) return -1;
    int nc = N / 2;
    for (int i = 1; i <= nc; ++i) {
        double th = M_PI / 2 / N * (2 * i - 1);
        a[i - 1] = (freq_c / N) * sin(th);
        b[i - 1] = (freq_c / N) * cos(th);
        c[i - 1] = freq_c * sin(th);
        d[i - 1] = freq_c * cos(th);
    }
    return nc;
}

/* a8: tracer halo fill, reference update_lagrangian_filter_state.jl:33-34 -> Oceananigans
 * fill_halo_regions!: Periodic = copy H cells from the opposite side; Bounded (default zero-flux
 * BC) = first halo cell takes the adjacent interior value, outer halo cells untouched
 * (visible in every golden array). */
static void fill_halos(lfo_field *f, const lfo_problem *P)
{
    const int *H = P->H;
    for (int ax = 0; ax < 3; ++ax) {
        if (P->topo[ax] == TOPO_FLAT || H[ax] == 0) continue;
        int n = f->n[ax];
        /* loop over the full parent extent in the two other axes */
        int o1 = (ax + 1) % 3, o2 = (ax + 2) % 3;
        long st[3] = {1, f->s[0], (long)f->s[0] * f->s[1]};
        for (int q2 = 0; q2 < f->s[o2]; ++q2)
            for (int q1 = 0; q1 < f->s[o1]; ++q1) {
                long base = q1 * st[o1] + q2 * st[o2];
                double *line = f->p + base;   /* element m (parent index) at line[m*st[ax]] */
                long s = st[ax];
                if (P->topo[ax] == TOPO_PERIODIC) {
                    for (int h = 0; h < H[ax]; ++h) {
                        line[h * s] = line[(h + n) * s];                          /* left halo  <- right interior */
                        line[(H[ax] + n + h) * s] = line[(H[ax] + h) * s];        /* right halo <- left interior  */
                    }
                } else { /* bounded, zero-flux */
                    line[(H[ax] - 1) * s] = line[H[ax] * s];
                    line[(H[ax] + n) * s] = line[(H[ax] + n - 1) * s];
                }
            }
    }
}

/* ---------------------------------------------------------------- lifecycle */

lfo_state *lfo_create(const lfo_problem *P)
{
    lfo_state *S = (lfo_state *)calloc(1, sizeof(lfo_state));
    S->P = *P;
    S->n_maps = 0;
    if (P->with_maps)
        for (int ax = 0; ax < 3; ++ax) if (P->vel_present[ax]) S->map_axis[S->n_maps++] = ax;
    S->n_fields = P->n_vars + S->n_maps;
    S->n_C = S->n_fields * P->n_pairs;
    S->n_T = P->single_exp ? S->n_C : 2 * S->n_C;
    S->U  = (lfo_field *)calloc(S->n_T, sizeof(lfo_field));
    S->G  = (lfo_field *)calloc(S->n_T, sizeof(lfo_field));
    S->Gm = (lfo_field *)calloc(S->n_T, sizeof(lfo_field));
    for (int t = 0; t < S->n_T; ++t) {
        field_alloc(&S->U[t], P, -1); field_alloc(&S->G[t], P, -1); field_alloc(&S->Gm[t], P, -1);
    }
    for (int ax = 0; ax < 3; ++ax) field_alloc(&S->vel[ax], P, ax);
    for (int v = 0; v < P->n_vars; ++v) field_alloc(&S->var[v], P, -1);
    field_alloc(&S->mask, P, -1);
    field_alloc(&S->imm, P, -1);
    field_alloc(&S->dzc, P, -1);
    S->last_dt = INFINITY;        /* Oceananigans Clock: last_dt = Inf */
    return S;
}

static void free_series(lfo_state *S)
{
    for (int ax = 0; ax < 3; ++ax) if (S->snap_vel[ax]) {
        for (int n = 0; n < S->n_snap; ++n) free(S->snap_vel[ax][n]);
        free(S->snap_vel[ax]); S->snap_vel[ax] = NULL;
    }
    for (int v = 0; v < LFO_MAXV; ++v) if (S->snap_var[v]) {
        for (int n = 0; n < S->n_snap; ++n) free(S->snap_var[v][n]);
        free(S->snap_var[v]); S->snap_var[v] = NULL;
    }
    free(S->snap_t); S->snap_t = NULL; S->n_snap = 0;
}

void lfo_destroy(lfo_state *S)
{
    if (!S) return;
    free_series(S);
    for (int t = 0; t < S->n_T; ++t) { free(S->U[t].p); free(S->G[t].p); free(S->Gm[t].p); }
    free(S->U); free(S->G); free(S->Gm);
    for (int ax = 0; ax < 3; ++ax) free(S->vel[ax].p);
    for (int v = 0; v < S->P.n_vars; ++v) free(S->var[v].p);
    free(S->mask.p);
    free(S->imm.p);
    free(S->dzc.p);
    free(S);
}

/* OpenMP thread control for the timed CPU baseline (bench.py): launchers such as torchrun export
 * OMP_NUM_THREADS=1, which must not silently shrink the baseline. */
int lfo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n; return 1;
#endif
}

int  lfo_n_tracers(const lfo_state *S) { return S->n_T; }
long lfo_center_len(const lfo_state *S) { return S->U[0].len; }
long lfo_vel_len(const lfo_state *S, int ax) { return S->vel[ax].len; }
double lfo_time(const lfo_state *S) { return S->time; }
void lfo_reset_clock(lfo_state *S) { S->time = 0; S->iteration = 0; S->last_dt = INFINITY; }   /* run_offline...jl:93 */

/* a14: create_input_data_on_disk, utils:97-180.  `times` are the ORIGINAL file times of the
 * snapshots in [T_start,T_end] in file order.  forward: t <- t - T_start, order kept.
 * backward: order reversed, t <- T_end - t, velocities negated (utils:154-173). */
void lfo_load_series(lfo_state *S, int n_snap, const double *times, double T_start, double T_end,
                     int backward, const double *const *u, const double *const *v, const double *const *w,
                     const double *const *const *vars)
{
    free_series(S);
    S->n_snap = n_snap;
    S->snap_t = (double *)malloc(sizeof(double) * n_snap);
    const double *const *vel_in[3] = {u, v, w};
    for (int ax = 0; ax < 3; ++ax) {
        if (!S->P.vel_present[ax]) continue;
        S->snap_vel[ax] = (double **)calloc(n_snap, sizeof(double *));
    }
    for (int q = 0; q < S->P.n_vars; ++q) S->snap_var[q] = (double **)calloc(n_snap, sizeof(double *));
    for (int n = 0; n < n_snap; ++n) {
        int src = backward ? n_snap - 1 - n : n;
        S->snap_t[n] = backward ? T_end - times[src] : times[src] - T_start;
        for (int ax = 0; ax < 3; ++ax) {
            if (!S->P.vel_present[ax]) continue;
            long len = S->vel[ax].len;
            double *dst = (double *)malloc(sizeof(double) * len);
            const double *s = vel_in[ax][src];
            if (backward) for (long q = 0; q < len; ++q) dst[q] = -s[q];
            else memcpy(dst, s, sizeof(double) * len);
            S->snap_vel[ax][n] = dst;
        }
        for (int q = 0; q < S->P.n_vars; ++q) {
            long len = S->var[q].len;
            double *dst = (double *)malloc(sizeof(double) * len);
            memcpy(dst, vars[q][src], sizeof(double) * len);
            S->snap_var[q][n] = dst;
        }
    }
}

void lfo_set_mask(lfo_state *S, const double *mask_parent)
{
    memcpy(S->mask.p, mask_parent, sizeof(double) * S->mask.len);
}

/* a7: update_input_data!, utils:1036-1058: parent(field) .= parent(fts[Time(t)]) with
 * Oceananigans' linear time interpolation psi2*n + psi1*(1-n), n=(t-t1)/(t2-t1); a time that
 * sits on a snapshot returns that snapshot. */
/* Immersed cells of an ImmersedBoundaryGrid (GridFittedBottom-style: whole cells), as a parent array of 0 / 1 flags
 * evaluated by the host from the bottom height (immersed_cell: z_centre <= bottom_height). */
void lfo_set_immersed(lfo_state *S, const double *imm_parent)
{
    memcpy(S->imm.p, imm_parent, sizeof(double) * S->imm.len);
    S->P.immersed = 1;
}

/* PartialCellBottom (examples/offline_filter_lee_wave.jl:53): cell heights dz(i,j,k) as a parent array.  Oceananigans'
 * partial-cell metrics: dz at an x (y) face = min of the two adjacent cell heights, so Ax = dy min(dz[i-1], dz[i]),
 * Ay = dx min(dz[j-1], dz[j]), Az = dx dy, V = dx dy dz[i,j,k]. */
void lfo_set_cell_heights(lfo_state *S, const double *dz_parent)
{
    memcpy(S->dzc.p, dz_parent, sizeof(double) * S->dzc.len);
    S->partial = 1;
}

/* mask_immersed_field!: tracers are zero inside the immersed boundary (update_lagrangian_filter_state.jl:22-24) */
static void mask_immersed(lfo_state *S)
{
    if (!S->P.immersed) return;
    for (int t = 0; t < S->n_T; ++t)
        for (long q = 0; q < S->U[t].len; ++q) if (S->imm.p[q] != 0.0) S->U[t].p[q] = 0.0;
}

static void time_interp(double *dst, long len, double *const *series, const double *ts, int ns, double t)
{
    int n1 = 0;
    while (n1 + 1 < ns && ts[n1 + 1] <= t) ++n1;           /* last snapshot with time <= t */
    if (n1 >= ns - 1 || ts[n1] == t || t <= ts[0]) {
        if (t <= ts[0]) n1 = 0;
        memcpy(dst, series[n1], sizeof(double) * len);
        return;
    }
    double nn = (t - ts[n1]) / (ts[n1 + 1] - ts[n1]);
    const double *p1 = series[n1], *p2 = series[n1 + 1];
    double om = 1 - nn;
    #pragma omp parallel for schedule(static)
    for (long q = 0; q < len; ++q) dst[q] = p2[q] * nn + p1[q] * om;
}

static void update_input_data(lfo_state *S, double t)
{
    for (int ax = 0; ax < 3; ++ax)
        if (S->P.vel_present[ax]) time_interp(S->vel[ax].p, S->vel[ax].len, S->snap_vel[ax], S->snap_t, S->n_snap, t);
    for (int v = 0; v < S->P.n_vars; ++v)
        time_interp(S->var[v].p, S->var[v].len, S->snap_var[v], S->snap_t, S->n_snap, t);
}

/* ---------------------------------------------------------------- numerics */

/* WENO5-Z, left-biased argument order (p,o,n | e,f); reference: Oceananigans WENO(order=5)
 * reached through div_Uc, lagrangian_filter_tendency_kernel_functions.jl:56.  SURVEY.md App. A.5. */
static inline double weno5z(double p, double o, double n, double e, double f)
{
    const double eps = 1e-8;
    double b0 = n * (10 * n - 31 * e + 11 * f) + e * (25 * e - 19 * f) + f * (4 * f);
    double b1 = o * (4 * o - 13 * n + 5 * e) + n * (13 * n - 13 * e) + e * (4 * e);
    double b2 = p * (4 * p - 19 * o + 11 * n) + o * (25 * o - 31 * n) + n * (10 * n);
    double q0 = n / 3 + 5 * e / 6 - f / 6;
    double q1 = -o / 6 + 5 * n / 6 + e / 3;
    double q2 = p / 3 - 7 * o / 6 + 11 * n / 6;
    double tau = fabs(b0 - b2);
    double r0 = tau / (b0 + eps), r1 = tau / (b1 + eps), r2 = tau / (b2 + eps);
    double a0 = 0.3 * (1 + r0 * r0), a1 = 0.6 * (1 + r1 * r1), a2 = 0.1 * (1 + r2 * r2);
    return (a0 * q0 + a1 * q1 + a2 * q2) / (a0 + a1 + a2);
}

/* WENO3-Z buffer scheme, left-biased order (o,n | e). */
static inline double weno3z(double o, double n, double e)
{
    const double eps = 1e-8;
    double b0 = (e - n) * (e - n), b1 = (n - o) * (n - o);
    double q0 = (n + e) / 2, q1 = -o / 2 + 3 * n / 2;
    double tau = fabs(b0 - b1);
    double r0 = tau / (b0 + eps), r1 = tau / (b1 + eps);
    double a0 = (2.0 / 3.0) * (1 + r0 * r0), a1 = (1.0 / 3.0) * (1 + r1 * r1);
    return (a0 * q0 + a1 * q1) / (a0 + a1);
}

/* WENO5-Z with the nonlinear weights evaluated in Float32 (EXPERIMENT, lfo_problem.weights_f32): the
 * candidates and the final combination stay FP64.  Used only to measure what Float32 weights would cost
 * in parity; never a product path. */
static inline double weno5z_w32(double p, double o, double n, double e, double f)
{
    const float eps = 1e-8f;
    float pf = (float)p, of = (float)o, nf = (float)n, ef = (float)e, ff = (float)f;
    float b0 = nf * (10 * nf - 31 * ef + 11 * ff) + ef * (25 * ef - 19 * ff) + ff * (4 * ff);
    float b1 = of * (4 * of - 13 * nf + 5 * ef) + nf * (13 * nf - 13 * ef) + ef * (4 * ef);
    float b2 = pf * (4 * pf - 19 * of + 11 * nf) + of * (25 * of - 31 * nf) + nf * (10 * nf);
    float tau = fabsf(b0 - b2);
    float r0 = tau / (b0 + eps), r1 = tau / (b1 + eps), r2 = tau / (b2 + eps);
    double a0 = 0.3f * (1 + r0 * r0), a1 = 0.6f * (1 + r1 * r1), a2 = 0.1f * (1 + r2 * r2);
    double q0 = n / 3 + 5 * e / 6 - f / 6;
    double q1 = -o / 6 + 5 * n / 6 + e / 3;
    double q2 = p / 3 - 7 * o / 6 + 11 * n / 6;
    return (a0 * q0 + a1 * q1 + a2 * q2) / (a0 + a1 + a2);
}

/* Reconstruction kinds of a scheme chain.  Every Oceananigans scheme carries a `buffer_scheme` that takes over
 * where its own stencil would reach outside the domain or into the immersed boundary; the chains (highest
 * order first) follow the constructors' defaults, e.g. the one printed at OfflineLagrangianFilter.jl:212-214:
 *   WENO(order=5) -> WENO(order=3) -> Centered(order=2)
 *   UpwindBiased(order=5) -> UpwindBiased(order=3) -> UpwindBiased(order=1)
 *   Centered(order=4) -> Centered(order=2)
 * Only the WENO(order=5) chain is pinned by the reference's golden file. */
enum { K_C2, K_C4, K_U1, K_U3, K_U5, K_W3, K_W5 };

static int scheme_chain(int scheme, int kinds[3], int bufs[3])
{
    switch (scheme) {
    case ADV_WENO5:     kinds[0] = K_W5; bufs[0] = 3; kinds[1] = K_W3; bufs[1] = 2; kinds[2] = K_C2; bufs[2] = 1; return 3;
    case ADV_WENO3:     kinds[0] = K_W3; bufs[0] = 2; kinds[1] = K_C2; bufs[1] = 1; return 2;
    case ADV_CENTERED4: kinds[0] = K_C4; bufs[0] = 2; kinds[1] = K_C2; bufs[1] = 1; return 2;
    case ADV_CENTERED2: kinds[0] = K_C2; bufs[0] = 1; return 1;
    case ADV_UPWIND5:   kinds[0] = K_U5; bufs[0] = 3; kinds[1] = K_U3; bufs[1] = 2; kinds[2] = K_U1; bufs[2] = 1; return 3;
    case ADV_UPWIND3:   kinds[0] = K_U3; bufs[0] = 2; kinds[1] = K_U1; bufs[1] = 1; return 2;
    case ADV_UPWIND1:   kinds[0] = K_U1; bufs[0] = 1; return 1;
    }
    return 0;
}

/* Face value at face m (1-based: between cells m-1 and m) along one axis.
 * cm points at cell m, s is the stride along the axis, N the cell count, q the advecting face velocity, imm
 * (or NULL) points at the immersed flag of cell m.  The first scheme of the chain whose symmetric stencil of
 * `buffer` cells on each side of the face (cells m-b .. m+b-1) lies inside a Bounded domain and outside the
 * immersed boundary is used (Oceananigans near_*_boundary / near_*_immersed_boundary; for WENO(order=5) this is
 * the face-index rule of SURVEY.md App. A.5).  Where not even the two adjacent cells are active the face is a
 * wall (velocity zero / flux masked) and the centred average is returned. */
static inline double face_value(const lfo_problem *P, const double *cm, const double *imm, long s, int m, int N, int topo, double q)
{
    int kinds[3], bufs[3];
    const int nk = scheme_chain(P->advect, kinds, bufs);
    int kind = K_C2;
    for (int c = 0; c < nk; ++c) {
        const int b = bufs[c];
        int ok = 1;
        if (topo == TOPO_BOUNDED && (m - b < 1 || m + b - 1 > N)) ok = 0;
        if (ok && imm) for (int r = -b; r < b; ++r) if (imm[r * s] != 0.0) { ok = 0; break; }
        if (ok) { kind = kinds[c]; break; }
    }
    const int left = q > 0;
    /* upwind-ordered stencil (p, o, n | e, f): n is the upwind cell */
    #define CELL(r) (left ? cm[((r) - 1) * s] : cm[-(r) * s])      /* r = -2..2 relative to n */
    switch (kind) {
    case K_W5: return P->weights_f32 ? weno5z_w32(CELL(-2), CELL(-1), CELL(0), CELL(1), CELL(2))
                                     : weno5z(CELL(-2), CELL(-1), CELL(0), CELL(1), CELL(2));
    case K_W3: return weno3z(CELL(-1), CELL(0), CELL(1));
    case K_U1: return CELL(0);
    case K_U3: return (-1.0 / 6.0) * CELL(-1) + (5.0 / 6.0) * CELL(0) + (1.0 / 3.0) * CELL(1);
    case K_U5: return (2.0 / 60.0) * CELL(-2) + (-13.0 / 60.0) * CELL(-1) + (47.0 / 60.0) * CELL(0)
                      + (27.0 / 60.0) * CELL(1) + (-3.0 / 60.0) * CELL(2);
    case K_C4: return (-1.0 / 12.0) * cm[-2 * s] + (7.0 / 12.0) * cm[-s] + (7.0 / 12.0) * cm[0] + (-1.0 / 12.0) * cm[s];
    default:   return (cm[-s] + cm[0]) / 2;
    }
    #undef CELL
}

/* a9 + a3 (+ a4): compute_tendencies! -> compute_Gc! -> tracer_tendency
 * (compute_lagrangian_filter_tendencies.jl:16-91, lagrangian_filter_tendency_kernel_functions.jl:36-61,
 * lagrangian_filtering_advection_operators.jl:19-21) with the forcing closures of
 * utils:487-565 (create_forcing utils:731-865) and relaxation utils:584-699. */
static void compute_tendencies(lfo_state *S)
{
    const lfo_problem *P = &S->P;
    const int Nx = P->N[0], Ny = P->N[1], Nz = P->N[2];
    const double dx = P->delta[0], dy = P->delta[1], dz = P->delta[2];
    const double Ax = dy * dz, Ay = dx * dz, Az = dx * dy, V = dx * dy * dz;
    const int np = P->n_pairs;

    for (int t = 0; t < S->n_T; ++t) {
        const int is_S = t >= S->n_C;
        const int tc = is_S ? t - S->n_C : t;       /* index inside the C (or S) block */
        const int fld = tc / np, pole = tc % np;
        const int is_map = fld >= P->n_vars;
        const int map_ax = is_map ? S->map_axis[fld - P->n_vars] : -1;
        const double c = P->c[pole], d = P->single_exp ? 0.0 : P->d[pole];
        const lfo_field *Uc = &S->U[fld * np + pole];                              /* cosine partner */
        const lfo_field *Us = P->single_exp ? NULL : &S->U[S->n_C + fld * np + pole]; /* sine partner   */
        const lfo_field *me = &S->U[t];
        lfo_field *G = &S->G[t];
        const long sc[3] = {1, me->s[0], (long)me->s[0] * me->s[1]};

        #pragma omp parallel for collapse(2) schedule(static)
        for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
        for (int i = 0; i < Nx; ++i) {
            const int ijk[3] = {i, j, k};
            long ic = fidx(me, P, i, j, k);
            const double *cp = me->p + ic;
            double div_flux = 0, div_u = 0;
            if (P->advect) {
                double a_lo[3] = {Ax, Ay, Az}, a_hi[3] = {Ax, Ay, Az}, Vc = V;
                if (S->partial) {
                    const double *hz = S->dzc.p + ic;
                    if (P->topo[0] != TOPO_FLAT) { a_lo[0] = dy * fmin(hz[-sc[0]], hz[0]); a_hi[0] = dy * fmin(hz[0], hz[sc[0]]); }
                    if (P->topo[1] != TOPO_FLAT) { a_lo[1] = dx * fmin(hz[-sc[1]], hz[0]); a_hi[1] = dx * fmin(hz[0], hz[sc[1]]); }
                    Vc = Az * hz[0];
                }
                for (int ax = 0; ax < 3; ++ax) {
                    if (P->topo[ax] == TOPO_FLAT || !P->vel_present[ax]) continue;
                    const lfo_field *vf = &S->vel[ax];
                    const long sv[3] = {1, vf->s[0], (long)vf->s[0] * vf->s[1]};
                    long iv = fidx(vf, P, i, j, k);
                    double q_lo = vf->p[iv], q_hi = vf->p[iv + sv[ax]];
                    int m = ijk[ax] + 1;  /* 1-based face index of the low face */
                    const double *ip = P->immersed ? S->imm.p + ic : NULL;
                    double c_lo = face_value(P, cp, ip, sc[ax], m, P->N[ax], P->topo[ax], q_lo);
                    double c_hi = face_value(P, cp + sc[ax], ip ? ip + sc[ax] : NULL, sc[ax], m + 1, P->N[ax], P->topo[ax], q_hi);
                    double F_lo = a_lo[ax] * q_lo * c_lo, F_hi = a_hi[ax] * q_hi * c_hi;
                    /* conditional_flux: no advective flux through a face of the immersed boundary */
                    if (ip && (ip[-sc[ax]] != 0.0 || ip[0] != 0.0)) F_lo = 0;
                    if (ip && (ip[0] != 0.0 || ip[sc[ax]] != 0.0)) F_hi = 0;
                    div_flux += (F_hi - F_lo);
                    div_u += (a_hi[ax] * q_hi - a_lo[ax] * q_lo);
                }
                div_flux = div_flux / Vc;
                div_u = div_u / Vc;
            }
            /* forcing */
            double forcing;
            double gC = Uc->p[ic], gS = Us ? Us->p[ic] : 0.0;
            double src = 0;   /* original variable (vars) or centred velocity (maps) */
            if (!is_map) src = S->var[fld].p[ic];
            else if (P->vel_present[map_ax]) {
                const lfo_field *vf = &S->vel[map_ax];
                const long sv[3] = {1, vf->s[0], (long)vf->s[0] * vf->s[1]};
                long iv = fidx(vf, P, i, j, k);
                src = (P->topo[map_ax] == TOPO_FLAT) ? vf->p[iv] : (vf->p[iv] + vf->p[iv + sv[map_ax]]) / 2;
            }
            const double n2 = c * c + d * d;
            if (P->single_exp) {
                if (!is_map) forcing = (-c * gC) + src;                                   /* utils:491,743 */
                else         forcing = (-c / n2 * src) + (-c * gC);                       /* utils:543 + 491 */
            } else if (!is_S) {
                if (!is_map) forcing = (-c * gC - d * gS) + src;                          /* utils:498,743 */
                else         forcing = (-c / n2 * src) + (-c * gC - d * gS);              /* utils:543 + 498 */
            } else {
                if (!is_map) forcing = (-c * gS + d * gC);                                /* utils:522 */
                else         forcing = (-d / n2 * src) + (-c * gS + d * gC);              /* utils:563 + 522 */
            }
            if (P->relax) {
                double g = cp[0], geq;
                if (P->single_exp)      geq = is_map ? (-1 / (c * c)) * src : src / c;     /* utils:592,659 */
                else if (!is_S)         geq = is_map ? src * (d * d - c * c) / (n2 * n2) : src * c / n2;  /* utils:602,669 */
                else                    geq = is_map ? src * (-2 * c * d) / (n2 * n2) : src * d / n2;     /* utils:633,697 */
                forcing += -1 / P->relax_timescale * (g - geq) * S->mask.p[ic];
            }
            G->p[ic] = (-div_flux + cp[0] * div_u) + forcing;
        }
    }
}

/* a8: update_state!, update_lagrangian_filter_state.jl:19-53 */
void lfo_update_state(lfo_state *S)
{
    mask_immersed(S);
    for (int t = 0; t < S->n_T; ++t) fill_halos(&S->U[t], &S->P);
    update_input_data(S, S->time);
    compute_tendencies(S);
}

/* a6: initialise_filtered_vars_from_data, utils:1086-1137 (whole parent arrays, halos too;
 * maps use Field(@at (Center,Center,Center) vel[Time(0)]) = interior average + halo fill). */
void lfo_init_from_data(lfo_state *S)
{
    const lfo_problem *P = &S->P;
    const int np = P->n_pairs;
    update_input_data(S, 0.0);
    lfo_field tmp; field_alloc(&tmp, P, -1);
    for (int fld = 0; fld < S->n_fields; ++fld) {
        const int is_map = fld >= P->n_vars;
        const double *srcp;
        if (!is_map) srcp = S->var[fld].p;
        else {
            int ax = S->map_axis[fld - P->n_vars];
            const lfo_field *vf = &S->vel[ax];
            const long sv[3] = {1, vf->s[0], (long)vf->s[0] * vf->s[1]};
            memset(tmp.p, 0, sizeof(double) * tmp.len);
            for (int k = 0; k < P->N[2]; ++k) for (int j = 0; j < P->N[1]; ++j) for (int i = 0; i < P->N[0]; ++i) {
                long iv = fidx(vf, P, i, j, k);
                tmp.p[fidx(&tmp, P, i, j, k)] = (P->topo[ax] == TOPO_FLAT) ? vf->p[iv] : (vf->p[iv] + vf->p[iv + sv[ax]]) / 2;
            }
            fill_halos(&tmp, P);
            srcp = tmp.p;
        }
        for (int pole = 0; pole < np; ++pole) {
            double c = P->c[pole], d = P->single_exp ? 0.0 : P->d[pole];
            double fC, fS = 0;
            if (P->single_exp) fC = is_map ? (-1 / (c * c)) : 1 / c;
            else if (!is_map) { fC = c / (c * c + d * d); fS = d / (c * c + d * d); }
            else { double n2 = c * c + d * d; fC = (d * d - c * c) / (n2 * n2); fS = -2 * c * d / (n2 * n2); }
            lfo_field *UC = &S->U[fld * np + pole];
            for (long q = 0; q < UC->len; ++q) UC->p[q] = fC * srcp[q];
            if (!P->single_exp) {
                lfo_field *US = &S->U[S->n_C + fld * np + pole];
                for (long q = 0; q < US->len; ++q) US->p[q] = fS * srcp[q];
            }
        }
    }
    free(tmp.p);
    S->time = 0; S->iteration = 0;
}

/* a14: change_sign_of_map_variables!, utils:1232-1251 (parent arrays) */
void lfo_negate_maps(lfo_state *S)
{
    const int np = S->P.n_pairs;
    for (int fld = S->P.n_vars; fld < S->n_fields; ++fld)
        for (int pole = 0; pole < np; ++pole) {
            lfo_field *f = &S->U[fld * np + pole];
            for (long q = 0; q < f->len; ++q) f->p[q] = -f->p[q];
            if (!S->P.single_exp) {
                f = &S->U[S->n_C + fld * np + pole];
                for (long q = 0; q < f->len; ++q) f->p[q] = -f->p[q];
            }
        }
}

/* a10 + a11: rk3_substep! (lagrangian_filter_rk3_substep.jl:31-61, U += dt(g G + z G-)) and
 * cache_previous_tendencies! (cache_lagrangian_filter_tendencies.jl:21-31), interior only. */
static void rk3_substep(lfo_state *S, double dt, double gamma, double zeta, int first)
{
    const lfo_problem *P = &S->P;
    for (int t = 0; t < S->n_T; ++t) {
        lfo_field *U = &S->U[t], *G = &S->G[t], *Gm = &S->Gm[t];
        #pragma omp parallel for collapse(2) schedule(static)
        for (int k = 0; k < P->N[2]; ++k) for (int j = 0; j < P->N[1]; ++j) for (int i = 0; i < P->N[0]; ++i) {
            long q = fidx(U, P, i, j, k);
            if (first) U->p[q] += dt * gamma * G->p[q];      /* Oceananigans: convert(FT, dt) * gamma1 * G1 */
            else       U->p[q] += dt * (gamma * G->p[q] + zeta * Gm->p[q]);
        }
    }
}

static void cache_tendencies(lfo_state *S)
{
    for (int t = 0; t < S->n_T; ++t) memcpy(S->Gm[t].p, S->G[t].p, sizeof(double) * S->G[t].len);
}

/* One RK3 step of size dt (Oceananigans time_step!(::RungeKutta3) driving the LagrangianFilter
 * method extensions).  Requires lfo_update_state() to have been called once
 * at iteration 0 (as Oceananigans does). */
/* QuasiAdamsBashforth2 step (lagrangian_filter_ab2_step.jl:24-51 + Oceananigans time_step!/_ab2_step_field!):
 * U += dt ((3/2 + chi) G - (1/2 + chi) G^-), an Euler step (chi = -1/2) at iteration 0 and whenever dt differs
 * from the previous step's. */
static void ab2_time_step(lfo_state *S, double dt)
{
    const lfo_problem *P = &S->P;
    if (S->iteration == 0) lfo_update_state(S);
    const int euler = dt != S->last_dt;
    const double chi = euler ? -0.5 : P->chi;
    for (int t = 0; t < S->n_T; ++t) {
        lfo_field *U = &S->U[t], *G = &S->G[t], *Gm = &S->Gm[t];
        for (int k = 0; k < P->N[2]; ++k) for (int j = 0; j < P->N[1]; ++j) for (int i = 0; i < P->N[0]; ++i) {
            long q = fidx(U, P, i, j, k);
            double Gu = (1.5 + chi) * G->p[q] - (euler ? 0.0 : (0.5 + chi) * Gm->p[q]);
            U->p[q] += dt * Gu;
        }
    }
    S->time = S->time + dt;
    S->last_dt = dt;
    for (int t = 0; t < S->n_T; ++t) memcpy(S->Gm[t].p, S->G[t].p, sizeof(double) * S->G[t].len);
    lfo_update_state(S);
    S->iteration += 1;
}

void lfo_time_step(lfo_state *S, double dt)
{
    if (S->P.timestepper == 1) { ab2_time_step(S, dt); return; }
    const double g1 = 8.0 / 15, g2 = 5.0 / 12, g3 = 3.0 / 4, z2 = -17.0 / 60, z3 = -5.0 / 12;
    double t0 = S->time, t1 = t0 + dt;
    if (S->iteration == 0) lfo_update_state(S);
    rk3_substep(S, dt, g1, 0, 1);
    S->time = t0 + g1 * dt;
    cache_tendencies(S);
    lfo_update_state(S);
    rk3_substep(S, dt, g2, z2, 0);
    S->time = S->time + (g2 + z2) * dt;
    cache_tendencies(S);
    lfo_update_state(S);
    rk3_substep(S, dt, g3, z3, 0);
    S->time = t1;
    cache_tendencies(S);
    lfo_update_state(S);
    S->iteration += 1;
}

/* tracer access (parent arrays) */
void lfo_get_tracer(const lfo_state *S, int t, double *out) { memcpy(out, S->U[t].p, sizeof(double) * S->U[t].len); }
void lfo_set_tracer(lfo_state *S, int t, const double *in) { memcpy(S->U[t].p, in, sizeof(double) * S->U[t].len); }
void lfo_get_tendency(const lfo_state *S, int t, double *out) { memcpy(out, S->G[t].p, sizeof(double) * S->G[t].len); }

/* a13: create_output_fields, utils:898-1012.  kind 0: <var>_Lagrangian_filtered (field = var index),
 * kind 1: xi_<vel> (field = map index), kind 2: <vel>_Lagrangian_filtered (mean velocity).
 * Computed on the interior, then halo-filled like any Oceananigans computed Field. */
void lfo_output(lfo_state *S, int kind, int field, double *out)
{
    const lfo_problem *P = &S->P;
    const int np = P->n_pairs;
    const int fld = kind == 0 ? field : P->n_vars + field;
    lfo_field o; field_alloc(&o, P, -1);
    for (int t = 0; t < S->n_T; ++t) fill_halos(&S->U[t], P);
    for (int k = 0; k < P->N[2]; ++k) for (int j = 0; j < P->N[1]; ++j) for (int i = 0; i < P->N[0]; ++i) {
        long q = fidx(&o, P, i, j, k);
        double tot = 0;
        for (int pole = 0; pole < np; ++pole) {
            double a = P->a[pole], b = P->b[pole], c = P->c[pole], d = P->d[pole];
            double gC = S->U[fld * np + pole].p[q];
            double term;
            if (P->single_exp) term = kind == 2 ? -a * c * gC : a * gC;
            else {
                double gS = S->U[S->n_C + fld * np + pole].p[q];
                term = kind == 2 ? (-a * c + b * d) * gC + (-a * d - b * c) * gS : a * gC + b * gS;
            }
            tot = pole == 0 ? term : tot + term;
        }
        o.p[q] = tot;
    }
    fill_halos(&o, P);
    memcpy(out, o.p, sizeof(double) * o.len);
    free(o.p);
}
