/*
 * lfb200.h -- C ABI of liblfb200.so: the B200-native (sm_100a) hot path of
 * OceananigansLagrangianFilter.jl's offline filter PDEs.
 *
 * This is the drop-in boundary.  It replaces what the reference does between
 * src/OfflineLagrangianFilter/run_offline_lagrangian_filter.jl:47 and :105, i.e. the
 * `LagrangianFilter` model (lagrangian_filter.jl:23-180), the five method extensions Oceananigans'
 * time stepper calls on it (rk3_substep! lagrangian_filter_rk3_substep.jl:11, cache_previous_tendencies!
 * cache_lagrangian_filter_tendencies.jl:21, update_state! update_lagrangian_filter_state.jl:19,
 * compute_tendencies! compute_lagrangian_filter_tendencies.jl:16, compute_flux_bc_tendencies! :99),
 * the update_input_data! callback (src/Utils/lagrangian_filter_utils.jl:1036-1058) and the helpers
 * that initialise / flip / combine the state (utils:1086-1137, 1232-1251, 898-1012;
 * src/Utils/post_processing_utils.jl:45-170).  Every entry point cites the reference interface it
 * stands in for.  The Julia-side `ccall` stubs are in INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no exceptions cross the boundary.  Every function returns
 *    LFB_OK (0) or a negative error code; lfb_last_error() returns the message of the last failure on
 *    the calling thread.
 *  - Host arrays are exactly Oceananigans' `parent(field)`: FP64 (or FP32 for the *_f32 variants),
 *    column-major, x fastest, halos included; a field located on a Face of a Bounded axis carries one
 *    extra entry on that axis (w: Nz+1 levels).  They are borrowed only for the duration of the call.
 *  - The library owns all device memory.  One handle is used from one host thread at a time;
 *    distinct handles are independent.  With several GPUs every rank (one process per GPU) owns one
 *    handle holding its y-slab; see lfb_comm_init.
 *  - There is no CPU fallback: creation fails with LFB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef LFB200_H
#define LFB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFB_VERSION 110           /* 0.1.1 */

#define LFB_MAX_COEFFS 8          /* coefficient sets (a,b,c,d), i.e. Butterworth order up to 16 */
#define LFB_MAX_VARS   8          /* variables to filter                                         */

/* error codes */
#define LFB_OK              0
#define LFB_ERR_ARG        -1     /* invalid argument / configuration                            */
#define LFB_ERR_CUDA       -2     /* CUDA runtime failure (no device, launch or memory error)    */
#define LFB_ERR_STATE      -3     /* call sequence error (e.g. stepping without resident data)   */
#define LFB_ERR_NO_DATA    -4     /* no resident snapshot brackets the requested time            */
#define LFB_ERR_COMM       -5     /* NCCL failure                                                */

/* grid topology per axis (Oceananigans Periodic / Bounded / Flat) */
#define LFB_PERIODIC 0
#define LFB_BOUNDED  1
#define LFB_FLAT     2

/* advection (OfflineFilterConfig.advection, OfflineLagrangianFilter.jl:139,246) */
#define LFB_ADVECTION_NONE      0 /* advection = nothing: Eulerian filter through the same path  */
#define LFB_ADVECTION_WENO5     1 /* WENO() default: WENO5-Z, buffers WENO3-Z and Centered2      */
/* the other schemes OfflineFilterConfig.advection admits, each with Oceananigans' buffer-scheme chain
 * (the scheme whose stencil fits between the walls / outside the immersed boundary is used) */
#define LFB_ADVECTION_CENTERED2 2 /* Centered(order=2)                                           */
#define LFB_ADVECTION_CENTERED4 3 /* Centered(order=4) -> Centered2                              */
#define LFB_ADVECTION_UPWIND1   4 /* UpwindBiased(order=1)                                       */
#define LFB_ADVECTION_UPWIND3   5 /* UpwindBiased(order=3) -> UpwindBiased1                      */
#define LFB_ADVECTION_UPWIND5   6 /* UpwindBiased(order=5) -> UpwindBiased3 -> UpwindBiased1     */
#define LFB_ADVECTION_WENO3     7 /* WENO(order=3) -> Centered2                                  */

/* time stepper (LagrangianFilter(...; timestepper), lagrangian_filter.jl:48,85) */
#define LFB_TIMESTEPPER_RK3 0     /* :RungeKutta3 (default), lagrangian_filter_rk3_substep.jl:11-61 */
#define LFB_TIMESTEPPER_AB2 1     /* :QuasiAdamsBashforth2, lagrangian_filter_ab2_step.jl:24-51  */

/* kernel selection (lfb_desc.kernel) */
#define LFB_KERNEL_AUTO    0
#define LFB_KERNEL_GENERIC 1      /* one thread per cell and coefficient pair, any topology      */
                                  /* ids 2, 3: the retired cp.async z-march kernels (rejected)   */
#define LFB_KERNEL_ZTMA    4      /* z-march with TMA-staged tiles; what LFB_KERNEL_AUTO selects on
                                   * 3-D grids (x, y, z Periodic or Bounded), u,v,w present,
                                   * WENO advection (boundary relaxation included), strict = 0       */

/* field ids for snapshots */
#define LFB_FIELD_U 0
#define LFB_FIELD_V 1
#define LFB_FIELD_W 2
#define LFB_FIELD_VAR0 3          /* LFB_FIELD_VAR0 + q = q-th entry of var_names_to_filter      */

/* output kinds of lfb_output (create_output_fields, utils:898-1012) */
#define LFB_OUT_FILTERED_VAR 0    /* <var>_Lagrangian_filtered (or _Eulerian_filtered)           */
#define LFB_OUT_XI           1    /* xi_<vel>                                                    */
#define LFB_OUT_MEAN_VEL     2    /* <vel>_Lagrangian_filtered                                   */

typedef struct lfb_handle lfb_handle;

/* Everything the kernels need from OfflineFilterConfig (OfflineLagrangianFilter.jl:83-113) and the
 * grid; validation of the user-facing options stays on the host side (:255-485). */
typedef struct lfb_desc {
    int32_t size[3];              /* Nx, Ny, Nz of THIS rank's slab                              */
    int32_t halo[3];              /* Hx, Hy, Hz (0 on Flat axes; 3 for the default WENO)         */
    int32_t topology[3];          /* LFB_PERIODIC / LFB_BOUNDED / LFB_FLAT of the GLOBAL grid    */
    double  spacing[3];           /* regular spacings; 1.0 on Flat axes                          */
    int32_t n_vars;               /* length(var_names_to_filter)                                 */
    int32_t velocity_mask;        /* bit0 u, bit1 v, bit2 w present in velocity_names            */
    int32_t n_coeffs;             /* filter_params.N_coeffs; 0 encodes the single exponential    */
    double  a[LFB_MAX_COEFFS], b[LFB_MAX_COEFFS], c[LFB_MAX_COEFFS], d[LFB_MAX_COEFFS];
    int32_t with_maps;            /* map_to_mean || compute_mean_velocities (utils:439,459)      */
    int32_t advection;            /* LFB_ADVECTION_*                                             */
    int32_t n_slots;              /* resident input snapshots (>= 2); InMemory(4) analogue       */
    int32_t device;               /* CUDA device ordinal                                         */
    int32_t strict;               /* 1: reference order of operations, no FMA contraction        */
    int32_t kernel;               /* LFB_KERNEL_*                                                */
    int32_t relaxation;           /* boundary_relaxation (OfflineLagrangianFilter.jl:108-111)    */
    double  relax_timescale;
    int32_t timestepper;          /* LFB_TIMESTEPPER_*                                           */
    int32_t immersed;             /* 1: ImmersedBoundaryGrid; the cell flags follow with lfb_set_immersed */
    double  chi;                  /* AB2 parameter (Oceananigans default 0.1); ignored by RK3   */
    int32_t reserved[4];          /* must be zero                                                */
} lfb_desc;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int         lfb_version(void);
const char *lfb_last_error(void);
/* replaces LagrangianFilter(grid; tracers, auxiliary_fields, forcing, advection), lagrangian_filter.jl:78-180 */
int         lfb_create(const lfb_desc *desc, lfb_handle **out);
int         lfb_destroy(lfb_handle *h);
int         lfb_sync(lfb_handle *h);                      /* wait for all queued device work     */

/* ---- tracer bookkeeping (create_filtered_vars, utils:423-470) -------------------------------- */
int lfb_n_tracers(const lfb_handle *h);
/* field: 0..n_vars-1 variables, then one map per present velocity in u,v,w order; pole: 0-based
 * coefficient index; sine: 0 -> _C tracer, 1 -> _S tracer.  Returns the tracer index or <0. */
int lfb_tracer_index(const lfb_handle *h, int field, int pole, int sine);
/* number of FP64 elements of a centre-located / velocity parent array of this handle */
int64_t lfb_parent_len(const lfb_handle *h, int field_id);

/* ---- input series (load_data utils:201-215; update_input_data! utils:1036-1058;
 *      create_input_data_on_disk utils:97-180) ------------------------------------------------- */
/* Make `parent` the content of snapshot slot `slot` for `field_id`, tagged with the pass time
 * `time` (t - T_start on the forward pass).  All fields of one slot must carry the same time.
 * The copy runs on a side stream and overlaps with queued stage kernels that do not read `slot`. */
int lfb_upload_snapshot(lfb_handle *h, int slot, int field_id, const double *parent, double time);
int lfb_upload_snapshot_f32(lfb_handle *h, int slot, int field_id, const float *parent, double time);
/* Asynchronous variant: returns as soon as the copy is queued.  `parent` must stay valid until
 * lfb_wait_transfers(); the copy overlaps with queued stage kernels only if `parent` is page-locked
 * (lfb_host_alloc or cudaHostRegister). */
int lfb_upload_snapshot_async(lfb_handle *h, int slot, int field_id, const void *parent, int is_f32, double time);
int lfb_invalidate_slot(lfb_handle *h, int slot);
/* Re-tag a resident slot with a new pass time without touching its data. */
int lfb_set_slot_time(lfb_handle *h, int slot, double time);
/* direction = +1: forward pass.  direction = -1: backward pass: velocities are negated on the fly
 * (utils:169-173).  Together with lfb_set_slot_time(slot, T_end - t) this replaces the reference's
 * rewrite of the whole input file (utils:154-173): nothing is copied or re-uploaded. */
int lfb_set_direction(lfb_handle *h, int direction);
/* relaxation mask evaluated at cell centres (mask_func, utils:584-699); parent layout */
int lfb_set_mask(lfb_handle *h, const double *parent);
/* Immersed cells of an ImmersedBoundaryGrid (lee-wave example, examples/offline_filter_lee_wave.jl:49-53):
 * parent array of 0 / 1 flags at cell centres, halos included (periodic halos filled like the data), evaluated
 * by the host from the bottom height.  Tracers are kept at zero inside the boundary (mask_immersed_field!,
 * update_lagrangian_filter_state.jl:22-24), no advective flux crosses a face that touches an immersed cell and
 * the reconstruction drops to the buffer scheme whose stencil stays outside the boundary (Oceananigans'
 * conditional fluxes / near-boundary interpolation reached through div_Uc,
 * lagrangian_filter_tendency_kernel_functions.jl:56).  Needs desc.immersed = 1.
 * cell_heights (or NULL: whole cells, GridFittedBottom): parent array of the cell heights dz(i,j,k) of a
 * PartialCellBottom (the lee-wave example's choice); they enter the flux divergence through Oceananigans'
 * partial-cell metrics Ax = dy min(dz[i-1], dz[i]), Ay = dx min(dz[j-1], dz[j]), Az = dx dy, V = dx dy dz.
 * Whole-cell immersed grids run on the TMA z-march kernel (face codes as one more TMA tile); with cell_heights the
 * generic kernel is used (LFB_KERNEL_ZTMA is then refused). */
int lfb_set_immersed(lfb_handle *h, const double *parent, const double *cell_heights);

/* ---- state ----------------------------------------------------------------------------------- */
/* initialise_filtered_vars_from_data, utils:1086-1137 (uses the snapshot at pass time 0) */
int lfb_init_from_data(lfb_handle *h);
/* change_sign_of_map_variables!, utils:1232-1251 */
int lfb_negate_maps(lfb_handle *h);
/* reset!(model.clock), run_offline_lagrangian_filter.jl:93 */
int lfb_reset_clock(lfb_handle *h);
double  lfb_time(const lfb_handle *h);
int64_t lfb_iteration(const lfb_handle *h);
/* parent(model.tracers[tracer]) with halos filled as update_state! leaves them */
int lfb_get_tracer(lfb_handle *h, int tracer, double *parent);
int lfb_set_tracer(lfb_handle *h, int tracer, const double *parent);
/* tendency G^n of the CURRENT state at the current time (compute_tendencies!); diagnostic */
int lfb_get_tendency(lfb_handle *h, int tracer, double *parent);

/* ---- time stepping --------------------------------------------------------------------------- */
/* One RK3 step of size dt = time_step!(model, dt): three fused stage kernels, each doing for every
 * tracer the input time interpolation, WENO flux divergence, c div(U), forcing, the RK3 substep and
 * the tendency caching.  The snapshots bracketing t, t+8dt/15 and t+2dt/3 must be resident. */
int lfb_step(lfb_handle *h, double dt);
/* Simulation time-step alignment (Oceananigans run!): steps with
 * dt_n = min(dt, next multiple of each align interval - t, t_stop - t) until t_stop. */
int lfb_run_until(lfb_handle *h, double t_stop, double dt, int n_align, const double *align_intervals);

/* ---- outputs (create_output_fields utils:898-1012; written with halos like JLD2Writer) ------- */
int lfb_output(lfb_handle *h, int kind, int field, double *parent);
int lfb_output_f32(lfb_handle *h, int kind, int field, float *parent);
/* Asynchronous output: the combination is queued behind the steps issued so far, exported into a ring of
 * device buffers and copied to `parent` on a separate stream, so the host can keep queuing time steps.
 * `parent` (FP64, or FP32 when is_f32 != 0) may be read after lfb_wait_transfers(). */
int lfb_output_async(lfb_handle *h, int kind, int field, void *parent, int is_f32);
/* wait for every asynchronous upload and output issued so far */
int lfb_wait_transfers(lfb_handle *h);
/* page-locked host memory for the asynchronous transfers */
int lfb_host_alloc(size_t bytes, void **ptr);
int lfb_host_free(void *ptr);
/* sum_forward_backward_contributions!, post:45-170, for one output time on host arrays:
 * out = fwd + sign * (w_hi*bwd_hi + (1-w_hi)*bwd_lo), computed on the device in FP64.
 * sign = +1 for filtered vars and xi, -1 for mean velocities (post:151-164). */
int lfb_combine(lfb_handle *h, int64_t n, const double *fwd, const double *bwd_lo, const double *bwd_hi,
                double w_hi, double sign, double *out);

/* ---- device-resident pass outputs -----------------------------------------------------------------
 * The reference writes each pass's outputs to forward_output.jld2 / backward_output.jld2 (JLD2Writer,
 * run_offline_lagrangian_filter.jl:77-80), reads them back, sums them (post:45-170) and reads the sum again for
 * the remap (post:327-762).  Here the outputs of both passes can stay on the device as dense parent arrays
 * ("kept" outputs, addressed by small integer ids), are combined there, gathered across the y-slabs of a multi-GPU
 * run, handed to lfb_remap_to_mean as device pointers and fetched once at the end. */
/* evaluate an output now (like lfb_output) and keep it: FP64, or rounded to FP32 as the reference's writer does */
int lfb_output_keep(lfb_handle *h, int kind, int field, int is_f32, int *id);
/* new FP64 kept array = fwd + sign * (w_hi * bwd_hi + (1 - w_hi) * bwd_lo)  (lfb_combine on kept arrays;
 * bwd_hi is ignored when w_hi == 0) */
int lfb_kept_combine(lfb_handle *h, int fwd, int bwd_lo, int bwd_hi, double w_hi, double sign, int *id);
/* copy a kept array to the host in its own precision (parent layout) */
int lfb_kept_fetch(lfb_handle *h, int id, void *parent);
/* asynchronous form: the copy runs on the download stream behind the work that produces the array; `parent` may be
 * read after lfb_wait_transfers() (page-locked memory for a real overlap); lfb_kept_release waits for it */
int lfb_kept_fetch_async(lfb_handle *h, int id, void *parent);
/* precision, parent extents and raw device pointer of a kept array (any of the outputs may be NULL) */
int lfb_kept_info(lfb_handle *h, int id, int *is_f32, int *extents3, void **device_ptr);
/* released arrays go to a per-handle pool and are reused by later kept outputs of the same size (no cudaFree, hence
 * no device synchronisation, between time steps); lfb_kept_trim returns the pooled memory to the device */
int lfb_kept_release(lfb_handle *h, int id);
int lfb_kept_trim(lfb_handle *h);
/* make sure the pool holds at least `count` blocks of `bytes` bytes ahead of time (before a pass), so that no cudaMalloc
 * happens between queued time steps */
int lfb_kept_reserve(lfb_handle *h, size_t bytes, int count);
/* COLLECTIVE over the communicator of lfb_comm_init: assemble the GLOBAL parent array of a kept output on rank
 * `root` from every rank's y-slab (interior rows from their owners, outer halo rows from the first / last rank).
 * *gid is the new kept id on `root`, -1 elsewhere; with one rank *gid = id. */
int lfb_kept_gather(lfb_handle *h, int id, int root, int *gid);
/* free / total device memory (the host decides whether the pass outputs fit on the device) */
int lfb_mem_info(lfb_handle *h, size_t *free_bytes, size_t *total_bytes);

/* ---- remap to the mean position (regrid_to_mean_position!, post:327-762) ------------------------ */
typedef struct lfb_remap_desc {
    int32_t size[3], halo[3], topology[3];   /* grid: two or three axes must be non-Flat                 */
    double  origin[3], spacing[3];           /* first face coordinate and spacing per axis               */
    int32_t has_map[3];                      /* a xi map exists for this axis (velocity in velocity_names) */
    int32_t npad;                            /* config.npad: periodic padding in cells (post:527-611)    */
    int32_t device;
    int32_t radius;                          /* search patch radius in cells; 0 = chosen from the data   */
    int32_t reserved[4];
} lfb_remap_desc;
/* For one output time: interpolate every field from the displaced positions x + xi to the grid nodes:
 * xi[axis] = parent array of xi_<vel> for that axis (NULL if no map), fields[v] / out[v] = parent arrays.
 * Piecewise-linear interpolation on the Delaunay triangulation of the normalised, periodically padded
 * cloud (scipy LinearNDInterpolator in the reference; tetrahedra on grids with three non-Flat axes), the
 * reference's edge fix on Bounded axes (numpy.interp along edge rows in 2-D, a 2-D interpolation on the
 * edge planes in 3-D), halos zero, NaN outside the hull.  The simplex search is shared by all fields.
 * Deliberate deviation: on a grid where a Bounded axis precedes a Periodic one the reference pads the wrong coordinate
 * column (post:523-611: `column_number` only advances inside the periodic branches); this function pads every
 * Periodic axis's own coordinate, as the reference's comments and documentation describe.  The two coincide on every
 * grid of the reference's examples and tests ((Periodic, Flat, Bounded), (Periodic, Periodic, Bounded)).
 * xi / fields / out may be host pointers or device pointers of desc.device (e.g. from lfb_kept_info): the copies
 * use unified addressing. */
int lfb_remap_to_mean(const lfb_remap_desc *desc, const double *const *xi, int n_fields, const double *const *fields,
                      double *const *out);
/* the work arrays of the 3-D remap are kept between calls (one call per output time); this returns them to the device */
int lfb_remap_release(void);

/* ---- multi-GPU: y-slab decomposition, one process per GPU -------------------------------------
 * With nranks > 1 every call that produces exported arrays or (re)initialises the state exchanges y-halo rows with
 * the neighbour ranks and is therefore COLLECTIVE: lfb_init_from_data, lfb_step / lfb_run_until, lfb_get_tracer,
 * lfb_get_tendency, lfb_output, lfb_output_f32, lfb_output_async, lfb_output_keep and lfb_kept_gather must be called
 * by all ranks in the same order (a rank that reads a diagnostic on its own deadlocks the others). */
/* NCCL unique id (128 bytes) created on rank 0 and broadcast by the host (torch.distributed / MPI). */
int lfb_comm_unique_id(void *id128);
/* Join the communicator.  Rank r owns rows [r*Ny/R, (r+1)*Ny/R) of the global grid; tracer y-halos
 * are exchanged with ranks r-1 / r+1 once per RK3 stage (ncclSend/ncclRecv over NVLink) instead of
 * being filled locally.  Stands in for Oceananigans' Distributed halo fill + compute_buffer_tendencies!
 * (compute_lagrangian_filter_buffer_tendencies.jl:8-38). */
int lfb_comm_init(lfb_handle *h, const void *id128, int rank, int nranks);

/* ---- introspection ---------------------------------------------------------------------------- */
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t lfb_kernel_launches(const lfb_handle *h);
/* device time (ms) spent in RK3 stage kernels since the last call with reset != 0, measured with
 * CUDA events on the compute stream, and the number of stage launches it covers */
int lfb_stage_time_ms(lfb_handle *h, int reset, double *ms, int64_t *launches);
/* CUDA-event stopwatch on the compute stream (the stream every stage kernel is launched on):
 * lfb_timer_start records an event, lfb_timer_stop records a second one, waits for it and returns
 * the device time between the two in milliseconds. */
int lfb_timer_start(lfb_handle *h);
int lfb_timer_stop(lfb_handle *h, double *ms);
/* name of the stage kernel in use: "generic" | "generic-strict" | "ztma" */
const char *lfb_kernel_name(const lfb_handle *h);
/* raw device pointers for zero-copy hosts (torch): what = 0 tracer (current), 1 tendency,
 * 2 snapshot slot (idx = slot * 16 + field_id); internal padded layout, see lfb_layout */
void *lfb_device_ptr(lfb_handle *h, int what, int idx);
/* internal layout: element (i,j,k) (0-based interior) lives at x_offset + i + pitch*((Hy + j) + rows*(Hz + k)) */
int lfb_layout(const lfb_handle *h, int64_t *x_offset, int64_t *pitch, int64_t *rows, int64_t *planes);

#ifdef __cplusplus
}
#endif
#endif /* LFB200_H */
