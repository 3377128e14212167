// lfb_api.cu -- host side of liblfb200.so: the C ABI declared in include/lfb200.h.
//
// Owns device memory, streams and the RK3 / snapshot bookkeeping; all arithmetic on field data is
// done by the kernels in lfb_generic.cu / lfb_ztma.cu.  No CPU fallback exists: every entry point
// that touches field data requires a CUDA device.
#include <cuda_runtime.h>
#include <nccl.h>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "lfb_internal.h"

// ---- kernel launchers (lfb_generic.cu, lfb_ztma.cu) ---------------------------------------------
extern "C" {
cudaError_t lfb_launch_stage_generic(const LfbStageParams *P, int strict, cudaStream_t stream);
cudaError_t lfb_launch_stage_ztma(const LfbStageParams *P, cudaStream_t stream);
int         lfb_ztma_supported(const LfbStageParams *P);
int         lfb_ztma_tile_rows(void);
cudaError_t lfb_launch_interp_inputs(const LfbInterpParams *P, cudaStream_t stream);
cudaError_t lfb_launch_fill_halo_axis(double *base, int64_t field_stride, int n_fields, const LfbLayout *L,
                                      int ax, int mode, cudaStream_t stream);
cudaError_t lfb_launch_import(double *dst, const void *src, int is_f32, const LfbLayout *L, int ex, int ey, int ez,
                              cudaStream_t stream);
cudaError_t lfb_launch_export(void *dst, int is_f32, const double *src, const LfbLayout *L, int ex, int ey, int ez,
                              cudaStream_t stream);
cudaError_t lfb_launch_centre_velocity(double *dst, const double *vel, const LfbLayout *L, int ax, double sign,
                                       cudaStream_t stream);
cudaError_t lfb_launch_scale(double *dst, const double *src, double factor, int64_t n, cudaStream_t stream);
cudaError_t lfb_launch_mask_immersed(double *U, const double *imm, int64_t len, int n_fields, cudaStream_t stream);
cudaError_t lfb_launch_build_face_codes(double *codes, const double *imm, const LfbLayout *L, cudaStream_t stream);
cudaError_t lfb_launch_output(double *out, const LfbOutputParams *P, cudaStream_t stream);
cudaError_t lfb_launch_combine(double *out, const double *fwd, const double *lo, const double *hi, double w,
                               double sign, int64_t n, cudaStream_t stream);
cudaError_t lfb_launch_combine_mixed(double *out, const void *fwd, int fwd_f32, const void *lo, int lo_f32, const void *hi, int hi_f32,
                                     double w, double sign, int64_t n, cudaStream_t stream);
cudaError_t lfb_launch_pack_y(double *buf, double *base, int64_t field_stride, int n_fields, const LfbLayout *L,
                              int side, int unpack, cudaStream_t stream);
}

// ---- error plumbing ------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" void lfb_internal_set_error(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(LFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail(LFB_ERR_COMM, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

#define N_FIELD_IDS (3 + LFB_MAX_VARS)

struct lfb_handle {
    lfb_desc   desc;
    LfbLayout  L;
    int        device;
    int        n_maps, n_fields, n_items, n_tracers, sub;   // sub = tracers per item (1 or 2)
    int        map_axis[3];
    LfbItem    items[LFB_MAX_ITEMS];
    int        single_exp, n_pairs;
    // device storage
    double    *U[2];
    double    *G;
    double    *scratch;                    // one field
    double    *mask;
    double    *imm;                        // immersed-cell flags (desc.immersed), set by lfb_set_immersed
    double    *dzc;                        // partial-cell heights (optional)
    double    *codes;                      // face codes of the immersed grid for the TMA z-march kernel
    int        imm_set;
    double     last_dt;                    // AB2: the previous step's dt (an Euler step whenever it changes)
    double    *zero_field;                 // velocity components that are not in velocity_names (TMA z-march path)
    double    *cur_in[N_FIELD_IDS];           // inputs at the current stage time (update_input_data!)
    // TMA z-march path: two sets of stage-time velocity fields; while a stage reads one, its helper warps fill the
    // other with the velocities of the next stage time (set 0 = cur_in[0..2], set 1 = vin_alt)
    double    *vin_alt[3];
    struct VinTag { int valid; double tau; int dir, s1, s2; uint64_t g1, g2; double t1, t2; } vin_tag[2];
    std::vector<uint64_t> slot_gen;        // bumped by every upload into the slot
    void      *staging;                    // dense parent array, device side
    size_t     staging_bytes;
    std::vector<double *> slot_field;      // [slot * N_FIELD_IDS + field_id]
    std::vector<double>   slot_time;
    std::vector<uint32_t> slot_valid;      // bit per field id
    std::vector<cudaEvent_t> slot_ready;   // import finished (recorded on the copy stream)
    std::vector<cudaEvent_t> slot_used;    // last stage kernel reading the slot (compute stream)
    std::vector<int>      slot_pending;    // compute stream has not yet waited for slot_ready
    int        cur;                        // index of the current U buffer
    double     time;
    int64_t    iteration;
    int        direction;
    cudaStream_t compute, copy, d2h;
    std::vector<void *> export_ring;       // dense parent arrays for asynchronous outputs
    std::vector<cudaEvent_t> ev_export, ev_fetched;
    std::vector<int> ring_used;
    int        export_next;
    cudaEvent_t  ev_tmp, ev_t0, ev_t1, ev_pre;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_free;   // recycled stage-timing event pairs
    // statistics
    int64_t    launches;
    double     stage_ms;
    int64_t    stage_launches;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;   // pending stage timings
    int        use_zmarch;
    int        fold;                       // the TMA kernel runs the (Flat y) grid folded: rows of 32 cells
    int        halos_valid;                // halos of U[cur] are up to date (kept so by the stage kernels)
    // distributed y-slabs
    cudaStream_t comm_stream;
    cudaEvent_t  ev_strips, ev_halo;
    ncclComm_t comm;
    int        rank, nranks;
    double    *send_lo, *send_hi, *recv_lo, *recv_hi;
    size_t     halo_elems;
    // device-resident outputs (lfb_output_keep ...): dense parent arrays
    struct Kept { void *p; int is_f32; int e[3]; cudaEvent_t fetched; };   // fetched: a pending asynchronous fetch
    std::vector<Kept> kept;
    // released kept arrays, reused by size: cudaMalloc / cudaFree of GB-sized blocks cost milliseconds and cudaFree
    // synchronises the device, which would stall the queue of time steps at every output time
    struct Pooled { void *p; size_t bytes; cudaEvent_t fetched; };
    std::vector<Pooled> kept_pool;
};

static inline int field_present(const lfb_handle *h, int field_id)
{
    if (field_id < 0) return 0;
    if (field_id < 3) return (h->desc.velocity_mask >> field_id) & 1;
    return field_id - 3 < h->desc.n_vars;
}

// interior extent of a field along each axis (N, or N+1 for a velocity on its own Bounded axis)
static void parent_extents(const lfb_handle *h, int field_id, int e[3])
{
    for (int ax = 0; ax < 3; ++ax) {
        int n = h->L.N[ax];
        if (field_id == ax && h->L.topo[ax] == LFB_BOUNDED && !(ax == 1 && h->nranks > 1 && h->rank != h->nranks - 1)) n += 1;
        e[ax] = n + 2 * h->L.H[ax];
    }
}

extern "C" int lfb_version(void) { return LFB_VERSION; }
extern "C" const char *lfb_last_error(void) { return g_err; }

static void build_layout(lfb_handle *h)
{
    LfbLayout &L = h->L;
    const lfb_desc &d = h->desc;
    for (int ax = 0; ax < 3; ++ax) {
        L.N[ax] = d.size[ax];
        L.H[ax] = d.halo[ax];
        L.topo[ax] = d.topology[ax];
        L.NG[ax] = d.size[ax];
        L.goff[ax] = 0;
        L.connected_lo[ax] = L.connected_hi[ax] = 0;
        // periodic images are written by the stage kernels when every cell has at most one image per axis
        L.fuse_halo[ax] = d.topology[ax] == LFB_PERIODIC && d.halo[ax] > 0 && d.size[ax] >= 2 * d.halo[ax];
    }
    L.xoff = L.H[0] > 0 ? LFB_XPAD : 0;
    int need = L.xoff + L.N[0] + L.H[0] + 1;
    L.pitch = (need + 15) / 16 * 16;
    L.rows = L.N[1] + 2 * L.H[1] + 1;
    L.planes = L.N[2] + 2 * L.H[2] + 1;
    L.plane = (int64_t)L.pitch * L.rows;
    L.len = L.plane * L.planes;
    L.stride[0] = 1; L.stride[1] = L.pitch; L.stride[2] = L.plane;
}

// stage kernel selection: the z-marching kernel when the grid / options allow it, else the generic one
static int select_kernel(lfb_handle *h)
{
    LfbStageParams P;
    memset(&P, 0, sizeof(P));
    P.L = h->L;
    P.advect = h->desc.advection;
    P.relax = h->desc.relaxation;
    P.mask = h->mask;
    P.imm = h->imm;
    for (int ax = 0; ax < 3; ++ax) P.vel_present[ax] = (h->desc.velocity_mask >> ax) & 1;
    P.j0 = 0; P.j1 = h->L.N[1];
    // 0 generic, 3 TMA-staged z-march (what `auto` picks when the grid allows it; 2 from lfb_ztma_supported: folded, Flat y)
    const int sup = lfb_ztma_supported(&P);
    const int fold_v = ((h->desc.velocity_mask >> 1) & 1) && h->desc.n_vars >= LFB_MAX_VARS;   // v travels in the last variable slot
    const int ok_tma = sup && !h->desc.strict && !(sup == 2 && fold_v);
    const char *need = "a 3-D grid (or a Flat y axis with a non-Flat x and z) with Nx >= 8, Ny >= 4, Nz >= 8, halos >= 3, at least "
                       "one velocity component, WENO(order=5) advection, whole cells if the grid is immersed (no cell_heights) and strict = 0";
    h->use_zmarch = 0;
    h->fold = 0;
    switch (h->desc.kernel) {
    case LFB_KERNEL_AUTO:
        // folded (Flat y) grids: a block marches Nz planes one after the other, so short lines (few blocks) are faster on
        // the generic kernel, which spreads every cell over the GPU: 400 x 80 runs 0.039 ms per stage there against
        // 0.071 ms folded, 8192 x 1024 runs 4.2 ms against 1.16 ms (scratch/bench_2d.py)
        h->use_zmarch = (ok_tma && !(sup == 2 && h->L.N[0] < 2048)) ? 3 : 0;
        break;
    case LFB_KERNEL_GENERIC: break;
    case LFB_KERNEL_ZTMA:
        if (!ok_tma) return fail(LFB_ERR_ARG, "the TMA z-marching kernel needs %s (and cuTensorMapEncodeTiled from the driver)", need);
        h->use_zmarch = 3;
        break;
    default:
        return fail(LFB_ERR_ARG, "kernel id %d: the cp.async z-march kernels (ids 2, 3) were retired; use AUTO, GENERIC or ZTMA", h->desc.kernel);
    }
    if (h->use_zmarch == 3 && sup == 2) {
        h->fold = 1;
        h->L.fuse_halo[0] = 0;           // the folded kernel does not write the periodic x images: filled before each stage
        h->halos_valid = 0;
    }
    return LFB_OK;
}

static int create_impl(lfb_handle *h)
{
    const lfb_desc *desc = &h->desc;
    build_layout(h);
    h->single_exp = desc->n_coeffs == 0;
    h->n_pairs = h->single_exp ? 1 : desc->n_coeffs;
    h->sub = h->single_exp ? 1 : 2;
    h->n_maps = 0;
    if (desc->with_maps)
        for (int ax = 0; ax < 3; ++ax) if ((desc->velocity_mask >> ax) & 1) h->map_axis[h->n_maps++] = ax;
    h->n_fields = desc->n_vars + h->n_maps;
    h->n_items = h->n_fields * h->n_pairs;
    if (h->n_items > LFB_MAX_ITEMS) return fail(LFB_ERR_ARG, "too many tracers (%d items > %d)", h->n_items, LFB_MAX_ITEMS);
    h->n_tracers = h->n_items * h->sub;
    for (int f = 0; f < h->n_fields; ++f)
        for (int p = 0; p < h->n_pairs; ++p) {
            LfbItem &it = h->items[f * h->n_pairs + p];
            it.is_map = f >= desc->n_vars;
            it.src = it.is_map ? h->map_axis[f - desc->n_vars] : f;
            it.tracer_c = (f * h->n_pairs + p) * h->sub;
            it.pad = 0;
            it.c = desc->c[p];
            it.d = h->single_exp ? 0.0 : desc->d[p];
        }
    const LfbLayout &L = h->L;
    const size_t fbytes = (size_t)L.len * sizeof(double);
    const int nt = h->n_tracers > 0 ? h->n_tracers : 1;
    CU(cudaMalloc(&h->U[0], fbytes * nt));
    CU(cudaMalloc(&h->U[1], fbytes * nt));
    CU(cudaMalloc(&h->G, fbytes * nt));
    CU(cudaMalloc(&h->scratch, fbytes));
    CU(cudaMemset(h->U[0], 0, fbytes * nt));
    CU(cudaMemset(h->U[1], 0, fbytes * nt));
    CU(cudaMemset(h->G, 0, fbytes * nt));
    CU(cudaMemset(h->scratch, 0, fbytes));
    if (desc->relaxation) { CU(cudaMalloc(&h->mask, fbytes)); CU(cudaMemset(h->mask, 0, fbytes)); }
    if (desc->immersed) { CU(cudaMalloc(&h->imm, fbytes)); CU(cudaMemset(h->imm, 0, fbytes)); }
    for (int f = 0; f < N_FIELD_IDS; ++f)
        if (field_present(h, f)) { CU(cudaMalloc(&h->cur_in[f], fbytes)); CU(cudaMemset(h->cur_in[f], 0, fbytes)); }
    h->staging_bytes = fbytes;     // dense parent arrays are never larger than the padded layout
    CU(cudaMalloc(&h->staging, h->staging_bytes));
    const int ns = desc->n_slots;
    h->slot_field.assign((size_t)ns * N_FIELD_IDS, nullptr);
    h->slot_time.assign(ns, 0.0);
    h->slot_valid.assign(ns, 0u);
    h->slot_pending.assign(ns, 0);
    h->slot_gen.assign(ns, 0);
    h->slot_ready.assign(ns, nullptr);
    h->slot_used.assign(ns, nullptr);
    for (int s = 0; s < ns; ++s) {
        CU(cudaEventCreateWithFlags(&h->slot_ready[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->slot_used[s], cudaEventDisableTiming));
        for (int f = 0; f < N_FIELD_IDS; ++f)
            if (field_present(h, f)) {
                CU(cudaMalloc(&h->slot_field[(size_t)s * N_FIELD_IDS + f], fbytes));
                CU(cudaMemset(h->slot_field[(size_t)s * N_FIELD_IDS + f], 0, fbytes));
            }
    }
    CU(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking));
    {
        // the small layout kernels behind the uploads (copy stream) and the downloads go first: a stage kernel that fills
        // every SM would otherwise keep the import of the next snapshot waiting until its own blocks are all issued
        int lo_prio = 0, hi_prio = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CU(cudaStreamCreateWithPriority(&h->copy, cudaStreamNonBlocking, hi_prio));
        CU(cudaStreamCreateWithPriority(&h->d2h, cudaStreamNonBlocking, hi_prio));
    }
    CU(cudaEventCreateWithFlags(&h->ev_tmp, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
    CU(cudaEventCreate(&h->ev_t0));
    CU(cudaEventCreate(&h->ev_t1));
    CU(cudaDeviceSynchronize());
    int rc = select_kernel(h);
    if (rc) return rc;
    if (h->use_zmarch == 3)
        for (int f = 0; f < 3; ++f)
            if (field_present(h, f)) { CU(cudaMalloc(&h->vin_alt[f], fbytes)); CU(cudaMemset(h->vin_alt[f], 0, fbytes)); }
    return LFB_OK;
}

extern "C" int lfb_create(const lfb_desc *desc, lfb_handle **out)
{
    if (!desc || !out) return fail(LFB_ERR_ARG, "null argument");
    *out = nullptr;
    for (int ax = 0; ax < 3; ++ax) {
        if (desc->size[ax] < 1) return fail(LFB_ERR_ARG, "size[%d] = %d", ax, desc->size[ax]);
        if (desc->topology[ax] < 0 || desc->topology[ax] > 2) return fail(LFB_ERR_ARG, "topology[%d] invalid", ax);
        if (desc->topology[ax] == LFB_FLAT && (desc->size[ax] != 1 || desc->halo[ax] != 0))
            return fail(LFB_ERR_ARG, "Flat axis %d must have size 1 and halo 0", ax);
        {
            // halo the scheme's stencil (and, on an immersed grid, its boundary test) reaches into
            const int adv = desc->advection;
            const int need_halo = adv == LFB_ADVECTION_NONE ? 0 : (adv == LFB_ADVECTION_WENO5 || adv == LFB_ADVECTION_UPWIND5) ? 3 :
                                  (adv == LFB_ADVECTION_CENTERED2 || adv == LFB_ADVECTION_UPWIND1) ? 1 : 2;
            if (desc->topology[ax] != LFB_FLAT && desc->halo[ax] < need_halo)
                return fail(LFB_ERR_ARG, "advection scheme %d needs halo >= %d on axis %d (got %d)", adv, need_halo, ax, desc->halo[ax]);
        }
        if (desc->halo[ax] < 0 || desc->halo[ax] > 8) return fail(LFB_ERR_ARG, "halo[%d] out of range", ax);
        if (!(desc->spacing[ax] > 0)) return fail(LFB_ERR_ARG, "spacing[%d] must be positive", ax);
        if (desc->topology[ax] == LFB_PERIODIC && desc->size[ax] < desc->halo[ax])
            return fail(LFB_ERR_ARG, "periodic axis %d shorter than its halo", ax);
    }
    if (desc->n_vars < 0 || desc->n_vars > LFB_MAX_VARS) return fail(LFB_ERR_ARG, "n_vars out of range");
    if (desc->n_coeffs < 0 || desc->n_coeffs > LFB_MAX_COEFFS) return fail(LFB_ERR_ARG, "n_coeffs out of range");
    if (desc->n_slots < 2 || desc->n_slots > 4096) return fail(LFB_ERR_ARG, "n_slots must be in [2, 4096]");
    if (desc->advection < LFB_ADVECTION_NONE || desc->advection > LFB_ADVECTION_WENO3)
        return fail(LFB_ERR_ARG, "unknown advection id %d", desc->advection);
    if (desc->timestepper != LFB_TIMESTEPPER_RK3 && desc->timestepper != LFB_TIMESTEPPER_AB2)
        return fail(LFB_ERR_ARG, "unknown timestepper id %d", desc->timestepper);
    if (desc->immersed != 0 && desc->immersed != 1) return fail(LFB_ERR_ARG, "immersed must be 0 or 1");
    if (desc->relaxation && !(desc->relax_timescale > 0)) return fail(LFB_ERR_ARG, "relax_timescale must be set");
    for (int q = 0; q < 4; ++q) if (desc->reserved[q]) return fail(LFB_ERR_ARG, "reserved fields must be zero");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(LFB_ERR_CUDA, "no CUDA device available (%s); liblfb200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (desc->device < 0 || desc->device >= ndev) return fail(LFB_ERR_ARG, "device %d not in [0,%d)", desc->device, ndev);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, desc->device));
    if (prop.major != 10)
        return fail(LFB_ERR_CUDA, "device %d is sm_%d%d; liblfb200 is built for sm_100a only", desc->device, prop.major, prop.minor);
    CU(cudaSetDevice(desc->device));

    lfb_handle *h = new lfb_handle();
    // every member that lfb_destroy looks at is null / empty from here on, so a failure anywhere below can be
    // routed through lfb_destroy without touching garbage or leaking what was already allocated
    h->desc = *desc;
    h->device = desc->device;
    h->U[0] = h->U[1] = h->G = h->scratch = h->mask = h->zero_field = h->imm = h->dzc = h->codes = nullptr;
    h->imm_set = 0; h->last_dt = INFINITY;
    h->staging = nullptr; h->staging_bytes = 0;
    for (int f = 0; f < N_FIELD_IDS; ++f) h->cur_in[f] = nullptr;
    for (int f = 0; f < 3; ++f) h->vin_alt[f] = nullptr;
    memset(h->vin_tag, 0, sizeof(h->vin_tag));
    h->compute = h->copy = h->d2h = h->comm_stream = nullptr;
    h->ev_tmp = h->ev_t0 = h->ev_t1 = h->ev_pre = h->ev_strips = h->ev_halo = nullptr;
    h->comm = nullptr; h->rank = 0; h->nranks = 1; h->halos_valid = 0;
    h->send_lo = h->send_hi = h->recv_lo = h->recv_hi = nullptr; h->halo_elems = 0;
    h->export_next = 0;
    h->cur = 0; h->time = 0; h->iteration = 0; h->direction = 1;
    h->launches = 0; h->stage_ms = 0; h->stage_launches = 0;
    h->use_zmarch = 0; h->fold = 0;
    int rc = create_impl(h);
    if (rc) { lfb_destroy(h); return rc; }
    *out = h;
    return LFB_OK;
}

extern "C" int lfb_destroy(lfb_handle *h)
{
    if (!h) return LFB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    // tolerates partially built handles (lfb_create routes its failures through here)
    if (h->comm) ncclCommDestroy(h->comm);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    for (cudaEvent_t e : {h->ev_strips, h->ev_halo, h->ev_tmp, h->ev_t0, h->ev_t1, h->ev_pre}) if (e) cudaEventDestroy(e);
    for (double *p : h->slot_field) if (p) cudaFree(p);
    for (cudaEvent_t e : h->slot_ready) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->slot_used) if (e) cudaEventDestroy(e);
    for (auto &pr : h->timing) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto &pr : h->timing_free) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (double *p : {h->U[0], h->U[1], h->G, h->scratch, h->mask, h->imm, h->dzc, h->codes, h->zero_field, h->send_lo, h->send_hi, h->recv_lo, h->recv_hi})
        if (p) cudaFree(p);
    for (int f = 0; f < N_FIELD_IDS; ++f) if (h->cur_in[f]) cudaFree(h->cur_in[f]);
    for (int f = 0; f < 3; ++f) if (h->vin_alt[f]) cudaFree(h->vin_alt[f]);
    if (h->staging) cudaFree(h->staging);
    for (void *p : h->export_ring) if (p) cudaFree(p);
    for (auto &k : h->kept) { if (k.p) cudaFree(k.p); if (k.fetched) cudaEventDestroy(k.fetched); }
    for (auto &k : h->kept_pool) { if (k.p) cudaFree(k.p); if (k.fetched) cudaEventDestroy(k.fetched); }
    for (cudaEvent_t e : h->ev_export) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_fetched) if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {h->compute, h->copy, h->d2h}) if (st) cudaStreamDestroy(st);
    delete h;
    return LFB_OK;
}

extern "C" int lfb_sync(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->copy));
    CU(cudaStreamSynchronize(h->compute));
    CU(cudaStreamSynchronize(h->d2h));
    return LFB_OK;
}

extern "C" int lfb_wait_transfers(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->copy));
    CU(cudaStreamSynchronize(h->d2h));
    return LFB_OK;
}

extern "C" int lfb_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr || bytes == 0) return fail(LFB_ERR_ARG, "bad argument");
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return LFB_OK;
}

extern "C" int lfb_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return LFB_OK;
}

extern "C" int lfb_n_tracers(const lfb_handle *h) { return h ? h->n_tracers : LFB_ERR_ARG; }

extern "C" int lfb_tracer_index(const lfb_handle *h, int field, int pole, int sine)
{
    if (!h || field < 0 || field >= h->n_fields || pole < 0 || pole >= h->n_pairs || sine < 0 || sine >= h->sub)
        return fail(LFB_ERR_ARG, "no tracer (field %d, pole %d, sine %d)", field, pole, sine);
    return (field * h->n_pairs + pole) * h->sub + sine;
}

extern "C" int64_t lfb_parent_len(const lfb_handle *h, int field_id)
{
    if (!h) return LFB_ERR_ARG;
    int e[3];
    parent_extents(h, field_id, e);
    return (int64_t)e[0] * e[1] * e[2];
}

// ---- input series --------------------------------------------------------------------------------
static int upload_impl(lfb_handle *h, int slot, int field_id, const void *parent, int is_f32, double time, int wait = 1)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (slot < 0 || slot >= h->desc.n_slots) return fail(LFB_ERR_ARG, "slot %d out of range", slot);
    if (!field_present(h, field_id)) return fail(LFB_ERR_ARG, "field id %d is not part of this handle", field_id);
    CU(cudaSetDevice(h->device));
    int e[3];
    parent_extents(h, field_id, e);
    const size_t n = (size_t)e[0] * e[1] * e[2];
    const size_t bytes = n * (is_f32 ? sizeof(float) : sizeof(double));
    // the slot may still be read by stage kernels queued earlier
    CU(cudaStreamWaitEvent(h->copy, h->slot_used[slot], 0));
    CU(cudaMemcpyAsync(h->staging, parent, bytes, cudaMemcpyHostToDevice, h->copy));
    CU(cudaEventRecord(h->ev_tmp, h->copy));
    double *dst = h->slot_field[(size_t)slot * N_FIELD_IDS + field_id];
    CU(lfb_launch_import(dst, h->staging, is_f32, &h->L, e[0], e[1], e[2], h->copy));
    h->launches++;
    CU(cudaEventRecord(h->slot_ready[slot], h->copy));
    h->slot_pending[slot] = 1;
    if (h->slot_valid[slot] && h->slot_time[slot] != time) h->slot_valid[slot] = 0;
    h->slot_time[slot] = time;
    h->slot_valid[slot] |= 1u << field_id;
    h->slot_gen[slot]++;
    // the host buffer is only borrowed: wait for the H2D copy (not for the import kernel)
    if (wait) CU(cudaEventSynchronize(h->ev_tmp));
    return LFB_OK;
}

extern "C" int lfb_upload_snapshot(lfb_handle *h, int slot, int field_id, const double *parent, double time)
{
    return upload_impl(h, slot, field_id, parent, 0, time);
}

extern "C" int lfb_upload_snapshot_f32(lfb_handle *h, int slot, int field_id, const float *parent, double time)
{
    return upload_impl(h, slot, field_id, parent, 1, time);
}

extern "C" int lfb_upload_snapshot_async(lfb_handle *h, int slot, int field_id, const void *parent, int is_f32, double time)
{
    return upload_impl(h, slot, field_id, parent, is_f32 != 0, time, 0);
}

extern "C" int lfb_invalidate_slot(lfb_handle *h, int slot)
{
    if (!h || slot < 0 || slot >= h->desc.n_slots) return fail(LFB_ERR_ARG, "slot out of range");
    h->slot_valid[slot] = 0;
    return LFB_OK;
}

extern "C" int lfb_set_slot_time(lfb_handle *h, int slot, double time)
{
    if (!h || slot < 0 || slot >= h->desc.n_slots) return fail(LFB_ERR_ARG, "slot out of range");
    h->slot_time[slot] = time;
    return LFB_OK;
}

extern "C" int lfb_set_direction(lfb_handle *h, int direction)
{
    if (!h || (direction != 1 && direction != -1)) return fail(LFB_ERR_ARG, "direction must be +1 or -1");
    h->direction = direction;
    return LFB_OK;
}

static uint32_t required_mask(const lfb_handle *h)
{
    uint32_t m = (uint32_t)h->desc.velocity_mask & 7u;
    for (int q = 0; q < h->desc.n_vars; ++q) m |= 1u << (3 + q);
    return m;
}

// slots bracketing pass time tau: s1 = latest slot with time <= tau, s2 = earliest slot with time > tau
static int bracket(const lfb_handle *h, double tau, int *s1, int *s2, double *w1, double *w2, int *interp, bool quiet = false)
{
    const uint32_t need = required_mask(h);
    int a = -1, b = -1;
    for (int s = 0; s < h->desc.n_slots; ++s) {
        if ((h->slot_valid[s] & need) != need) continue;
        const double t = h->slot_time[s];
        if (t <= tau) { if (a < 0 || t > h->slot_time[a]) a = s; }
        else          { if (b < 0 || t < h->slot_time[b]) b = s; }
    }
    if (a < 0) return quiet ? LFB_ERR_NO_DATA : fail(LFB_ERR_NO_DATA, "no resident snapshot at or before pass time %.17g", tau);
    *s1 = a;
    if (h->slot_time[a] == tau) { *s2 = a; *w1 = 1; *w2 = 0; *interp = 0; return LFB_OK; }
    if (b < 0) return quiet ? LFB_ERR_NO_DATA : fail(LFB_ERR_NO_DATA, "no resident snapshot after pass time %.17g", tau);
    const double t1 = h->slot_time[a], t2 = h->slot_time[b];
    const double n = (tau - t1) / (t2 - t1);
    *s2 = b; *w2 = n; *w1 = 1 - n; *interp = 1;
    return LFB_OK;
}

extern "C" int lfb_set_mask(lfb_handle *h, const double *parent)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (!h->mask) return fail(LFB_ERR_STATE, "handle was created without relaxation");
    CU(cudaSetDevice(h->device));
    int e[3];
    parent_extents(h, -1, e);
    CU(cudaStreamSynchronize(h->copy));
    CU(cudaMemcpyAsync(h->staging, parent, (size_t)e[0] * e[1] * e[2] * sizeof(double), cudaMemcpyHostToDevice, h->compute));
    CU(lfb_launch_import(h->mask, h->staging, 0, &h->L, e[0], e[1], e[2], h->compute));
    h->launches++;
    CU(cudaStreamSynchronize(h->compute));
    return LFB_OK;
}

// face codes of an immersed grid for the TMA z-march kernel (rebuilt when the slab decomposition changes the global indices)
static int build_face_codes(lfb_handle *h)
{
    if (h->use_zmarch != 3 || !h->imm) return LFB_OK;
    const size_t fbytes = (size_t)h->L.len * sizeof(double);
    if (!h->codes) { CU(cudaMalloc(&h->codes, fbytes)); }
    CU(cudaMemsetAsync(h->codes, 0, fbytes, h->compute));
    CU(lfb_launch_build_face_codes(h->codes, h->imm, &h->L, h->compute));
    h->launches++;
    CU(cudaStreamSynchronize(h->compute));
    return LFB_OK;
}

extern "C" int lfb_set_immersed(lfb_handle *h, const double *parent, const double *cell_heights)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (!h->imm) return fail(LFB_ERR_STATE, "handle was created with desc.immersed = 0");
    CU(cudaSetDevice(h->device));
    int e[3];
    parent_extents(h, -1, e);
    CU(cudaStreamSynchronize(h->copy));
    CU(cudaMemcpyAsync(h->staging, parent, (size_t)e[0] * e[1] * e[2] * sizeof(double), cudaMemcpyHostToDevice, h->compute));
    CU(lfb_launch_import(h->imm, h->staging, 0, &h->L, e[0], e[1], e[2], h->compute));
    h->launches++;
    CU(cudaStreamSynchronize(h->compute));
    if (cell_heights) {
        const size_t fbytes = (size_t)h->L.len * sizeof(double);
        if (!h->dzc) { CU(cudaMalloc(&h->dzc, fbytes)); }
        CU(cudaMemset(h->dzc, 0, fbytes));
        CU(cudaMemcpyAsync(h->staging, cell_heights, (size_t)e[0] * e[1] * e[2] * sizeof(double), cudaMemcpyHostToDevice, h->compute));
        CU(lfb_launch_import(h->dzc, h->staging, 0, &h->L, e[0], e[1], e[2], h->compute));
        h->launches++;
        CU(cudaStreamSynchronize(h->compute));
    } else if (h->dzc) { cudaFree(h->dzc); h->dzc = nullptr; }
    if (h->use_zmarch == 3 && h->dzc) {
        // partial cells change the face areas and cell volumes: generic kernel
        if (h->desc.kernel == LFB_KERNEL_ZTMA)
            return fail(LFB_ERR_ARG, "the TMA z-marching kernel handles immersed grids with whole cells only (no cell_heights)");
        if (h->fold)        // the generic kernel writes the periodic x images itself again
            h->L.fuse_halo[0] = h->L.topo[0] == LFB_PERIODIC && h->L.H[0] > 0 && h->L.N[0] >= 2 * h->L.H[0];
        h->use_zmarch = 0; h->fold = 0;
    }
    {
        int rc = build_face_codes(h);
        if (rc) return rc;
    }
    h->imm_set = 1;
    h->halos_valid = 0;
    return LFB_OK;
}

// ---- halo handling -------------------------------------------------------------------------------
// Periodic halos of the current tracer buffer (what the stencils read).  Bounded-axis halos are never
// read by the stencils and are only produced for exported arrays.  Axes whose images
// the stage kernels write themselves (LfbLayout::fuse_halo) only need this after the state was set from
// outside (force = 1).
static int fill_periodic_halos(lfb_handle *h, double *base, int n_fields, int force)
{
    for (int ax = 0; ax < 3; ++ax) {
        if (h->L.topo[ax] != LFB_PERIODIC) continue;
        if (ax == 1 && h->nranks > 1) continue;     // exchanged with the neighbour ranks instead
        if (h->L.fuse_halo[ax] && !force) continue;
        CU(lfb_launch_fill_halo_axis(base, h->L.len, n_fields, &h->L, ax, 0, h->compute));
        h->launches++;
    }
    return LFB_OK;
}

static int fill_all_halos(lfb_handle *h, double *base, int n_fields)
{
    for (int ax = 0; ax < 3; ++ax) {
        if (h->L.topo[ax] == LFB_FLAT || h->L.H[ax] == 0) continue;
        int mode = h->L.topo[ax] == LFB_PERIODIC ? 0 : 1;
        if (ax == 1 && h->nranks > 1) {
            // y halos come from the neighbour ranks (exchange_y); only the outer edges of a Bounded axis are mirrored here
            if (h->L.topo[1] != LFB_BOUNDED) continue;
            if (h->rank == 0) mode = 2;
            else if (h->rank == h->nranks - 1) mode = 3;
            else continue;
        }
        CU(lfb_launch_fill_halo_axis(base, h->L.len, n_fields, &h->L, ax, mode, h->compute));
        h->launches++;
    }
    return LFB_OK;
}

// y-halo exchange between neighbouring slabs: pack the H edge rows of every tracer, ncclSend/Recv
// them to ranks r-1 / r+1 (periodic wrap when the global y axis is Periodic), unpack into the halos.
static int exchange_y(lfb_handle *h, double *base, int n_fields, cudaStream_t st = nullptr)
{
    if (!st) st = h->compute;
    if (h->nranks <= 1 || h->L.H[1] == 0) return LFB_OK;
    const bool periodic = h->L.topo[1] == LFB_PERIODIC;
    const int lo = h->rank > 0 ? h->rank - 1 : (periodic ? h->nranks - 1 : -1);
    const int hi = h->rank < h->nranks - 1 ? h->rank + 1 : (periodic ? 0 : -1);
    const size_t per_field = (size_t)(h->L.N[0] + 2 * h->L.H[0]) * h->L.H[1] * (h->L.N[2] + 2 * h->L.H[2]);
    const size_t count = per_field * n_fields;
    if (lo >= 0) { CU(lfb_launch_pack_y(h->send_lo, base, h->L.len, n_fields, &h->L, 0, 0, st)); h->launches++; }
    if (hi >= 0) { CU(lfb_launch_pack_y(h->send_hi, base, h->L.len, n_fields, &h->L, 1, 0, st)); h->launches++; }
    // sends first (low edge, high edge), then receives (high halo, low halo): with two ranks on a periodic
    // axis both neighbours are the same peer and NCCL matches operations per peer in program order
    NC(ncclGroupStart());
    if (lo >= 0) NC(ncclSend(h->send_lo, count, ncclDouble, lo, h->comm, st));
    if (hi >= 0) NC(ncclSend(h->send_hi, count, ncclDouble, hi, h->comm, st));
    if (hi >= 0) NC(ncclRecv(h->recv_hi, count, ncclDouble, hi, h->comm, st));
    if (lo >= 0) NC(ncclRecv(h->recv_lo, count, ncclDouble, lo, h->comm, st));
    NC(ncclGroupEnd());
    if (lo >= 0) { CU(lfb_launch_pack_y(h->recv_lo, base, h->L.len, n_fields, &h->L, 0, 1, st)); h->launches++; }
    if (hi >= 0) { CU(lfb_launch_pack_y(h->recv_hi, base, h->L.len, n_fields, &h->L, 1, 1, st)); h->launches++; }
    return LFB_OK;
}

// Make the halos of U[cur] valid before a stage reads them.
static int refresh_halos(lfb_handle *h, double *base, int n_fields)
{
    const int force = !h->halos_valid;
    if (force && h->imm) {
        // mask_immersed_field! of update_state!: the stage kernels keep immersed cells at zero themselves, so this is
        // only needed after the state was set from outside (initialisation, lfb_set_tracer)
        if (!h->imm_set) return fail(LFB_ERR_STATE, "desc.immersed = 1 but lfb_set_immersed was not called");
        CU(lfb_launch_mask_immersed(base, h->imm, h->L.len, n_fields, h->compute));
        h->launches++;
    }
    int rc = fill_periodic_halos(h, base, n_fields, force);
    if (rc) return rc;
    if (h->nranks > 1 && force) { rc = exchange_y(h, base, n_fields); if (rc) return rc; }
    h->halos_valid = 1;
    return LFB_OK;
}

// ---- state ---------------------------------------------------------------------------------------
extern "C" int lfb_init_from_data(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    int s1, s2, interp; double w1, w2;
    int rc = bracket(h, 0.0, &s1, &s2, &w1, &w2, &interp);
    if (rc) return rc;
    if (interp) return fail(LFB_ERR_NO_DATA, "initialisation needs a resident snapshot at pass time 0");
    CU(cudaStreamWaitEvent(h->compute, h->slot_ready[s1], 0));
    h->slot_pending[s1] = 0;
    const LfbLayout &L = h->L;
    const double sign = h->direction < 0 ? -1.0 : 1.0;
    for (int f = 0; f < h->n_fields; ++f) {
        const bool is_map = f >= h->desc.n_vars;
        const double *src;
        if (!is_map) src = h->slot_field[(size_t)s1 * N_FIELD_IDS + 3 + f];
        else {
            const int ax = h->map_axis[f - h->desc.n_vars];
            CU(cudaMemsetAsync(h->scratch, 0, (size_t)L.len * sizeof(double), h->compute));
            CU(lfb_launch_centre_velocity(h->scratch, h->slot_field[(size_t)s1 * N_FIELD_IDS + ax], &L, ax, sign, h->compute));
            h->launches++;
            rc = fill_all_halos(h, h->scratch, 1);
            if (rc) return rc;
            rc = exchange_y(h, h->scratch, 1);
            if (rc) return rc;
            src = h->scratch;
        }
        for (int p = 0; p < h->n_pairs; ++p) {
            const double c = h->desc.c[p], d = h->single_exp ? 0.0 : h->desc.d[p];
            double fC, fS = 0;
            if (h->single_exp) fC = is_map ? (-1 / (c * c)) : 1 / c;
            else if (!is_map) { fC = c / (c * c + d * d); fS = d / (c * c + d * d); }
            else { const double n2 = c * c + d * d; fC = (d * d - c * c) / (n2 * n2); fS = -2 * c * d / (n2 * n2); }
            const int tc = (f * h->n_pairs + p) * h->sub;
            for (int b = 0; b < 2; ++b) {
                CU(lfb_launch_scale(h->U[b] + (int64_t)tc * L.len, src, fC, L.len, h->compute));
                h->launches++;
                if (!h->single_exp) {
                    CU(lfb_launch_scale(h->U[b] + (int64_t)(tc + 1) * L.len, src, fS, L.len, h->compute));
                    h->launches++;
                }
            }
        }
    }
    h->cur = 0; h->time = 0; h->iteration = 0;
    h->halos_valid = 0;
    if (h->imm && h->imm_set) {
        // the reference's first update_state! (mask_immersed_field!, update_lagrangian_filter_state.jl:22-24) runs before the
        // output writer sees iteration 0: the initial state is masked here, not only before the first stage
        for (int b = 0; b < 2; ++b) { CU(lfb_launch_mask_immersed(h->U[b], h->imm, L.len, h->n_tracers, h->compute)); h->launches++; }
    }
    return LFB_OK;
}

extern "C" int lfb_negate_maps(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    const LfbLayout &L = h->L;
    for (int f = h->desc.n_vars; f < h->n_fields; ++f)
        for (int p = 0; p < h->n_pairs; ++p)
            for (int s = 0; s < h->sub; ++s)
                for (int b = 0; b < 2; ++b) {
                    double *U = h->U[b] + (int64_t)((f * h->n_pairs + p) * h->sub + s) * L.len;
                    CU(lfb_launch_scale(U, U, -1.0, L.len, h->compute));
                    h->launches++;
                }
    return LFB_OK;
}

extern "C" int lfb_reset_clock(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    h->time = 0; h->iteration = 0; h->last_dt = INFINITY;      // Oceananigans' reset!(clock): last_dt = Inf
    return LFB_OK;
}

extern "C" double lfb_time(const lfb_handle *h) { return h ? h->time : NAN; }
extern "C" int64_t lfb_iteration(const lfb_handle *h) { return h ? h->iteration : -1; }

static int download(lfb_handle *h, const double *src, void *parent, int is_f32, int field_id)
{
    int e[3];
    parent_extents(h, field_id, e);
    const size_t n = (size_t)e[0] * e[1] * e[2];
    CU(cudaStreamSynchronize(h->copy));      // staging is shared with uploads
    CU(lfb_launch_export(h->staging, is_f32, src, &h->L, e[0], e[1], e[2], h->compute));
    h->launches++;
    CU(cudaMemcpyAsync(parent, h->staging, n * (is_f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, h->compute));
    CU(cudaStreamSynchronize(h->compute));
    return LFB_OK;
}

extern "C" int lfb_get_tracer(lfb_handle *h, int tracer, double *parent)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (tracer < 0 || tracer >= h->n_tracers) return fail(LFB_ERR_ARG, "tracer %d out of range", tracer);
    CU(cudaSetDevice(h->device));
    double *U = h->U[h->cur] + (int64_t)tracer * h->L.len;
    int rc = fill_all_halos(h, U, 1);
    if (rc) return rc;
    rc = exchange_y(h, U, 1);
    if (rc) return rc;
    return download(h, U, parent, 0, -1);
}

extern "C" int lfb_set_tracer(lfb_handle *h, int tracer, const double *parent)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (tracer < 0 || tracer >= h->n_tracers) return fail(LFB_ERR_ARG, "tracer %d out of range", tracer);
    CU(cudaSetDevice(h->device));
    int e[3];
    parent_extents(h, -1, e);
    CU(cudaStreamSynchronize(h->copy));
    CU(cudaMemcpyAsync(h->staging, parent, (size_t)e[0] * e[1] * e[2] * sizeof(double), cudaMemcpyHostToDevice, h->compute));
    for (int b = 0; b < 2; ++b) {
        CU(lfb_launch_import(h->U[b] + (int64_t)tracer * h->L.len, h->staging, 0, &h->L, e[0], e[1], e[2], h->compute));
        h->launches++;
    }
    CU(cudaStreamSynchronize(h->compute));
    h->halos_valid = 0;
    return LFB_OK;
}

// ---- time stepping -------------------------------------------------------------------------------
static lfb_handle::VinTag make_tag(const lfb_handle *h, double tau, int s1, int s2)
{
    lfb_handle::VinTag t;
    t.valid = 1; t.tau = tau; t.dir = h->direction; t.s1 = s1; t.s2 = s2;
    t.g1 = h->slot_gen[s1]; t.g2 = h->slot_gen[s2]; t.t1 = h->slot_time[s1]; t.t2 = h->slot_time[s2];
    return t;
}

static bool same_tag(const lfb_handle::VinTag &a, const lfb_handle::VinTag &b)
{
    return a.valid && b.valid && a.tau == b.tau && a.dir == b.dir && a.s1 == b.s1 && a.s2 == b.s2 && a.g1 == b.g1 &&
           a.g2 == b.g2 && a.t1 == b.t1 && a.t2 == b.t2;
}

// One fused RK3 stage at pass time tau.  tau_next (NaN: unknown) is the time of the stage that follows: the TMA
// z-march kernel interpolates that stage's velocities on the side.
static int launch_stage(lfb_handle *h, double tau, double dt, double gamma, double zeta, int first, int store_G, int update,
                        double tau_next = NAN)
{
    LfbStageParams P;
    memset(&P, 0, sizeof(P));
    P.L = h->L;
    P.U_old = h->U[h->cur];
    P.U_new = h->U[1 - h->cur];
    P.G = h->G;
    P.n_items = h->n_items;
    P.single_exp = h->single_exp;
    P.advect = h->desc.advection;
    P.relax = h->desc.relaxation;
    P.imm = h->imm;
    P.dzc = h->dzc;
    P.codes = (h->use_zmarch == 3 && update) ? h->codes : nullptr;
    memcpy(P.items, h->items, sizeof(LfbItem) * h->n_items);
    int s1, s2, interp; double w1, w2;
    int rc = bracket(h, tau, &s1, &s2, &w1, &w2, &interp);
    if (rc) return rc;
    const bool ztma = h->use_zmarch == 3 && update;
    const bool backward = h->direction < 0;
    // the next stage's bracketing snapshots, if they are resident already (the side job reads them)
    int n1 = -1, n2 = -1, ninterp = 0; double nw1 = 1, nw2 = 0;
    static const bool no_prefetch = getenv("LFB_NO_PREFETCH") != nullptr;      // A/B measurements only
    bool prefetch = ztma && !no_prefetch && h->L.N[0] % 2 == 0 /* 128-bit accesses up to column Nx */ && tau_next == tau_next && bracket(h, tau_next, &n1, &n2, &nw1, &nw2, &ninterp, true) == LFB_OK &&
                    (ninterp || backward);
    for (int s : {s1, s2, prefetch ? n1 : s1, prefetch ? n2 : s1})
        if (h->slot_pending[s]) { CU(cudaStreamWaitEvent(h->compute, h->slot_ready[s], 0)); h->slot_pending[s] = 0; }
    // update_input_data!: inputs at the stage time (aliased to the slot when no arithmetic is needed)
    int vset = -1;                     // velocity set read by this stage (TMA path), -1: aliased to a slot
    {
        LfbInterpParams I;
        memset(&I, 0, sizeof(I));
        I.len = h->L.len; I.interp = interp; I.w1 = w1; I.w2 = w2;
        const lfb_handle::VinTag want = make_tag(h, tau, s1, s2);
        const bool need_vel = interp || backward;
        if (ztma && need_vel) {
            for (int b = 0; b < 2; ++b) if (same_tag(h->vin_tag[b], want)) vset = b;
            if (vset < 0) {            // not prefetched (first stage after a start, re-tagged slots, ...): interpolate now
                vset = h->vin_tag[0].valid && !h->vin_tag[1].valid ? 1 : 0;
                h->vin_tag[vset] = want;
            } else I.interp = -1;      // marks "velocities already there"
        } else if (!ztma && need_vel) h->vin_tag[0].valid = 0;     // the generic path overwrites set 0
        for (int f = 0; f < N_FIELD_IDS; ++f) {
            if (!field_present(h, f)) continue;
            const double *a = h->slot_field[(size_t)s1 * N_FIELD_IDS + f];
            const double *b = h->slot_field[(size_t)s2 * N_FIELD_IDS + f];
            const bool neg = backward && f < 3;
            const double *use = a;
            if (f == 1 && ztma && h->fold) {
                // Flat y: v is not advected with; it forces its xi_v map from its two snapshots (last variable slot)
                P.var_fused = 1; P.var_w1 = w1; P.var_w2 = w2;
                P.var[LFB_MAX_VARS - 1] = a; P.var2[LFB_MAX_VARS - 1] = interp ? b : a;
                P.vsign = backward ? -1.0 : 1.0;
                continue;
            }
            if (f >= 3 && ztma) {
                // the TMA z-march kernel reads both snapshots of a variable and interpolates in its forcing term
                P.var_fused = 1; P.var_w1 = w1; P.var_w2 = w2;
                P.var[f - 3] = a; P.var2[f - 3] = interp ? b : a;
                continue;
            }
            if (interp || neg) {
                double *dst = (f < 3 && vset == 1) ? h->vin_alt[f] : h->cur_in[f];
                if (!(f < 3 && I.interp < 0)) {
                    const int q = I.n_fields++;
                    I.s1[q] = a; I.s2[q] = b; I.dst[q] = dst; I.negate[q] = neg;
                }
                use = dst;
            }
            if (f < 3) P.vel[f] = use; else P.var[f - 3] = use;
        }
        if (I.interp < 0) I.interp = interp;
        if (I.n_fields) { CU(lfb_launch_interp_inputs(&I, h->compute)); h->launches++; }
    }
    for (int ax = 0; ax < 3; ++ax) P.vel_present[ax] = (h->desc.velocity_mask >> ax) & 1;
    P.fold = ztma && h->fold;
    if (ztma)
        for (int ax = 0; ax < 3; ++ax)
            if (!P.vel_present[ax] && !(ax == 1 && h->fold)) {                  // a component that is not advected with: zero face velocity
                if (!h->zero_field) {
                    CU(cudaMalloc(&h->zero_field, (size_t)h->L.len * sizeof(double)));
                    CU(cudaMemsetAsync(h->zero_field, 0, (size_t)h->L.len * sizeof(double), h->compute));
                }
                P.vel[ax] = h->zero_field;
            }
    int pset = -1;
    if (prefetch) {
        // the set this stage does not read receives the next stage's velocities
        pset = vset == 0 ? 1 : 0;
        const lfb_handle::VinTag next = make_tag(h, tau_next, n1, n2);
        if (same_tag(h->vin_tag[pset], next)) pset = -1;           // already there
        else {
            P.pf.on = 1; P.pf.negate = backward; P.pf.w1 = ninterp ? nw1 : 1.0; P.pf.w2 = ninterp ? nw2 : 0.0;
            for (int f = 0; f < 3; ++f) {
                if (!field_present(h, f) || (f == 1 && h->fold)) continue;
                P.pf.s1[f] = h->slot_field[(size_t)n1 * N_FIELD_IDS + f];
                P.pf.s2[f] = h->slot_field[(size_t)(ninterp ? n2 : n1) * N_FIELD_IDS + f];
                P.pf.dst[f] = pset == 1 ? h->vin_alt[f] : h->cur_in[f];
            }
            h->vin_tag[pset].valid = 0;
        }
    }
    P.mask = h->mask;
    const double dx = h->desc.spacing[0], dy = h->desc.spacing[1], dz = h->desc.spacing[2];
    P.area[0] = dy * dz; P.area[1] = dx * dz; P.area[2] = dx * dy; P.volume = dx * dy * dz;
    P.spacing[0] = dx; P.spacing[1] = dy; P.spacing[2] = dz;
    P.relax_timescale = h->desc.relax_timescale;
    P.dt = dt; P.gamma = gamma; P.zeta = zeta;
    P.first_stage = first; P.store_G = store_G; P.update = update;
    P.j0 = 0; P.j1 = h->L.N[1];
    P.scratch = h->scratch;

    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!h->timing_free.empty()) { ev = h->timing_free.back(); h->timing_free.pop_back(); }
    else { CU(cudaEventCreate(&ev.first)); CU(cudaEventCreate(&ev.second)); }
    CU(cudaEventRecord(ev.first, h->compute));
    const int Ny = h->L.N[1];
    // rows per edge strip: one z-march tile row / the halo width
    const int strip = h->use_zmarch == 3 ? lfb_ztma_tile_rows() : h->L.H[1];
    const bool overlap = update && h->nranks > 1 && Ny >= 4 * strip;
    // Bounded y on the TMA kernel: only the tile rows next to the walls need the general instantiation (buffer schemes);
    // rows at least one tile row away from both walls run the all-periodic one
    const bool split_walls = ztma && !h->fold && !h->imm && h->L.topo[1] == LFB_BOUNDED && Ny >= 4 * strip;
    // Bounded x likewise: the x tiles whose faces all lie in 4 .. Nx - 2 (tiles 1 .. xt_hi - 1) need no buffer schemes
    const int n_xt = (h->L.N[0] + 31) / 32, xt_hi = h->L.N[0] >= 35 ? (h->L.N[0] - 35) / 32 + 1 : 0;
    const bool split_x = ztma && !h->fold && !h->imm && h->L.topo[0] == LFB_BOUNDED && xt_hi >= 3;
    auto launch_rows = [&](int j0, int j1, int jb0, int jb1, cudaStream_t st) -> int {
        P.j0 = j0; P.j1 = j1; P.jb0 = jb0; P.jb1 = jb1;
        const int g0 = h->L.goff[1] + j0, g1 = h->L.goff[1] + j1;
        P.plain_rows = h->L.topo[1] != LFB_BOUNDED || (split_walls && jb1 <= jb0 && g0 >= strip && g1 <= h->L.NG[1] - strip);
        if (ztma) CU(lfb_launch_stage_ztma(&P, st));
        else {
            CU(lfb_launch_stage_generic(&P, h->desc.strict, st));
            if (jb1 > jb0) { P.j0 = jb0; P.j1 = jb1; CU(lfb_launch_stage_generic(&P, h->desc.strict, st)); h->launches++; }
        }
        h->launches++;
        return LFB_OK;
    };
    auto launch = [&](int j0, int j1, int jb0, int jb1, cudaStream_t st) -> int {
        P.xt0 = 0; P.xt_n = 0; P.plain_cols = 0;
        if (!split_x) return launch_rows(j0, j1, jb0, jb1, st);
        // x tiles next to the west wall, next to the east wall (general instantiation), and the ones between
        int rc2;
        P.xt0 = 0; P.xt_n = 1; P.plain_cols = 0;
        if ((rc2 = launch_rows(j0, j1, jb0, jb1, st))) return rc2;
        P.xt0 = xt_hi; P.xt_n = n_xt - xt_hi;
        if ((rc2 = launch_rows(j0, j1, jb0, jb1, st))) return rc2;
        P.xt0 = 1; P.xt_n = xt_hi - 1; P.plain_cols = 1;
        return launch_rows(j0, j1, jb0, jb1, st);
    };
    if (!overlap && split_walls && h->nranks == 1) {
        rc = launch(0, strip, Ny - strip, Ny, h->compute);            // the two wall strips, one grid
        if (rc) return rc;
        rc = launch(strip, Ny - strip, 0, 0, h->compute);             // everything else
        if (rc) return rc;
    } else if (!overlap) {
        rc = launch(0, Ny, 0, 0, h->compute);
        if (rc) return rc;
        if (update && h->nranks > 1) { rc = exchange_y(h, P.U_new, h->n_tracers); if (rc) return rc; }
    } else {
        // Slab decomposition (stands in for Oceananigans' async halo fill + compute_buffer_tendencies!,
        // compute_lagrangian_filter_buffer_tendencies.jl:8-38).  The two edge strips are ONE grid on the
        // high-priority communication stream, followed there by their halo exchange; the interior is a second
        // grid on the compute stream.  Nothing orders the two grids of a stage, so the strip blocks are placed
        // first and the interior blocks fill the SMs behind them without a wave boundary in between.
        CU(cudaEventRecord(h->ev_pre, h->compute));               // the previous stage's interior + this stage's inputs
        CU(cudaStreamWaitEvent(h->comm_stream, h->ev_pre, 0));
        rc = launch(0, strip, Ny - strip, Ny, h->comm_stream);
        if (rc) return rc;
        rc = exchange_y(h, P.U_new, h->n_tracers, h->comm_stream);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev_halo, h->comm_stream));
        rc = launch(strip, Ny - strip, 0, 0, h->compute);
        if (rc) return rc;
        CU(cudaStreamWaitEvent(h->compute, h->ev_halo, 0));   // the next stage reads the strips and the received halo rows
    }
    CU(cudaEventRecord(ev.second, h->compute));
    h->timing.push_back(ev);
    if (pset >= 0) h->vin_tag[pset] = make_tag(h, tau_next, n1, n2);
    for (int s : {s1, s2, prefetch ? n1 : s1, prefetch ? n2 : s1}) CU(cudaEventRecord(h->slot_used[s], h->compute));
    return LFB_OK;
}

static int drain_timings(lfb_handle *h, bool wait)
{
    size_t done = 0;
    for (auto &pr : h->timing) {
        if (!wait && cudaEventQuery(pr.second) != cudaSuccess) break;
        if (wait) CU(cudaEventSynchronize(pr.second));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, pr.first, pr.second));
        h->stage_ms += ms;
        h->stage_launches++;
        h->timing_free.push_back(pr);
        ++done;
    }
    h->timing.erase(h->timing.begin(), h->timing.begin() + done);
    return LFB_OK;
}

extern "C" int lfb_step(lfb_handle *h, double dt)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    if (!(dt > 0)) return fail(LFB_ERR_ARG, "dt must be positive");
    CU(cudaSetDevice(h->device));
    if (h->timing.size() > 64) drain_timings(h, false);
    if (h->desc.timestepper == LFB_TIMESTEPPER_AB2) {
        // QuasiAdamsBashforth2 (lagrangian_filter_ab2_step.jl:24-51 -> Oceananigans _ab2_step_field!):
        // U += dt ((3/2 + chi) G^n - (1/2 + chi) G^-), an Euler step (chi = -1/2) at the first step after a clock reset
        // and whenever dt differs from the previous step's.  One fused stage: G^n of the current state at the current
        // time, the update, and G^n kept as the next step's G^-.
        const bool euler = dt != h->last_dt;
        const double chi = euler ? -0.5 : h->desc.chi;
        int rc = refresh_halos(h, h->U[h->cur], h->n_tracers);
        if (rc) return rc;
        rc = launch_stage(h, h->time, dt, 1.5 + chi, -(0.5 + chi), euler, 1, 1, h->time + dt);
        if (rc) return rc;
        h->cur ^= 1;
        h->time = h->time + dt;
        h->last_dt = dt;
        h->iteration += 1;
        return LFB_OK;
    }
    // Oceananigans RungeKutta3 (Le & Moin 1991) coefficients and stage times
    const double g1 = 8.0 / 15, g2 = 5.0 / 12, g3 = 3.0 / 4, z2 = -17.0 / 60, z3 = -5.0 / 12;
    const double t0 = h->time, t1 = t0 + dt;
    const double tau1 = t0 + g1 * dt;
    const double tau2 = tau1 + (g2 + z2) * dt;
    const double taus[3] = {t0, tau1, tau2};
    const double gam[3] = {g1, g2, g3}, zet[3] = {0.0, z2, z3};
    const double next_tau[3] = {tau1, tau2, t1};       // the stage that follows (the next step starts at t1)
    for (int s = 0; s < 3; ++s) {
        int rc = refresh_halos(h, h->U[h->cur], h->n_tracers);
        if (rc) return rc;
        rc = launch_stage(h, taus[s], dt, gam[s], zet[s], s == 0, s < 2, 1, next_tau[s]);
        if (rc) return rc;
        h->cur ^= 1;          // the stage kernels (and the slab exchange) leave the halos of the new state valid
    }
    h->time = t1;
    h->iteration += 1;
    return LFB_OK;
}

extern "C" int lfb_run_until(lfb_handle *h, double t_stop, double dt, int n_align, const double *align_intervals)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    if (!(dt > 0)) return fail(LFB_ERR_ARG, "dt must be positive");
    if (n_align < 0 || (n_align > 0 && !align_intervals)) return fail(LFB_ERR_ARG, "bad align intervals");
    while (h->time < t_stop) {
        const double t = h->time;
        double step = dt, target = t + dt;
        bool snapped = false;
        auto consider = [&](double when) {
            const double cand = when - t;
            if (cand < step) { step = cand; target = when; snapped = true; }
        };
        for (int q = 0; q < n_align; ++q) {
            const double iv = align_intervals[q];
            if (!(iv > 0)) return fail(LFB_ERR_ARG, "align interval must be positive");
            consider((floor(t / iv) + 1) * iv);
        }
        consider(t_stop);
        if (!(step > 0)) return fail(LFB_ERR_STATE, "time step alignment produced dt = %g at t = %.17g", step, t);
        int rc = lfb_step(h, step);
        if (rc) return rc;
        // land exactly on the aligned time when t + (when - t) rounds one ulp off
        if (snapped && fabs(h->time - target) <= 4 * fabs(target) * 2.220446049250313e-16) h->time = target;
    }
    return LFB_OK;
}

extern "C" int lfb_get_tendency(lfb_handle *h, int tracer, double *parent)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    if (tracer < 0 || tracer >= h->n_tracers) return fail(LFB_ERR_ARG, "tracer %d out of range", tracer);
    CU(cudaSetDevice(h->device));
    int rc = refresh_halos(h, h->U[h->cur], h->n_tracers);
    if (rc) return rc;
    // AB2 keeps G^- of the previous step in G between steps: set it aside while the diagnostic overwrites G
    double *keep = nullptr;
    const size_t gbytes = (size_t)h->L.len * sizeof(double) * h->n_tracers;
    if (h->desc.timestepper == LFB_TIMESTEPPER_AB2) {
        CU(cudaMalloc(&keep, gbytes));
        CU(cudaMemcpyAsync(keep, h->G, gbytes, cudaMemcpyDeviceToDevice, h->compute));
    }
    rc = launch_stage(h, h->time, 0.0, 0.0, 0.0, 1, 1, 0);
    if (!rc) rc = download(h, h->G + (int64_t)tracer * h->L.len, parent, 0, -1);
    if (keep) {
        cudaMemcpyAsync(h->G, keep, gbytes, cudaMemcpyDeviceToDevice, h->compute);
        cudaStreamSynchronize(h->compute);
        cudaFree(keep);
    }
    return rc;
}

// ---- outputs -------------------------------------------------------------------------------------
// queue the output combination into h->scratch (interior, then the halo pattern of an Oceananigans computed Field)
static int output_compute(lfb_handle *h, int kind, int field)
{
    int fld;
    if (kind == LFB_OUT_FILTERED_VAR) {
        if (field < 0 || field >= h->desc.n_vars) return fail(LFB_ERR_ARG, "variable %d out of range", field);
        fld = field;
    } else if (kind == LFB_OUT_XI || kind == LFB_OUT_MEAN_VEL) {
        if (field < 0 || field >= h->n_maps) return fail(LFB_ERR_ARG, "map %d out of range", field);
        fld = h->desc.n_vars + field;
    } else return fail(LFB_ERR_ARG, "unknown output kind %d", kind);
    LfbOutputParams P;
    memset(&P, 0, sizeof(P));
    P.L = h->L;
    P.U = h->U[h->cur];
    P.n_terms = h->n_pairs;
    P.single_exp = h->single_exp;
    for (int p = 0; p < h->n_pairs; ++p) {
        const double a = h->desc.a[p], b = h->desc.b[p], c = h->desc.c[p], d = h->desc.d[p];
        P.tracer_c[p] = (fld * h->n_pairs + p) * h->sub;
        if (kind == LFB_OUT_MEAN_VEL) {
            if (h->single_exp) { P.wc[p] = -a * c; P.ws[p] = 0; }
            else { P.wc[p] = -a * c + b * d; P.ws[p] = -a * d - b * c; }
        } else { P.wc[p] = a; P.ws[p] = h->single_exp ? 0 : b; }
    }
    CU(cudaMemsetAsync(h->scratch, 0, (size_t)h->L.len * sizeof(double), h->compute));
    CU(lfb_launch_output(h->scratch, &P, h->compute));
    h->launches++;
    int rc = fill_all_halos(h, h->scratch, 1);
    if (rc) return rc;
    return exchange_y(h, h->scratch, 1);
}

static int output_impl(lfb_handle *h, int kind, int field, void *parent, int is_f32)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    int rc = output_compute(h, kind, field);
    if (rc) return rc;
    return download(h, h->scratch, parent, is_f32, -1);
}

extern "C" int lfb_output_async(lfb_handle *h, int kind, int field, void *parent, int is_f32)
{
    if (!h || !parent) return fail(LFB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    const int R = 4;
    if (h->export_ring.empty()) {
        h->export_ring.assign(R, nullptr);
        h->ev_export.resize(R); h->ev_fetched.resize(R); h->ring_used.assign(R, 0);
        for (int r = 0; r < R; ++r) {
            CU(cudaMalloc(&h->export_ring[r], h->staging_bytes));
            CU(cudaEventCreateWithFlags(&h->ev_export[r], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&h->ev_fetched[r], cudaEventDisableTiming));
        }
    }
    const int r = h->export_next++ % R;
    if (h->ring_used[r]) CU(cudaStreamWaitEvent(h->compute, h->ev_fetched[r], 0));   // previous copy out of this buffer
    int rc = output_compute(h, kind, field);
    if (rc) return rc;
    int e[3];
    parent_extents(h, -1, e);
    const size_t n = (size_t)e[0] * e[1] * e[2];
    CU(lfb_launch_export(h->export_ring[r], is_f32 != 0, h->scratch, &h->L, e[0], e[1], e[2], h->compute));
    h->launches++;
    CU(cudaEventRecord(h->ev_export[r], h->compute));
    CU(cudaStreamWaitEvent(h->d2h, h->ev_export[r], 0));
    CU(cudaMemcpyAsync(parent, h->export_ring[r], n * (is_f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, h->d2h));
    CU(cudaEventRecord(h->ev_fetched[r], h->d2h));
    h->ring_used[r] = 1;
    return LFB_OK;
}

extern "C" int lfb_output(lfb_handle *h, int kind, int field, double *parent) { return output_impl(h, kind, field, parent, 0); }
extern "C" int lfb_output_f32(lfb_handle *h, int kind, int field, float *parent) { return output_impl(h, kind, field, parent, 1); }

extern "C" int lfb_combine(lfb_handle *h, int64_t n, const double *fwd, const double *bwd_lo, const double *bwd_hi,
                           double w_hi, double sign, double *out)
{
    if (!h || !fwd || !bwd_lo || !out || n <= 0) return fail(LFB_ERR_ARG, "bad argument");
    if (w_hi != 0.0 && !bwd_hi) return fail(LFB_ERR_ARG, "bwd_hi required when w_hi != 0");
    CU(cudaSetDevice(h->device));
    // chunk through the scratch/staging buffers: [fwd | lo | hi] thirds of staging, result in scratch
    const int64_t cap = (int64_t)(h->staging_bytes / sizeof(double)) / 3;
    if (cap <= 0) return fail(LFB_ERR_STATE, "staging buffer too small");
    CU(cudaStreamSynchronize(h->copy));
    double *d_f = (double *)h->staging, *d_lo = d_f + cap, *d_hi = d_lo + cap;
    for (int64_t off = 0; off < n; off += cap) {
        const int64_t m = n - off < cap ? n - off : cap;
        CU(cudaMemcpyAsync(d_f, fwd + off, m * sizeof(double), cudaMemcpyHostToDevice, h->compute));
        CU(cudaMemcpyAsync(d_lo, bwd_lo + off, m * sizeof(double), cudaMemcpyHostToDevice, h->compute));
        if (w_hi != 0.0) CU(cudaMemcpyAsync(d_hi, bwd_hi + off, m * sizeof(double), cudaMemcpyHostToDevice, h->compute));
        CU(lfb_launch_combine(h->scratch, d_f, d_lo, d_hi, w_hi, sign, m, h->compute));
        h->launches++;
        CU(cudaMemcpyAsync(out + off, h->scratch, m * sizeof(double), cudaMemcpyDeviceToHost, h->compute));
        CU(cudaStreamSynchronize(h->compute));
    }
    // scratch doubles as the zero-initialised work field of outputs: nothing to restore, it is
    // memset before every use
    return LFB_OK;
}


// ---- device-resident pass outputs ------------------------------------------------------------------
static int kept_new(lfb_handle *h, int is_f32, const int e[3], int *id)
{
    const size_t n = (size_t)e[0] * e[1] * e[2];
    const size_t bytes = n * (is_f32 ? sizeof(float) : sizeof(double));
    void *p = nullptr;
    for (size_t q = 0; q < h->kept_pool.size() && !p; ++q)
        if (h->kept_pool[q].bytes == bytes) {
            // stream order protects the readers queued on the compute stream; a download still in flight on the d2h
            // stream is waited for by the compute stream, not by the host
            if (h->kept_pool[q].fetched) {
                CU(cudaStreamWaitEvent(h->compute, h->kept_pool[q].fetched, 0));
                cudaEventDestroy(h->kept_pool[q].fetched);
            }
            p = h->kept_pool[q].p;
            h->kept_pool.erase(h->kept_pool.begin() + q);
        }
    if (!p) {
        cudaError_t err = cudaMalloc(&p, bytes);
        if (err != cudaSuccess && !h->kept_pool.empty()) {          // give the pooled blocks back and retry
            cudaGetLastError();
            cudaDeviceSynchronize();
            for (auto &k : h->kept_pool) { cudaFree(k.p); if (k.fetched) cudaEventDestroy(k.fetched); }
            h->kept_pool.clear();
            err = cudaMalloc(&p, bytes);
        }
        if (err != cudaSuccess) {
            cudaGetLastError();
            return fail(LFB_ERR_CUDA, "device memory exhausted while keeping an output (%s); use the host path", cudaGetErrorString(err));
        }
    }
    lfb_handle::Kept k;
    k.p = p; k.is_f32 = is_f32; k.e[0] = e[0]; k.e[1] = e[1]; k.e[2] = e[2]; k.fetched = nullptr;
    for (size_t q = 0; q < h->kept.size(); ++q)
        if (!h->kept[q].p) { h->kept[q] = k; *id = (int)q; return LFB_OK; }
    h->kept.push_back(k);
    *id = (int)h->kept.size() - 1;
    return LFB_OK;
}

static lfb_handle::Kept *kept_get(lfb_handle *h, int id)
{
    if (!h || id < 0 || id >= (int)h->kept.size() || !h->kept[id].p) return nullptr;
    return &h->kept[id];
}

extern "C" int lfb_output_keep(lfb_handle *h, int kind, int field, int is_f32, int *id)
{
    if (!h || !id) return fail(LFB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    int rc = output_compute(h, kind, field);
    if (rc) return rc;
    int e[3];
    parent_extents(h, -1, e);
    rc = kept_new(h, is_f32 != 0, e, id);
    if (rc) return rc;
    CU(lfb_launch_export(h->kept[*id].p, is_f32 != 0, h->scratch, &h->L, e[0], e[1], e[2], h->compute));
    h->launches++;
    return LFB_OK;
}

extern "C" int lfb_kept_combine(lfb_handle *h, int fwd, int bwd_lo, int bwd_hi, double w_hi, double sign, int *id)
{
    if (!h || !id) return fail(LFB_ERR_ARG, "null argument");
    lfb_handle::Kept *f = kept_get(h, fwd), *lo = kept_get(h, bwd_lo), *hi = w_hi != 0.0 ? kept_get(h, bwd_hi) : lo;
    if (!f || !lo || !hi) return fail(LFB_ERR_ARG, "unknown kept output id");
    for (int a = 0; a < 3; ++a)
        if (f->e[a] != lo->e[a] || f->e[a] != hi->e[a]) return fail(LFB_ERR_ARG, "kept outputs differ in shape");
    CU(cudaSetDevice(h->device));
    const lfb_handle::Kept F = *f, LO = *lo, HI = *hi;            // kept_new may move the vector
    int rc = kept_new(h, 0, F.e, id);
    if (rc) return rc;
    const int64_t n = (int64_t)F.e[0] * F.e[1] * F.e[2];
    CU(lfb_launch_combine_mixed((double *)h->kept[*id].p, F.p, F.is_f32, LO.p, LO.is_f32, HI.p, HI.is_f32, w_hi, sign, n, h->compute));
    h->launches++;
    return LFB_OK;
}

extern "C" int lfb_kept_fetch(lfb_handle *h, int id, void *parent)
{
    lfb_handle::Kept *k = kept_get(h, id);
    if (!k || !parent) return fail(LFB_ERR_ARG, "unknown kept output id / null pointer");
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)k->e[0] * k->e[1] * k->e[2];
    CU(cudaMemcpyAsync(parent, k->p, n * (k->is_f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, h->compute));
    CU(cudaStreamSynchronize(h->compute));
    return LFB_OK;
}

// Asynchronous fetch: queued behind what produces the array, copied on the download stream so that the host can keep
// queuing time steps; `parent` may be read after lfb_wait_transfers() (page-locked memory for a real overlap).
// lfb_kept_release waits for a pending fetch of the array it frees.
extern "C" int lfb_kept_fetch_async(lfb_handle *h, int id, void *parent)
{
    lfb_handle::Kept *k = kept_get(h, id);
    if (!k || !parent) return fail(LFB_ERR_ARG, "unknown kept output id / null pointer");
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)k->e[0] * k->e[1] * k->e[2];
    if (!k->fetched) CU(cudaEventCreateWithFlags(&k->fetched, cudaEventDisableTiming));
    CU(cudaEventRecord(h->ev_tmp, h->compute));
    CU(cudaStreamWaitEvent(h->d2h, h->ev_tmp, 0));
    CU(cudaMemcpyAsync(parent, k->p, n * (k->is_f32 ? sizeof(float) : sizeof(double)), cudaMemcpyDeviceToHost, h->d2h));
    CU(cudaEventRecord(k->fetched, h->d2h));
    return LFB_OK;
}

extern "C" int lfb_kept_info(lfb_handle *h, int id, int *is_f32, int *extents3, void **device_ptr)
{
    lfb_handle::Kept *k = kept_get(h, id);
    if (!k) return fail(LFB_ERR_ARG, "unknown kept output id");
    if (is_f32) *is_f32 = k->is_f32;
    if (extents3) for (int a = 0; a < 3; ++a) extents3[a] = k->e[a];
    if (device_ptr) {
        // the pointer is handed to work outside the handle's streams (lfb_remap_to_mean): finish what produces it
        CU(cudaSetDevice(h->device));
        CU(cudaStreamSynchronize(h->compute));
        *device_ptr = k->p;
    }
    return LFB_OK;
}

extern "C" int lfb_kept_release(lfb_handle *h, int id)
{
    lfb_handle::Kept *k = kept_get(h, id);
    if (!k) return fail(LFB_ERR_ARG, "unknown kept output id");
    // no host synchronisation: the block goes to the pool together with the event of a download that may still be
    // in flight (see kept_new)
    lfb_handle::Pooled pl;
    pl.p = k->p; pl.bytes = (size_t)k->e[0] * k->e[1] * k->e[2] * (k->is_f32 ? sizeof(float) : sizeof(double)); pl.fetched = k->fetched;
    h->kept_pool.push_back(pl);
    k->p = nullptr; k->fetched = nullptr;
    return LFB_OK;
}

// Put `count` blocks of `bytes` bytes into the pool ahead of time, so that kept outputs created between time steps find
// their memory there (a cudaMalloc of a GB-sized block between queued stage kernels stalls the queue for milliseconds).
extern "C" int lfb_kept_reserve(lfb_handle *h, size_t bytes, int count)
{
    if (!h || bytes == 0 || count < 0) return fail(LFB_ERR_ARG, "bad argument");
    CU(cudaSetDevice(h->device));
    int have = 0;                                   // "at least `count`": blocks of this size that are pooled already count
    for (auto &k : h->kept_pool) if (k.bytes == bytes) ++have;
    for (int q = have; q < count; ++q) {
        lfb_handle::Pooled pl;
        pl.p = nullptr; pl.bytes = bytes; pl.fetched = nullptr;
        cudaError_t err = cudaMalloc(&pl.p, bytes);
        if (err != cudaSuccess) { cudaGetLastError(); return fail(LFB_ERR_CUDA, "device memory exhausted while reserving kept outputs (%s)", cudaGetErrorString(err)); }
        h->kept_pool.push_back(pl);
    }
    return LFB_OK;
}

extern "C" int lfb_kept_trim(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->compute));
    CU(cudaStreamSynchronize(h->d2h));
    for (auto &k : h->kept_pool) { cudaFree(k.p); if (k.fetched) cudaEventDestroy(k.fetched); }
    h->kept_pool.clear();
    return LFB_OK;
}

// Gather the y-slabs of a kept output on rank `root` (collective over the slab communicator): the interior rows of
// every rank, the low halo rows of rank 0 and the high halo rows of the last rank -> the GLOBAL parent array.
// temporary device blocks of the gather, taken from / returned to the pool of the kept outputs: everything the gather does is
// ordered on the compute stream, so a block may be handed out again while earlier work on it is still queued there
static void *pool_take(lfb_handle *h, size_t bytes)
{
    for (size_t q = 0; q < h->kept_pool.size(); ++q)
        if (h->kept_pool[q].bytes == bytes) {
            if (h->kept_pool[q].fetched) { cudaStreamWaitEvent(h->compute, h->kept_pool[q].fetched, 0); cudaEventDestroy(h->kept_pool[q].fetched); }
            void *p = h->kept_pool[q].p;
            h->kept_pool.erase(h->kept_pool.begin() + q);
            return p;
        }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

static void pool_give(lfb_handle *h, void *p, size_t bytes)
{
    if (!p) return;
    lfb_handle::Pooled pl;
    pl.p = p; pl.bytes = bytes; pl.fetched = nullptr;
    h->kept_pool.push_back(pl);
}

extern "C" int lfb_kept_gather(lfb_handle *h, int id, int root, int *gid)
{
    if (!h || !gid) return fail(LFB_ERR_ARG, "null argument");
    lfb_handle::Kept *kp = kept_get(h, id);
    if (!kp) return fail(LFB_ERR_ARG, "unknown kept output id");
    if (root < 0 || root >= h->nranks) return fail(LFB_ERR_ARG, "root %d out of range", root);
    *gid = -1;
    if (h->nranks == 1) { *gid = id; return LFB_OK; }
    CU(cudaSetDevice(h->device));
    const lfb_handle::Kept K = *kp;
    const int H = h->L.H[1], ny = h->L.N[1], R = h->nranks;
    const size_t item = K.is_f32 ? sizeof(float) : sizeof(double);
    const int ex = K.e[0], ey = K.e[1], ez = K.e[2];
    const int EY = ny * R + 2 * H;
    auto rows_of = [&](int r, int *lo, int *hi) { *lo = r == 0 ? 0 : H; *hi = r == R - 1 ? ny + 2 * H : ny + H; };
    int lo, hi;
    rows_of(h->rank, &lo, &hi);
    // pack my rows: per z plane a contiguous run of (hi - lo) * ex elements.  No host synchronisation anywhere: the packs,
    // the NCCL transfers and the unpacks are ordered on the compute stream.
    const size_t my_row_bytes = (size_t)(hi - lo) * ex * item;
    void *pack = pool_take(h, my_row_bytes * ez);
    if (!pack) return fail(LFB_ERR_CUDA, "device memory exhausted in the gather of slab outputs");
    CU(cudaMemcpy2DAsync(pack, my_row_bytes, (const char *)K.p + (size_t)lo * ex * item, (size_t)ey * ex * item, my_row_bytes, ez,
                         cudaMemcpyDeviceToDevice, h->compute));
    int rc = LFB_OK;
    if (h->rank == root) {
        int ge[3] = {ex, EY, ez};
        rc = kept_new(h, K.is_f32, ge, gid);
        if (rc) { pool_give(h, pack, my_row_bytes * ez); return rc; }
        char *G = (char *)h->kept[*gid].p;
        std::vector<void *> tmp(R, nullptr);
        std::vector<size_t> tmp_bytes(R, 0);
        for (int r = 0; r < R && !rc; ++r) {
            if (r == root) continue;
            int rlo, rhi;
            rows_of(r, &rlo, &rhi);
            tmp_bytes[r] = (size_t)(rhi - rlo) * ex * item * ez;
            tmp[r] = pool_take(h, tmp_bytes[r]);
            if (!tmp[r]) rc = fail(LFB_ERR_CUDA, "device memory exhausted in the gather of slab outputs");
        }
        if (!rc) {
            ncclResult_t nr = ncclGroupStart();
            for (int r = 0; r < R && nr == ncclSuccess; ++r)
                if (r != root) nr = ncclRecv(tmp[r], tmp_bytes[r], ncclChar, r, h->comm, h->compute);
            if (nr == ncclSuccess) nr = ncclGroupEnd(); else ncclGroupEnd();
            if (nr != ncclSuccess) rc = fail(LFB_ERR_COMM, "gather of slab outputs failed: %s", ncclGetErrorString(nr));
        }
        for (int r = 0; r < R && !rc; ++r) {
            int rlo, rhi;
            rows_of(r, &rlo, &rhi);
            const size_t row_bytes = (size_t)(rhi - rlo) * ex * item;
            const void *src = r == root ? pack : tmp[r];
            cudaError_t ce = cudaMemcpy2DAsync(G + (size_t)(r * ny + rlo) * ex * item, (size_t)EY * ex * item, src, row_bytes, row_bytes, ez,
                                               cudaMemcpyDeviceToDevice, h->compute);
            if (ce != cudaSuccess) rc = fail(LFB_ERR_CUDA, "cudaMemcpy2DAsync: %s", cudaGetErrorString(ce));
        }
        for (int r = 0; r < R; ++r) pool_give(h, tmp[r], tmp_bytes[r]);
    } else {
        ncclResult_t nr = ncclSend(pack, my_row_bytes * ez, ncclChar, root, h->comm, h->compute);
        if (nr != ncclSuccess) rc = fail(LFB_ERR_COMM, "gather of slab outputs failed: %s", ncclGetErrorString(nr));
    }
    pool_give(h, pack, my_row_bytes * ez);
    return rc;
}

extern "C" int lfb_mem_info(lfb_handle *h, size_t *free_bytes, size_t *total_bytes)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return LFB_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------
extern "C" int lfb_comm_unique_id(void *id128)
{
    if (!id128) return fail(LFB_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return LFB_OK;
}

extern "C" int lfb_comm_init(lfb_handle *h, const void *id128, int rank, int nranks)
{
    if (!h || !id128) return fail(LFB_ERR_ARG, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(LFB_ERR_ARG, "bad rank %d / %d", rank, nranks);
    if (h->comm) return fail(LFB_ERR_STATE, "communicator already initialised");
    if (h->L.topo[1] == LFB_FLAT) return fail(LFB_ERR_ARG, "y-slab decomposition needs a non-Flat y axis");
    if (nranks > 1 && h->L.N[1] < h->L.H[1]) return fail(LFB_ERR_ARG, "slab thinner than the halo");
    CU(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NC(ncclCommInitRank(&h->comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks;
    // global indexing of this slab: equal slabs, rank-major (the host guarantees Ny_global = R * Ny_local)
    h->L.NG[1] = h->L.N[1] * nranks;
    h->L.goff[1] = h->L.N[1] * rank;
    h->L.connected_lo[1] = rank > 0 || h->L.topo[1] == LFB_PERIODIC;
    h->L.connected_hi[1] = rank < nranks - 1 || h->L.topo[1] == LFB_PERIODIC;
    if (nranks > 1) {
        h->L.fuse_halo[1] = 0;                        // y halos come from the neighbour ranks
        for (int ax = 0; ax < 3; ax += 2)
            if (h->L.topo[ax] == LFB_PERIODIC && !h->L.fuse_halo[ax])
                return fail(LFB_ERR_ARG, "slab decomposition needs periodic axes at least twice as long as their halo");
        int lo_prio = 0, hi_prio = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CU(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi_prio));
        CU(cudaEventCreateWithFlags(&h->ev_strips, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
    }
    h->halos_valid = 0;
    const size_t per_field = (size_t)(h->L.N[0] + 2 * h->L.H[0]) * h->L.H[1] * (h->L.N[2] + 2 * h->L.H[2]);
    h->halo_elems = per_field * (h->n_tracers > 0 ? h->n_tracers : 1);
    CU(cudaMalloc(&h->send_lo, h->halo_elems * sizeof(double)));
    CU(cudaMalloc(&h->send_hi, h->halo_elems * sizeof(double)));
    CU(cudaMalloc(&h->recv_lo, h->halo_elems * sizeof(double)));
    CU(cudaMalloc(&h->recv_hi, h->halo_elems * sizeof(double)));
    if (nranks > 2) {
        // NCCL sets up a point-to-point connection the first time two ranks talk: the halo exchange only ever connects
        // neighbours, the gather of the outputs (lfb_kept_gather) any pair.  One tiny all-to-all here moves those set-ups
        // (tens of milliseconds each) out of the first pipeline run.
        double *w = nullptr;
        CU(cudaMalloc(&w, 2 * (size_t)nranks * sizeof(double)));
        CU(cudaMemset(w, 0, 2 * (size_t)nranks * sizeof(double)));
        NC(ncclGroupStart());
        for (int r = 0; r < nranks; ++r) {
            if (r == rank) continue;
            NC(ncclSend(w + r, 1, ncclDouble, r, h->comm, h->compute));
            NC(ncclRecv(w + nranks + r, 1, ncclDouble, r, h->comm, h->compute));
        }
        NC(ncclGroupEnd());
        CU(cudaStreamSynchronize(h->compute));
        cudaFree(w);
    }
    int rc = select_kernel(h);
    if (rc) return rc;
    if (h->imm_set) {
        if (h->use_zmarch == 3 && h->dzc) { h->use_zmarch = 0; h->fold = 0; }
        rc = build_face_codes(h);
    }
    return rc;
}

// ---- introspection -------------------------------------------------------------------------------
extern "C" int64_t lfb_kernel_launches(const lfb_handle *h) { return h ? h->launches : -1; }

extern "C" int lfb_stage_time_ms(lfb_handle *h, int reset, double *ms, int64_t *launches)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    int rc = drain_timings(h, true);
    if (rc) return rc;
    if (ms) *ms = h->stage_ms;
    if (launches) *launches = h->stage_launches;
    if (reset) { h->stage_ms = 0; h->stage_launches = 0; }
    return LFB_OK;
}

extern "C" int lfb_timer_start(lfb_handle *h)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev_t0, h->compute));
    return LFB_OK;
}

extern "C" int lfb_timer_stop(lfb_handle *h, double *ms)
{
    if (!h || !ms) return fail(LFB_ERR_ARG, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev_t1, h->compute));
    CU(cudaEventSynchronize(h->ev_t1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, h->ev_t0, h->ev_t1));
    *ms = f;
    return LFB_OK;
}

extern "C" const char *lfb_kernel_name(const lfb_handle *h)
{
    if (!h) return "";
    if (h->use_zmarch == 3) return "ztma";
    return h->desc.strict ? "generic-strict" : "generic";
}

extern "C" void *lfb_device_ptr(lfb_handle *h, int what, int idx)
{
    if (!h) return nullptr;
    if (what == 0 && idx >= 0 && idx < h->n_tracers) return h->U[h->cur] + (int64_t)idx * h->L.len;
    if (what == 1 && idx >= 0 && idx < h->n_tracers) return h->G + (int64_t)idx * h->L.len;
    if (what == 2) {
        const int slot = idx / 16, f = idx % 16;
        if (slot >= 0 && slot < h->desc.n_slots && f < N_FIELD_IDS) return h->slot_field[(size_t)slot * N_FIELD_IDS + f];
    }
    return nullptr;
}

extern "C" int lfb_layout(const lfb_handle *h, int64_t *x_offset, int64_t *pitch, int64_t *rows, int64_t *planes)
{
    if (!h) return fail(LFB_ERR_ARG, "null handle");
    if (x_offset) *x_offset = h->L.xoff;
    if (pitch) *pitch = h->L.pitch;
    if (rows) *rows = h->L.rows;
    if (planes) *planes = h->L.planes;
    return LFB_OK;
}
