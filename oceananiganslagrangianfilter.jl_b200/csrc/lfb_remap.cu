// lfb_remap.cu -- remap of filtered fields to the mean position for grids with two non-Flat axes
// (reference regrid_to_mean_position!, src/Utils/post_processing_utils.jl:327-762).
//
// The reference interpolates the scattered data {(x + xi, f*)} to the grid nodes with
// scipy.interpolate.LinearNDInterpolator, i.e. a Delaunay triangulation (Qhull) of the normalised,
// periodically padded point cloud followed by barycentric interpolation, rebuilt per variable and per
// output time.  Here no global triangulation is built.  The Delaunay triangulation of points in general
// position is unique, so the triangle that contains a node can be constructed locally: starting from a
// Delaunay edge (nearest data point a and its own nearest neighbour b), the Delaunay triangle on the
// node's side of the edge is (a, b, c) with c the point that sees ab under the largest angle; if the node
// lies beyond one of the two new edges the walk continues from that edge.  Every step is one pass over a
// patch of lattice neighbours of the node, all nodes are independent (one thread each), and the result
// is shared by all variables of an output time.  The circumcircle of the triangle found is checked
// against the patch, so a patch that was too small is detected and the caller retries with a larger one.
// The bounded-axis edge rows are then overwritten by the reference's 1-D interpolation along the row
// (numpy.interp, post:641-745) and the halos zeroed (post:748).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "lfb200.h"

struct RemapParams {
    int    n[2];            // interior points along the two true axes
    int    H[2];            // halos along them
    int    e[2];            // parent extents n + 2H
    int    periodic[2];
    int    has_map[2];
    double origin[2], delta[2], L[2];
    double mn[2], rg[2];    // normalisation of the padded cloud
    double pad[2];          // width npad * delta of the periodic padding (post:527-611)
    int    r[2];            // patch radius in cells
    double maxdisp[2];      // max |xi| in cells (for the patch sufficiency check)
    int64_t stride[2];      // element strides of the two axes in the parent arrays
};

__device__ __forceinline__ double orient2d(double ax, double ay, double bx, double by, double cx, double cy)
{
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
}

struct Cand { double x, y; int64_t idx; };

// Position along axis q of the periodic image of lattice point `qi` that sits at lattice offset `w` periods from it,
// or false if the reference's padded cloud does not hold that image: the cloud has the point wrapped into
// [lo, lo + L) (post:427-429), its copy one period down if the wrapped point lies within `pad` of the upper end and
// its copy one period up if it lies within `pad` of the lower end (post:527-611) -- nothing further out, which
// changes the triangulation near periodic edges when the normalised cells are strongly anisotropic.
__device__ __forceinline__ bool image_position(const RemapParams &P, int q, int qi, int w, double disp, double &x)
{
    const double lo = P.origin[q];
    const double X = lo + P.delta[q] * (qi + 0.5) + disp;
    if (!P.periodic[q]) { x = X; return true; }
    const double fl = floor((X - lo) / P.L[q]);
    const double Xw = X - fl * P.L[q];
    const int m = w + (int)fl;                 // image index relative to the wrapped point
    if (m == 0) { x = Xw; return true; }
    if (m == -1 && Xw > lo + P.L[q] - P.pad[q] && Xw < lo + P.L[q]) { x = Xw - P.L[q]; return true; }
    if (m == 1 && Xw < lo + P.pad[q] && Xw > lo) { x = Xw + P.L[q]; return true; }
    return false;
}

// lattice neighbour (p0, p1) of the patch: normalised position of its (periodic image of the) data point
__device__ __forceinline__ bool candidate(const RemapParams &P, const double *__restrict__ xi0, const double *__restrict__ xi1,
                                          int p0, int p1, Cand &c)
{
    int q0 = p0, q1 = p1, w0 = 0, w1 = 0;
    if (P.periodic[0]) { while (q0 < 0) { q0 += P.n[0]; --w0; } while (q0 >= P.n[0]) { q0 -= P.n[0]; ++w0; } }
    else if (p0 < 0 || p0 >= P.n[0]) return false;
    if (P.periodic[1]) { while (q1 < 0) { q1 += P.n[1]; --w1; } while (q1 >= P.n[1]) { q1 -= P.n[1]; ++w1; } }
    else if (p1 < 0 || p1 >= P.n[1]) return false;
    const int64_t idx = (int64_t)(q0 + P.H[0]) * P.stride[0] + (int64_t)(q1 + P.H[1]) * P.stride[1];
    double x0, x1;
    if (!image_position(P, 0, q0, w0, P.has_map[0] ? xi0[idx] : 0.0, x0)) return false;
    if (!image_position(P, 1, q1, w1, P.has_map[1] ? xi1[idx] : 0.0, x1)) return false;
    c.x = (x0 - P.mn[0]) / P.rg[0];
    c.y = (x1 - P.mn[1]) / P.rg[1];
    c.idx = idx;
    return true;
}

// one thread per interior node: Delaunay triangle containing the node + barycentric weights
__global__ void lfb_remap2d_locate(const RemapParams P, const double *__restrict__ xi0, const double *__restrict__ xi1,
                                   int64_t *__restrict__ tri, double *__restrict__ wgt, int *__restrict__ too_small)
{
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y;
    if (i0 >= P.n[0]) return;
    const int64_t node = (int64_t)i1 * P.n[0] + i0;
    const double qx = (P.origin[0] + P.delta[0] * (i0 + 0.5) - P.mn[0]) / P.rg[0];
    const double qy = (P.origin[1] + P.delta[1] * (i1 + 0.5) - P.mn[1]) / P.rg[1];
    const int r0 = P.r[0], r1 = P.r[1];
    Cand a, b, c, t;
    int pa0 = 0, pa1 = 0, pb0 = 0, pb1 = 0, pc0 = 0, pc1 = 0;
    // a: nearest data point to the node
    double best = INFINITY;
    bool have = false;
    for (int d1 = -r1; d1 <= r1; ++d1)
        for (int d0 = -r0; d0 <= r0; ++d0) {
            if (!candidate(P, xi0, xi1, i0 + d0, i1 + d1, t)) continue;
            const double dd = (t.x - qx) * (t.x - qx) + (t.y - qy) * (t.y - qy);
            if (dd < best) { best = dd; a = t; pa0 = i0 + d0; pa1 = i1 + d1; have = true; }
        }
    // b: nearest neighbour of a (a Delaunay edge)
    best = INFINITY;
    bool have_b = false;
    for (int d1 = -r1; d1 <= r1; ++d1)
        for (int d0 = -r0; d0 <= r0; ++d0) {
            const int p0 = i0 + d0, p1 = i1 + d1;
            if ((p0 == pa0 && p1 == pa1) || !candidate(P, xi0, xi1, p0, p1, t)) continue;
            const double dd = (t.x - a.x) * (t.x - a.x) + (t.y - a.y) * (t.y - a.y);
            if (dd < best) { best = dd; b = t; pb0 = p0; pb1 = p1; have_b = true; }
        }
    bool ok = have && have_b;
    if (ok && orient2d(a.x, a.y, b.x, b.y, qx, qy) < 0) { t = a; a = b; b = t; int s = pa0; pa0 = pb0; pb0 = s; s = pa1; pa1 = pb1; pb1 = s; }
    bool found = false;
    for (int step = 0; ok && step < 64; ++step) {
        // c: the point to the left of a->b that sees ab under the largest angle (smallest cotangent)
        double bd = 0, bc = -1;
        bool any = false;
        for (int d1 = -r1; d1 <= r1; ++d1)
            for (int d0 = -r0; d0 <= r0; ++d0) {
                const int p0 = i0 + d0, p1 = i1 + d1;
                if ((p0 == pa0 && p1 == pa1) || (p0 == pb0 && p1 == pb1)) continue;
                if (!candidate(P, xi0, xi1, p0, p1, t)) continue;
                const double cr = (a.x - t.x) * (b.y - t.y) - (a.y - t.y) * (b.x - t.x);
                if (!(cr > 0)) continue;
                const double dt = (a.x - t.x) * (b.x - t.x) + (a.y - t.y) * (b.y - t.y);
                if (!any || dt * bc < bd * cr) { bd = dt; bc = cr; c = t; pc0 = p0; pc1 = p1; any = true; }
            }
        if (!any) { ok = false; break; }                 // the node is outside the convex hull
        const double o1 = orient2d(b.x, b.y, c.x, c.y, qx, qy), o2 = orient2d(c.x, c.y, a.x, a.y, qx, qy);
        if (o1 >= 0 && o2 >= 0) { found = true; break; }
        if (o1 < 0 && (o1 <= o2 || o2 >= 0)) { a = c; pa0 = pc0; pa1 = pc1; }      // leave through edge b-c: new edge c->b
        else                                 { b = c; pb0 = pc0; pb1 = pc1; }      // leave through edge c-a: new edge a->c
    }
    if (!found) { tri[3 * node] = -1; tri[3 * node + 1] = -1; tri[3 * node + 2] = -1; wgt[3 * node] = wgt[3 * node + 1] = wgt[3 * node + 2] = 0; return; }
    const double area = orient2d(a.x, a.y, b.x, b.y, c.x, c.y);
    const double wa = orient2d(b.x, b.y, c.x, c.y, qx, qy) / area, wb = orient2d(c.x, c.y, a.x, a.y, qx, qy) / area;
    tri[3 * node] = a.idx; tri[3 * node + 1] = b.idx; tri[3 * node + 2] = c.idx;
    wgt[3 * node] = wa; wgt[3 * node + 1] = wb; wgt[3 * node + 2] = 1.0 - wa - wb;
    // patch sufficiency: the circumcircle must stay inside the part of the plane the patch is guaranteed to cover
    const double bx = b.x - a.x, by = b.y - a.y, cx = c.x - a.x, cy = c.y - a.y;
    const double dd = 2.0 * (bx * cy - by * cx);
    const double ux = (cy * (bx * bx + by * by) - by * (cx * cx + cy * cy)) / dd, uy = (bx * (cx * cx + cy * cy) - cx * (bx * bx + by * by)) / dd;
    const double rad = sqrt(ux * ux + uy * uy);
    const double ccx = a.x + ux, ccy = a.y + uy;
    const double h0 = P.delta[0] / P.rg[0], h1 = P.delta[1] / P.rg[1];
    const double safe0 = (r0 - P.maxdisp[0] - 0.5) * h0, safe1 = (r1 - P.maxdisp[1] - 0.5) * h1;
    const bool edge0_lo = !P.periodic[0] && i0 - r0 < 0, edge0_hi = !P.periodic[0] && i0 + r0 >= P.n[0];
    const bool edge1_lo = !P.periodic[1] && i1 - r1 < 0, edge1_hi = !P.periodic[1] && i1 + r1 >= P.n[1];
    // nodes on the first / last row of a Bounded axis are overwritten by the 1-D edge interpolation; their hull
    // triangles can be arbitrarily flat, so they are not checked
    if ((!P.periodic[0] && (i0 == 0 || i0 == P.n[0] - 1)) || (!P.periodic[1] && (i1 == 0 || i1 == P.n[1] - 1))) return;
    bool bad = false;
    if (!edge0_lo && ccx - rad < qx - safe0) bad = true;
    if (!edge0_hi && ccx + rad > qx + safe0) bad = true;
    if (!edge1_lo && ccy - rad < qy - safe1) bad = true;
    if (!edge1_hi && ccy + rad > qy + safe1) bad = true;
    if (bad) atomicAdd(too_small, 1);
}

// out(node) = sum_i w_i f(tri_i) for one variable; halos zero; NaN outside the hull
__global__ void lfb_remap2d_apply(const RemapParams P, const int64_t *__restrict__ tri, const double *__restrict__ wgt,
                                  const double *__restrict__ f, double *__restrict__ out)
{
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y;
    if (i0 >= P.n[0]) return;
    const int64_t node = (int64_t)i1 * P.n[0] + i0;
    const int64_t o = (int64_t)(i0 + P.H[0]) * P.stride[0] + (int64_t)(i1 + P.H[1]) * P.stride[1];
    if (tri[3 * node] < 0) { out[o] = nan(""); return; }
    out[o] = wgt[3 * node] * f[tri[3 * node]] + wgt[3 * node + 1] * f[tri[3 * node + 1]] + wgt[3 * node + 2] * f[tri[3 * node + 2]];
}

// Bounded-axis edge fix (post:641-745, two true dimensions): along the row at interior index `edge` of the
// bounded axis `bd`, numpy.interp of the row's own data points (un-normalised coordinate of the other axis,
// periodic images included) evaluated at the row's nodes.
__global__ void lfb_remap2d_edge(const RemapParams P, const double *__restrict__ xi_other, const double *__restrict__ f,
                                 double *__restrict__ out, int bd, int edge)
{
    const int od = 1 - bd;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n[od]) return;
    const double x = P.origin[od] + P.delta[od] * (i + 0.5);
    const int r = P.r[od] + 1;
    double lo_x = -INFINITY, lo_v = 0, hi_x = INFINITY, hi_v = 0;
    for (int d = -r; d <= r; ++d) {
        int p = i + d;
        double s = 0;
        int q = p;
        if (P.periodic[od]) { int w = 0; while (q < 0) { q += P.n[od]; --w; } while (q >= P.n[od]) { q -= P.n[od]; ++w; } s = w * P.L[od]; }
        else if (p < 0 || p >= P.n[od]) continue;
        const int64_t idx = (int64_t)(q + P.H[od]) * P.stride[od] + (int64_t)(edge + P.H[bd]) * P.stride[bd];
        const double xp = P.origin[od] + P.delta[od] * (q + 0.5) + (P.has_map[od] ? xi_other[idx] : 0.0) + s;
        const double v = f[idx];
        if (xp <= x) { if (xp > lo_x) { lo_x = xp; lo_v = v; } }
        else         { if (xp < hi_x) { hi_x = xp; hi_v = v; } }
    }
    double res;
    if (lo_x == -INFINITY) res = hi_v;                      // numpy.interp clamps outside the data range
    else if (hi_x == INFINITY) res = lo_v;
    else { const double slope = (hi_v - lo_v) / (hi_x - lo_x); res = slope * (x - lo_x) + lo_v; }
    out[(int64_t)(i + P.H[od]) * P.stride[od] + (int64_t)(edge + P.H[bd]) * P.stride[bd]] = res;
}

extern "C" void lfb_internal_set_error(const char *msg);      // lfb_api.cu: message returned by lfb_last_error()
#define RFAIL(code, ...) do { char m_[256]; snprintf(m_, sizeof(m_), __VA_ARGS__); lfb_internal_set_error(m_); return code; } while (0)
#define RCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) RFAIL(LFB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

extern "C" int lfb_remap_to_mean(const lfb_remap_desc *D, const double *const *xi, int n_fields, const double *const *fields,
                                 double *const *out)
{
    if (!D || !xi || !fields || !out || n_fields < 1) RFAIL(LFB_ERR_ARG, "bad argument");
    int ax[3], nt = 0;
    for (int a = 0; a < 3; ++a) if (D->topology[a] != LFB_FLAT) ax[nt++] = a;
    if (nt != 2) RFAIL(LFB_ERR_ARG, "the device remap handles grids with exactly two non-Flat axes (got %d)", nt);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) RFAIL(LFB_ERR_CUDA, "no CUDA device available; liblfb200 has no CPU fallback");
    RCU(cudaSetDevice(D->device));
    RemapParams P;
    memset(&P, 0, sizeof(P));
    int64_t ext[3], st[3];
    for (int a = 0; a < 3; ++a) ext[a] = D->size[a] + 2 * D->halo[a];
    st[0] = 1; st[1] = ext[0]; st[2] = ext[0] * ext[1];
    const int64_t len = ext[0] * ext[1] * ext[2];
    for (int q = 0; q < 2; ++q) {
        const int a = ax[q];
        P.n[q] = D->size[a]; P.H[q] = D->halo[a]; P.e[q] = (int)ext[a];
        P.periodic[q] = D->topology[a] == LFB_PERIODIC;
        P.has_map[q] = D->has_map[a] && xi[a];
        P.origin[q] = D->origin[a]; P.delta[q] = D->spacing[a]; P.L[q] = D->spacing[a] * D->size[a];
        P.stride[q] = st[a];
    }
    // normalisation extents of the padded cloud (post:527-628) and the largest displacements
    for (int q = 0; q < 2; ++q) {
        const int a = ax[q], o = ax[1 - q];
        double mn = INFINITY, mx = -INFINITY, md = 0;
        const double lo = P.origin[q], hi = P.origin[q] + P.L[q], pad = D->npad * P.delta[q];
        for (int j = 0; j < D->size[o]; ++j)
            for (int i = 0; i < D->size[a]; ++i) {
                const int64_t idx = (int64_t)(i + D->halo[a]) * st[a] + (int64_t)(j + D->halo[o]) * st[o];
                const double d = P.has_map[q] ? xi[a][idx] : 0.0;
                double X = P.origin[q] + P.delta[q] * (i + 0.5) + d;
                if (fabs(d) / P.delta[q] > md) md = fabs(d) / P.delta[q];
                if (P.periodic[q] && P.has_map[q]) X -= floor((X - lo) / P.L[q]) * P.L[q];
                if (X < mn) mn = X;
                if (X > mx) mx = X;
                if (P.periodic[q]) {
                    if (X > hi - pad && X < hi) { if (X - P.L[q] < mn) mn = X - P.L[q]; }
                    if (X < lo + pad && X > lo) { if (X + P.L[q] > mx) mx = X + P.L[q]; }
                }
            }
        if (!(md < 1e300)) RFAIL(LFB_ERR_ARG, "non-finite displacement");
        P.mn[q] = mn; P.rg[q] = mx - mn; P.maxdisp[q] = md; P.pad[q] = pad;
    }
    double *d_xi[2] = {nullptr, nullptr}, *d_f = nullptr, *d_out = nullptr, *d_w = nullptr;
    int64_t *d_tri = nullptr;
    int *d_flag = nullptr;
    const int64_t nodes = (int64_t)P.n[0] * P.n[1];
    auto cleanup = [&]() { cudaFree(d_xi[0]); cudaFree(d_xi[1]); cudaFree(d_f); cudaFree(d_out); cudaFree(d_w); cudaFree(d_tri); cudaFree(d_flag); };
    for (int q = 0; q < 2; ++q)
        if (P.has_map[q]) {
            RCU(cudaMalloc(&d_xi[q], len * sizeof(double)));
            RCU(cudaMemcpy(d_xi[q], xi[ax[q]], len * sizeof(double), cudaMemcpyHostToDevice));
        }
    RCU(cudaMalloc(&d_f, len * sizeof(double)));
    RCU(cudaMalloc(&d_out, len * sizeof(double)));
    RCU(cudaMalloc(&d_w, 3 * nodes * sizeof(double)));
    RCU(cudaMalloc(&d_tri, 3 * nodes * sizeof(int64_t)));
    RCU(cudaMalloc(&d_flag, sizeof(int)));
    // patch radius: displacement + anisotropy of the normalised cells + margin; doubled while the check fails
    const double h0 = P.delta[0] / P.rg[0], h1 = P.delta[1] / P.rg[1];
    int r[2];
    r[0] = (int)ceil(P.maxdisp[0]) + 3 + (h0 < h1 ? (int)fmin(ceil(h1 / h0) - 1, 12.0) : 0);
    r[1] = (int)ceil(P.maxdisp[1]) + 3 + (h1 < h0 ? (int)fmin(ceil(h0 / h1) - 1, 12.0) : 0);
    if (D->radius > 0) r[0] = r[1] = D->radius;
    dim3 block(64, 1, 1), grid((P.n[0] + 63) / 64, P.n[1], 1);
    int flag = 1;
    for (int attempt = 0; attempt < 4 && flag; ++attempt) {
        P.r[0] = r[0]; P.r[1] = r[1];
        RCU(cudaMemset(d_flag, 0, sizeof(int)));
        lfb_remap2d_locate<<<grid, block>>>(P, d_xi[0], d_xi[1], d_tri, d_w, d_flag);
        RCU(cudaGetLastError());
        RCU(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) { r[0] = r[0] * 2; r[1] = r[1] * 2; }
        if (r[0] > P.n[0]) r[0] = P.n[0];
        if (r[1] > P.n[1]) r[1] = P.n[1];
    }
    if (flag) { cleanup(); RFAIL(LFB_ERR_STATE, "remap patch too small even at radius (%d, %d): displacements too large for the local search", P.r[0], P.r[1]); }
    for (int v = 0; v < n_fields; ++v) {
        RCU(cudaMemcpy(d_f, fields[v], len * sizeof(double), cudaMemcpyHostToDevice));
        RCU(cudaMemset(d_out, 0, len * sizeof(double)));                       // halos stay zero (post:748)
        lfb_remap2d_apply<<<grid, block>>>(P, d_tri, d_w, d_f, d_out);
        RCU(cudaGetLastError());
        for (int q = 0; q < 2; ++q) {                                            // bounded dims in axis order (post:641)
            if (D->topology[ax[q]] != LFB_BOUNDED) continue;
            const int od = 1 - q;
            for (int edge : {0, P.n[q] - 1}) {
                lfb_remap2d_edge<<<(P.n[od] + 63) / 64, 64>>>(P, d_xi[od], d_f, d_out, q, edge);
                RCU(cudaGetLastError());
            }
        }
        RCU(cudaMemcpy(out[v], d_out, len * sizeof(double), cudaMemcpyDeviceToHost));
    }
    cleanup();
    return LFB_OK;
}
