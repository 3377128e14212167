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
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
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
    // one copy per point: the reference shifts its copy down first, and a copy that was shifted down is not shifted up
    // again (post:537-541), so on an axis shorter than 2 npad cells the points within `pad` of both ends have no upper copy
    if (m == 1 && Xw < lo + P.pad[q] && Xw > lo && !(Xw > lo + P.L[q] - P.pad[q])) { x = Xw + P.L[q]; return true; }
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

// ======================================================================================================================
// Three non-Flat axes.  Same idea one dimension up: the Delaunay tetrahedron that contains a node is built locally.
// The nearest data point a and its nearest neighbour b form a Delaunay edge; the point c with the smallest circumcircle
// through a, b makes (a, b, c) a Delaunay face (the sphere having that circle as a great circle is empty: any point
// inside it would give a smaller circle through a, b); the Delaunay tetrahedron on the node's side of a face (a, b, c)
// is completed by the point d whose circumsphere centre lies lowest along the face normal; if the node is outside
// (a, b, c, d) the walk continues through the face with the most negative barycentric coordinate.  The circumsphere of
// the tetrahedron found is checked against the patch of lattice neighbours that was searched.
struct RemapAxis {
    int     n, H, periodic, has_map, r;
    double  origin, delta, L, mn, rg, pad, maxdisp;
    int64_t stride;
};
struct Remap3Params { RemapAxis ax[3]; };

// as image_position, for one RemapAxis (reference padding rule: post:427-429, 527-611)
__device__ __forceinline__ bool axis_image(const RemapAxis &A, int qi, int w, double disp, double &x)
{
    const double lo = A.origin;
    const double X = lo + A.delta * (qi + 0.5) + disp;
    if (!A.periodic) { x = X; return true; }
    const double fl = floor((X - lo) / A.L);
    const double Xw = X - fl * A.L;
    const int m = w + (int)fl;
    if (m == 0) { x = Xw; return true; }
    if (m == -1 && Xw > lo + A.L - A.pad && Xw < lo + A.L) { x = Xw - A.L; return true; }
    if (m == 1 && Xw < lo + A.pad && Xw > lo && !(Xw > lo + A.L - A.pad)) { x = Xw + A.L; return true; }    // see image_position
    return false;
}

struct Cand3 { double x, y, z; int64_t idx; int p0, p1, p2; };

__device__ __forceinline__ bool candidate3(const Remap3Params &P, const double *__restrict__ xi0, const double *__restrict__ xi1,
                                           const double *__restrict__ xi2, int p0, int p1, int p2, Cand3 &c)
{
    int q[3] = {p0, p1, p2}, w[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const RemapAxis &A = P.ax[a];
        if (A.periodic) { while (q[a] < 0) { q[a] += A.n; --w[a]; } while (q[a] >= A.n) { q[a] -= A.n; ++w[a]; } }
        else if (q[a] < 0 || q[a] >= A.n) return false;
    }
    const int64_t idx = (int64_t)(q[0] + P.ax[0].H) * P.ax[0].stride + (int64_t)(q[1] + P.ax[1].H) * P.ax[1].stride +
                        (int64_t)(q[2] + P.ax[2].H) * P.ax[2].stride;
    double x0, x1, x2;
    if (!axis_image(P.ax[0], q[0], w[0], P.ax[0].has_map ? xi0[idx] : 0.0, x0)) return false;
    if (!axis_image(P.ax[1], q[1], w[1], P.ax[1].has_map ? xi1[idx] : 0.0, x1)) return false;
    if (!axis_image(P.ax[2], q[2], w[2], P.ax[2].has_map ? xi2[idx] : 0.0, x2)) return false;
    c.x = (x0 - P.ax[0].mn) / P.ax[0].rg;
    c.y = (x1 - P.ax[1].mn) / P.ax[1].rg;
    c.z = (x2 - P.ax[2].mn) / P.ax[2].rg;
    c.idx = idx; c.p0 = p0; c.p1 = p1; c.p2 = p2;
    return true;
}

__device__ __forceinline__ bool same3(const Cand3 &a, int p0, int p1, int p2) { return a.p0 == p0 && a.p1 == p1 && a.p2 == p2; }

// 6 x signed volume of (a, b, c, d): positive when d lies on the side of plane (a, b, c) its normal (b-a) x (c-a) points to
__device__ __forceinline__ double orient3(const Cand3 &a, const Cand3 &b, const Cand3 &c, double dx, double dy, double dz)
{
    const double bx = b.x - a.x, by = b.y - a.y, bz = b.z - a.z, cx = c.x - a.x, cy = c.y - a.y, cz = c.z - a.z;
    const double ex = dx - a.x, ey = dy - a.y, ez = dz - a.z;
    return (by * cz - bz * cy) * ex + (bz * cx - bx * cz) * ey + (bx * cy - by * cx) * ez;
}

// Delaunay tetrahedron containing node (i0, i1, i2) + barycentric weights.  `src(p0, p1, p2, t)` delivers the data point
// at lattice offset (p0, p1, p2) (false if the reference's cloud does not hold that image): from global memory, or from
// the patch a block has staged in shared memory.
// slk[2 a], slk[2 a + 1]: bounds of (position - lattice position) along axis a, in normalised units, over every data point the
// source can deliver to this node (the staged patch of the block: the displacement maps are smooth, so the local range is
// a fraction of the global +-maxdisp cells; the global bounds otherwise)
template <class Src>
__device__ __forceinline__ void remap3_locate_node(const Remap3Params &P, const Src &src, int i0, int i1, int i2,
                                                   int64_t *__restrict__ tet, double *__restrict__ wgt, int *__restrict__ too_small,
                                                   const double *slk)
{
    const int64_t node = ((int64_t)i2 * P.ax[1].n + i1) * P.ax[0].n + i0;
    const double qx = (P.ax[0].origin + P.ax[0].delta * (i0 + 0.5) - P.ax[0].mn) / P.ax[0].rg;
    const double qy = (P.ax[1].origin + P.ax[1].delta * (i1 + 0.5) - P.ax[1].mn) / P.ax[1].rg;
    const double qz = (P.ax[2].origin + P.ax[2].delta * (i2 + 0.5) - P.ax[2].mn) / P.ax[2].rg;
    const int r0 = P.ax[0].r, r1 = P.ax[1].r, r2 = P.ax[2].r;
    Cand3 a, b, c, d, t;
    auto none = [&]() {
        for (int v = 0; v < 4; ++v) { tet[4 * node + v] = -1; wgt[4 * node + v] = 0.0; }
    };
    // Nodes on the first / last plane of a Bounded axis take their value from the edge-plane interpolation (post:641-745),
    // which overwrites whatever is found here -- and they lie outside the hull of the cloud, where the search would
    // run through every pass up to the full patch before giving up (1.6 % of the nodes, a quarter of the work).
    if ((!P.ax[0].periodic && (i0 == 0 || i0 == P.ax[0].n - 1)) || (!P.ax[1].periodic && (i1 == 0 || i1 == P.ax[1].n - 1)) ||
        (!P.ax[2].periodic && (i2 == 0 || i2 == P.ax[2].n - 1))) { none(); return; }
    bool found = false;
    double wa = 0, wb = 0, wc = 0, wd = 0, scx = 0, scy = 0, scz = 0, srad2 = 0;
    // Up to four passes over growing patches.  The early ones search an inner patch (a fraction of the candidates) and
    // then check that no point of the shell around it lies inside the circumsphere of the tetrahedron found: the Delaunay
    // tetrahedron that contains the node is the unique one with an empty circumsphere, so a result that passes is the
    // result of the full search.  Otherwise (rare) the next pass searches the next larger patch, the last one the full patch.
    // second patch: displacement + 1 cell; third: half way to the full patch
    const int f0 = min(r0, (int)ceil(P.ax[0].maxdisp) + 1), f1 = min(r1, (int)ceil(P.ax[1].maxdisp) + 1), f2 = min(r2, (int)ceil(P.ax[2].maxdisp) + 1);
    // innermost patch: the displacement alone (at least one cell).  With displacements below one cell every vertex of the
    // enclosing tetrahedron is a direct lattice neighbour of the node (27 candidates instead of 125); the emptiness check
    // over the shell decides, as for every patch but the full one.
    const int g0 = min(f0, max(1, (int)ceil(P.ax[0].maxdisp))), g1 = min(f1, max(1, (int)ceil(P.ax[1].maxdisp))), g2 = min(f2, max(1, (int)ceil(P.ax[2].maxdisp)));
    int pR0 = -1, pR1 = -1, pR2 = -1;
    for (int pass = 0; pass < 4 && !found; ++pass) {
        const int R0 = pass == 0 ? g0 : pass == 1 ? f0 : pass == 2 ? (f0 + r0 + 1) / 2 : r0;
        const int R1 = pass == 0 ? g1 : pass == 1 ? f1 : pass == 2 ? (f1 + r1 + 1) / 2 : r1;
        const int R2 = pass == 0 ? g2 : pass == 1 ? f2 : pass == 2 ? (f2 + r2 + 1) / 2 : r2;
        if (R0 == pR0 && R1 == pR1 && R2 == pR2) continue;                                 // the same patch as the pass before
        pR0 = R0; pR1 = R1; pR2 = R2;
        const bool full_patch = R0 == r0 && R1 == r1 && R2 == r2;
        // a: the node's own data point (it lies within its displacement of the node, and its nearest neighbour -- the
        // Delaunay edge the walk starts from -- is a direct lattice neighbour, inside the innermost patch; starting from
        // the data point nearest to the node instead made that patch fail on anisotropic grids, where the nearest point
        // is often a neighbour in x whose own nearest neighbour lies two cells from the node).  Any vertex will do: the
        // walk ends in the tetrahedron that contains the node.
        double best = INFINITY;
        bool have = src(i0, i1, i2, a);
        if (!have) {
            for (int d2 = -R2; d2 <= R2; ++d2)
                for (int d1 = -R1; d1 <= R1; ++d1)
                    for (int d0 = -R0; d0 <= R0; ++d0) {
                        if (!src(i0 + d0, i1 + d1, i2 + d2, t)) continue;
                        const double dd = (t.x - qx) * (t.x - qx) + (t.y - qy) * (t.y - qy) + (t.z - qz) * (t.z - qz);
                        if (dd < best) { best = dd; a = t; have = true; }
                    }
        }
        if (!have) continue;
        // b: nearest neighbour of a (a Delaunay edge)
        best = INFINITY; have = false;
        for (int d2 = -R2; d2 <= R2; ++d2)
            for (int d1 = -R1; d1 <= R1; ++d1)
                for (int d0 = -R0; d0 <= R0; ++d0) {
                    const int p0 = i0 + d0, p1 = i1 + d1, p2 = i2 + d2;
                    if (same3(a, p0, p1, p2) || !src(p0, p1, p2, t)) continue;
                    const double dd = (t.x - a.x) * (t.x - a.x) + (t.y - a.y) * (t.y - a.y) + (t.z - a.z) * (t.z - a.z);
                    if (dd < best) { best = dd; b = t; have = true; }
                }
        if (!have) continue;
        // c: smallest circumcircle through a, b: R^2 = |ab|^2 |ac|^2 |bc|^2 / (4 |ab x ac|^2)
        {
            const double ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, uu = ux * ux + uy * uy + uz * uz;
            double bn = 0, bd = -1;
            have = false;
            for (int d2 = -R2; d2 <= R2; ++d2)
                for (int d1 = -R1; d1 <= R1; ++d1)
                    for (int d0 = -R0; d0 <= R0; ++d0) {
                        const int p0 = i0 + d0, p1 = i1 + d1, p2 = i2 + d2;
                        if (same3(a, p0, p1, p2) || same3(b, p0, p1, p2) || !src(p0, p1, p2, t)) continue;
                        const double vx = t.x - a.x, vy = t.y - a.y, vz = t.z - a.z, vv = vx * vx + vy * vy + vz * vz;
                        const double wx = t.x - b.x, wy = t.y - b.y, wz = t.z - b.z, ww = wx * wx + wy * wy + wz * wz;
                        const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
                        const double cc = cx * cx + cy * cy + cz * cz;
                        if (!(cc > 1e-30 * uu * vv)) continue;                  // collinear
                        const double num = uu * vv * ww;
                        if (!have || num * bd < bn * cc) { bn = num; bd = cc; c = t; have = true; }
                    }
            if (!have) continue;
        }
        if (orient3(a, b, c, qx, qy, qz) < 0) { t = b; b = c; c = t; }
        for (int step = 0; step < 96; ++step) {
            // circumcentre o and radius of the face, unit normal towards the node's side
            const double ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z, vx = c.x - a.x, vy = c.y - a.y, vz = c.z - a.z;
            const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
            const double nn = nx * nx + ny * ny + nz * nz, uu = ux * ux + uy * uy + uz * uz, vv = vx * vx + vy * vy + vz * vz;
            // o - a = (|u|^2 (v x n) + |v|^2 (n x u)) / (2 |n|^2)
            const double ox = (uu * (vy * nz - vz * ny) + vv * (ny * uz - nz * uy)) / (2 * nn);
            const double oy = (uu * (vz * nx - vx * nz) + vv * (nz * ux - nx * uz)) / (2 * nn);
            const double oz = (uu * (vx * ny - vy * nx) + vv * (nx * uy - ny * ux)) / (2 * nn);
            const double R2c = ox * ox + oy * oy + oz * oz;
            const double inv_n = rsqrt(nn);
            const double ex = nx * inv_n, ey = ny * inv_n, ez = nz * inv_n;
            // fourth vertex: the point on the node's side whose sphere through the face has the lowest centre along the
            // normal, t = (|p - o|^2 - R^2) / (2 h); candidates are compared as fractions (no division per candidate)
            double bnum = 0, bh = -1;
            bool any = false;
            for (int d2 = -R2; d2 <= R2; ++d2)
                for (int d1 = -R1; d1 <= R1; ++d1)
                    for (int d0 = -R0; d0 <= R0; ++d0) {
                        const int p0 = i0 + d0, p1 = i1 + d1, p2 = i2 + d2;
                        if (!src(p0, p1, p2, t)) continue;
                        const double px = t.x - a.x - ox, py = t.y - a.y - oy, pz = t.z - a.z - oz;
                        const double h = px * ex + py * ey + pz * ez;
                        if (!(h > 1e-14)) continue;                             // not on the node's side of the face
                        if (same3(a, p0, p1, p2) || same3(b, p0, p1, p2) || same3(c, p0, p1, p2)) continue;
                        const double num = px * px + py * py + pz * pz - R2c;
                        if (!any || num * bh < bnum * h) { bnum = num; bh = h; d = t; any = true; }
                    }
            if (!any) break;                                                    // the node is outside the convex hull
            const double bt = bnum / (2 * bh);
            const double vol = orient3(a, b, c, d.x, d.y, d.z);
            // barycentric coordinates of the node: replace one vertex by the node in the orientation determinant
            Cand3 qn; qn.x = qx; qn.y = qy; qn.z = qz;
            wd = orient3(a, b, c, qx, qy, qz) / vol;
            wa = orient3(qn, b, c, d.x, d.y, d.z) / vol;
            wb = orient3(a, qn, c, d.x, d.y, d.z) / vol;
            wc = orient3(a, b, qn, d.x, d.y, d.z) / vol;
            if (wa >= 0 && wb >= 0 && wc >= 0) {
                found = true;
                scx = a.x + ox + bt * ex; scy = a.y + oy + bt * ey; scz = a.z + oz + bt * ez;
                srad2 = R2c + bt * bt;
                break;
            }
            // leave through the face opposite the most negative coordinate; orient the new face towards the node
            if (wa <= wb && wa <= wc)      { a = d; }          // face (d, b, c)
            else if (wb <= wa && wb <= wc) { b = d; }          // face (a, d, c)
            else                           { c = d; }          // face (a, b, d)
            if (orient3(a, b, c, qx, qy, qz) < 0) { t = b; b = c; c = t; }
        }
        if (found && !full_patch) {
            // Is the circumsphere empty of the points of the outer shell?  Only lattice points whose cell can reach the
            // sphere need a look: a point at lattice offset d sits within maxdisp cells of q + d h along every axis.
            const double lim = srad2 * (1.0 - 1e-10), rad = sqrt(srad2);
            const double h0 = P.ax[0].delta / P.ax[0].rg, h1 = P.ax[1].delta / P.ax[1].rg, h2 = P.ax[2].delta / P.ax[2].rg;
            // a point at lattice offset d sits at q + d h + s with s inside [slk lo, slk hi] along every axis
            const int lo0 = max(-r0, (int)floor((scx - rad - qx - slk[1]) / h0)), hi0 = min(r0, (int)ceil((scx + rad - qx - slk[0]) / h0));
            const int lo1 = max(-r1, (int)floor((scy - rad - qy - slk[3]) / h1)), hi1 = min(r1, (int)ceil((scy + rad - qy - slk[2]) / h1));
            const int lo2 = max(-r2, (int)floor((scz - rad - qz - slk[5]) / h2)), hi2 = min(r2, (int)ceil((scz + rad - qz - slk[4]) / h2));
            // ... and of that box only the part of every x row the sphere can reach: a row whose y / z distance to the centre
            // (less the slack) exceeds the radius is skipped, and the others are cut to the chord of the sphere at that
            // distance (the box alone holds 180 lattice points per node at 512 x 512 x 128 and 460 at 1024 x 1024 x 128: more
            // than the scans of the search itself)
            const double wide2 = srad2 * (1.0 + 1e-9);
            for (int d2 = lo2; d2 <= hi2 && found; ++d2) {
                const double zc = qz + d2 * h2 - scz, ez = fmax(fmax(zc + slk[4], -(zc + slk[5])), 0.0), rz2 = wide2 - ez * ez;
                if (rz2 < 0) continue;
                for (int d1 = lo1; d1 <= hi1 && found; ++d1) {
                    const double yc = qy + d1 * h1 - scy, ey = fmax(fmax(yc + slk[2], -(yc + slk[3])), 0.0), w2 = rz2 - ey * ey;
                    if (w2 < 0) continue;
                    const double w = sqrt(w2);
                    const int a0 = max(lo0, (int)floor((scx - w - qx - slk[1]) / h0)), b0 = min(hi0, (int)ceil((scx + w - qx - slk[0]) / h0));
                    const bool inner_row = d2 >= -R2 && d2 <= R2 && d1 >= -R1 && d1 <= R1;
                    for (int d0 = a0; d0 <= b0; ++d0) {
                        if (inner_row && d0 >= -R0 && d0 <= R0) { d0 = R0; continue; }      // skip the inner patch
                        if (!src(i0 + d0, i1 + d1, i2 + d2, t)) continue;
                        const double dd = (t.x - scx) * (t.x - scx) + (t.y - scy) * (t.y - scy) + (t.z - scz) * (t.z - scz);
                        if (dd < lim) { found = false; break; }
                    }
                }
            }
        }
    }
    if (!found) { none(); return; }
    {
        const Cand3 *vv[4] = {&a, &b, &c, &d};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            int q[3] = {vv[v]->p0, vv[v]->p1, vv[v]->p2};
            int64_t idx = 0;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const RemapAxis &A = P.ax[ax];
                if (A.periodic) { q[ax] %= A.n; if (q[ax] < 0) q[ax] += A.n; }
                idx += (int64_t)(q[ax] + A.H) * A.stride;
            }
            tet[4 * node + v] = idx;
        }
    }
    wgt[4 * node] = wa; wgt[4 * node + 1] = wb; wgt[4 * node + 2] = wc; wgt[4 * node + 3] = 1.0 - wa - wb - wc;
    // patch sufficiency: the circumsphere must stay inside the part of space the patch is guaranteed to cover.  Nodes
    // on the first / last plane of a Bounded axis are overwritten by the edge-plane interpolation and are not checked.
    const int ii[3] = {i0, i1, i2};
    const double qq[3] = {qx, qy, qz}, sc[3] = {scx, scy, scz};
    const double rad = sqrt(srad2);
    bool bad = false;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const RemapAxis &A = P.ax[q];
        if (!A.periodic && (ii[q] == 0 || ii[q] == A.n - 1)) return;
        const double safe = (A.r - A.maxdisp - 0.5) * A.delta / A.rg;
        const bool edge_lo = !A.periodic && ii[q] - A.r < 0, edge_hi = !A.periodic && ii[q] + A.r >= A.n;
        if (!edge_lo && sc[q] - rad < qq[q] - safe) bad = true;
        if (!edge_hi && sc[q] + rad > qq[q] + safe) bad = true;
    }
    if (bad) atomicAdd(too_small, 1);
}

// candidate sources
struct Src3Global {
    const Remap3Params &P;
    const double *xi0, *xi1, *xi2;
    __device__ __forceinline__ bool operator()(int p0, int p1, int p2, Cand3 &t) const { return candidate3(P, xi0, xi1, xi2, p0, p1, p2, t); }
};

// one thread per interior node, candidates evaluated from global memory (any patch radius)
__global__ void lfb_remap3d_locate(const Remap3Params P, const double *__restrict__ xi0, const double *__restrict__ xi1,
                                   const double *__restrict__ xi2, int64_t *__restrict__ tet, double *__restrict__ wgt,
                                   int *__restrict__ too_small)
{
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
    if (i0 >= P.ax[0].n) return;
    const Src3Global src = {P, xi0, xi1, xi2};
    double slk[6];
    for (int a = 0; a < 3; ++a) { slk[2 * a + 1] = P.ax[a].maxdisp * P.ax[a].delta / P.ax[a].rg * (1.0 + 1e-12); slk[2 * a] = -slk[2 * a + 1]; }
    remap3_locate_node(P, src, i0, i1, i2, tet, wgt, too_small, slk);
}

// The same with the patch of a block of 32 x BY nodes staged once in shared memory (normalised positions of every
// lattice neighbour any node of the block can reach; x = +inf marks an image the cloud does not hold), so a candidate
// costs three shared-memory loads instead of three global loads, a wrap, a floor and three divisions.
#define R3_BX 32
struct Src3Shared {
    const double *px, *py, *pz;
    int b0, b1, b2, e0, e1;          // lattice offset of patch entry 0 and the patch extents in x, y
    __device__ __forceinline__ bool operator()(int p0, int p1, int p2, Cand3 &t) const
    {
        const int e = ((p2 - b2) * e1 + (p1 - b1)) * e0 + (p0 - b0);
        const double x = px[e];
        if (!(x < 1e300)) return false;
        t.x = x; t.y = py[e]; t.z = pz[e]; t.p0 = p0; t.p1 = p1; t.p2 = p2; t.idx = -1;
        return true;
    }
};

template <int R3_BY>
__global__ void __launch_bounds__(R3_BX * R3_BY, 1) lfb_remap3d_locate_smem(const Remap3Params P, const double *__restrict__ xi0,
                                                                        const double *__restrict__ xi1, const double *__restrict__ xi2,
                                                                        int64_t *__restrict__ tet, double *__restrict__ wgt,
                                                                        int *__restrict__ too_small)
{
    extern __shared__ double r3_smem[];
    const int r0 = P.ax[0].r, r1 = P.ax[1].r, r2 = P.ax[2].r;
    const int e0 = R3_BX + 2 * r0, e1 = R3_BY + 2 * r1, e2 = 2 * r2 + 1, ne = e0 * e1 * e2;
    double *px = r3_smem, *py = px + ne, *pz = py + ne;
    const int n0 = blockIdx.x * R3_BX, n1 = blockIdx.y * R3_BY, i2 = blockIdx.z;
    const int b0 = n0 - r0, b1 = n1 - r1, b2 = i2 - r2;
    // while staging: the range of (position - lattice position) over the patch, per axis (lo as a maximum of the negated value)
    double sl[6] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        const int q0 = e % e0, q1 = (e / e0) % e1, q2 = e / (e0 * e1);
        Cand3 t;
        if (candidate3(P, xi0, xi1, xi2, b0 + q0, b1 + q1, b2 + q2, t)) {
            px[e] = t.x; py[e] = t.y; pz[e] = t.z;
            const double v[3] = {t.x, t.y, t.z};
            const int pp[3] = {b0 + q0, b1 + q1, b2 + q2};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double off = v[a] - (P.ax[a].origin + P.ax[a].delta * (pp[a] + 0.5) - P.ax[a].mn) / P.ax[a].rg;
                sl[2 * a] = fmax(sl[2 * a], -off); sl[2 * a + 1] = fmax(sl[2 * a + 1], off);
            }
        } else { px[e] = INFINITY; py[e] = 0; pz[e] = 0; }
    }
    __shared__ double r3_slack[R3_BY][6];
#pragma unroll
    for (int v = 0; v < 6; ++v) {
        for (int o = 16; o > 0; o >>= 1) sl[v] = fmax(sl[v], __shfl_xor_sync(0xffffffffu, sl[v], o));
        if ((threadIdx.x & 31) == 0) r3_slack[threadIdx.x >> 5][v] = sl[v];
    }
    __syncthreads();
    const int i0 = n0 + threadIdx.x % R3_BX, i1 = n1 + threadIdx.x / R3_BX;
    if (i0 >= P.ax[0].n || i1 >= P.ax[1].n) return;
    double slk[6];
#pragma unroll
    for (int v = 0; v < 6; ++v) {
        double m = r3_slack[0][v];
        for (int w = 1; w < R3_BY; ++w) m = fmax(m, r3_slack[w][v]);
        // widened by rounding room; an empty patch (no data point at all) leaves the global bound
        const double g = P.ax[v / 2].maxdisp * P.ax[v / 2].delta / P.ax[v / 2].rg;
        m = m > -1e300 ? m + 1e-12 : g;
        slk[v] = (v & 1) ? m : -m;
    }
    const Src3Shared src = {px, py, pz, b0, b1, b2, e0, e1};
    remap3_locate_node(P, src, i0, i1, i2, tet, wgt, too_small, slk);
}

__global__ void lfb_remap3d_apply(const Remap3Params P, const int64_t *__restrict__ tet, const double *__restrict__ wgt,
                                  const double *__restrict__ f, double *__restrict__ out)
{
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, i1 = blockIdx.y, i2 = blockIdx.z;
    if (i0 >= P.ax[0].n) return;
    const int64_t node = ((int64_t)i2 * P.ax[1].n + i1) * P.ax[0].n + i0;
    const int64_t o = (int64_t)(i0 + P.ax[0].H) * P.ax[0].stride + (int64_t)(i1 + P.ax[1].H) * P.ax[1].stride +
                      (int64_t)(i2 + P.ax[2].H) * P.ax[2].stride;
    if (tet[4 * node] < 0) { out[o] = nan(""); return; }
    out[o] = wgt[4 * node] * f[tet[4 * node]] + wgt[4 * node + 1] * f[tet[4 * node + 1]] + wgt[4 * node + 2] * f[tet[4 * node + 2]] +
             wgt[4 * node + 3] * f[tet[4 * node + 3]];
}

extern "C" void lfb_internal_set_error(const char *msg);      // lfb_api.cu: message returned by lfb_last_error()
#define RFAIL(code, ...) do { char m_[256]; snprintf(m_, sizeof(m_), __VA_ARGS__); lfb_internal_set_error(m_); return code; } while (0)
#define RCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) RFAIL(LFB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

// normalisation extents of the reference's padded cloud along axis `a` (post:527-628) over the points of the index box
// [lo, hi) and the largest displacement in cells: per-block partial results (min, max, max |xi| / delta)
struct ExtentParams {
    int     lo[3], n[3], H[3];      // box origin and extents (interior indices), halos
    int64_t st[3];
    int     a, periodic, has_map;
    double  x0, dl, L, pad;
};

__global__ void lfb_remap_extent(const ExtentParams E, const double *__restrict__ xi_a, double *__restrict__ partial)
{
    const int64_t total = (int64_t)E.n[0] * E.n[1] * E.n[2];
    double mn = INFINITY, mx = -INFINITY, md = 0;
    const double xe = E.x0 + E.L;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        int ii[3];
        ii[0] = E.lo[0] + (int)(q % E.n[0]); ii[1] = E.lo[1] + (int)((q / E.n[0]) % E.n[1]); ii[2] = E.lo[2] + (int)(q / ((int64_t)E.n[0] * E.n[1]));
        const int64_t idx = (int64_t)(ii[0] + E.H[0]) * E.st[0] + (int64_t)(ii[1] + E.H[1]) * E.st[1] + (int64_t)(ii[2] + E.H[2]) * E.st[2];
        const double d = E.has_map ? xi_a[idx] : 0.0;
        double X = E.x0 + E.dl * (ii[E.a] + 0.5) + d;
        md = fmax(md, fabs(d) / E.dl);
        if (!(fabs(d) < 1e300)) md = INFINITY;                                  // NaN / inf displacement
        if (E.periodic && E.has_map) X -= floor((X - E.x0) / E.L) * E.L;
        mn = fmin(mn, X); mx = fmax(mx, X);
        if (E.periodic) {
            if (X > xe - E.pad && X < xe) mn = fmin(mn, X - E.L);
            else if (X < E.x0 + E.pad && X > E.x0) mx = fmax(mx, X + E.L);
        }
    }
    __shared__ double smn[256], smx[256], smd[256];
    smn[threadIdx.x] = mn; smx[threadIdx.x] = mx; smd[threadIdx.x] = md;
    __syncthreads();
    for (int s2 = blockDim.x / 2; s2 > 0; s2 >>= 1) {
        if ((int)threadIdx.x < s2) {
            smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + s2]);
            smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + s2]);
            smd[threadIdx.x] = fmax(smd[threadIdx.x], smd[threadIdx.x + s2]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[3 * blockIdx.x] = smn[0]; partial[3 * blockIdx.x + 1] = smx[0]; partial[3 * blockIdx.x + 2] = smd[0]; }
}

#define EXT_BLOCKS 592
static int cloud_extent(const lfb_remap_desc *D, const double *d_xi_a, double *d_partial, int a, const int lo[3], const int hi[3],
                        const int64_t st[3], double *mn_out, double *rg_out, double *md_out)
{
    ExtentParams E;
    memset(&E, 0, sizeof(E));
    for (int q = 0; q < 3; ++q) { E.lo[q] = lo[q]; E.n[q] = hi[q] - lo[q]; E.H[q] = D->halo[q]; E.st[q] = st[q]; }
    E.a = a; E.periodic = D->topology[a] == LFB_PERIODIC; E.has_map = D->has_map[a] && d_xi_a;
    E.x0 = D->origin[a]; E.dl = D->spacing[a]; E.L = E.dl * D->size[a]; E.pad = D->npad * E.dl;
    lfb_remap_extent<<<EXT_BLOCKS, 256>>>(E, d_xi_a, d_partial);
    static double hp[3 * EXT_BLOCKS];
    if (cudaGetLastError() != cudaSuccess || cudaMemcpy(hp, d_partial, sizeof(hp), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    double mn = INFINITY, mx = -INFINITY, md = 0;
    for (int b2 = 0; b2 < EXT_BLOCKS; ++b2) { mn = fmin(mn, hp[3 * b2]); mx = fmax(mx, hp[3 * b2 + 1]); md = fmax(md, hp[3 * b2 + 2]); }
    if (!(md < 1e300)) return -1;
    *mn_out = mn; *rg_out = mx - mn; *md_out = md;
    return 0;
}

// Work arrays of the 3-D remap are cached between calls (one call per output time, always the same sizes): allocating and
// freeing 14 GB per call at 1024 x 1024 x 128 costs more than the edge planes and the field loop together.
// lfb_remap_release() returns them to the device.
struct WsBlock { void *p; size_t bytes; int device; };
static std::vector<WsBlock> g_ws_free, g_ws_used;
static std::mutex g_ws_lock;

static cudaError_t ws_alloc(void **p, size_t bytes)
{
    std::lock_guard<std::mutex> guard(g_ws_lock);
    int dev = 0;
    cudaGetDevice(&dev);
    for (size_t q = 0; q < g_ws_free.size(); ++q)
        if (g_ws_free[q].bytes == bytes && g_ws_free[q].device == dev) {
            *p = g_ws_free[q].p;
            g_ws_used.push_back(g_ws_free[q]);
            g_ws_free.erase(g_ws_free.begin() + q);
            return cudaSuccess;
        }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && !g_ws_free.empty()) {           // give the cached blocks back and retry
        cudaGetLastError();
        for (auto &b : g_ws_free) cudaFree(b.p);
        g_ws_free.clear();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) g_ws_used.push_back({*p, bytes, dev});
    return e;
}

static void ws_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> guard(g_ws_lock);
    for (size_t q = 0; q < g_ws_used.size(); ++q)
        if (g_ws_used[q].p == p) { g_ws_free.push_back(g_ws_used[q]); g_ws_used.erase(g_ws_used.begin() + q); return; }
    cudaFree(p);
}

extern "C" int lfb_remap_release(void)
{
    std::lock_guard<std::mutex> guard(g_ws_lock);
    cudaDeviceSynchronize();
    for (auto &b : g_ws_free) cudaFree(b.p);
    g_ws_free.clear();
    return LFB_OK;
}

// grids with three non-Flat axes: 3-D walk for the bulk, then the reference's edge-plane interpolation on every
// Bounded axis (a 2-D LinearNDInterpolator over the points of the first / last index plane, normalised by the extents
// of that cut, post:678-745), halos zero
static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static int remap3(const lfb_remap_desc *D, const double *const *xi, int n_fields, const double *const *fields, double *const *out)
{
    const bool timing = getenv("LFB200_REMAP_TIMING") != nullptr;
    const double t_start = now_s();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) RFAIL(LFB_ERR_CUDA, "no CUDA device available; liblfb200 has no CPU fallback");
    RCU(cudaSetDevice(D->device));
    int64_t ext[3], st[3];
    for (int a = 0; a < 3; ++a) ext[a] = D->size[a] + 2 * D->halo[a];
    st[0] = 1; st[1] = ext[0]; st[2] = ext[0] * ext[1];
    const int64_t len = ext[0] * ext[1] * ext[2];
    Remap3Params P;
    memset(&P, 0, sizeof(P));
    const int lo[3] = {0, 0, 0}, hi[3] = {D->size[0], D->size[1], D->size[2]};
    double *d_xi[3] = {nullptr, nullptr, nullptr}, *d_f = nullptr, *d_out = nullptr, *d_w = nullptr, *d_pw = nullptr, *d_part = nullptr;
    int64_t *d_tet = nullptr, *d_ptri = nullptr;
    int *d_flag = nullptr;
    const int64_t nodes = (int64_t)D->size[0] * D->size[1] * D->size[2];
    auto cleanup = [&]() {
        for (int a = 0; a < 3; ++a) ws_free(d_xi[a]);
        ws_free(d_f); ws_free(d_out); ws_free(d_w); ws_free(d_tet); ws_free(d_flag); ws_free(d_pw); ws_free(d_ptri); ws_free(d_part);
    };
#define RCU3(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); RFAIL(LFB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)
    for (int a = 0; a < 3; ++a)
        if (D->has_map[a] && xi[a]) {
            RCU3(ws_alloc((void **)&d_xi[a], len * sizeof(double)));
            RCU3(cudaMemcpy(d_xi[a], xi[a], len * sizeof(double), cudaMemcpyDefault));
        }
    RCU3(ws_alloc((void **)&d_part, 3 * EXT_BLOCKS * sizeof(double)));
    for (int a = 0; a < 3; ++a) {
        RemapAxis &A = P.ax[a];
        A.n = D->size[a]; A.H = D->halo[a]; A.periodic = D->topology[a] == LFB_PERIODIC; A.has_map = D->has_map[a] && xi[a];
        A.origin = D->origin[a]; A.delta = D->spacing[a]; A.L = A.delta * A.n; A.pad = D->npad * A.delta; A.stride = st[a];
        const int ce = cloud_extent(D, d_xi[a], d_part, a, lo, hi, st, &A.mn, &A.rg, &A.maxdisp);
        if (ce) { cleanup(); RFAIL(ce == -1 ? LFB_ERR_ARG : LFB_ERR_CUDA, ce == -1 ? "non-finite displacement" : "extent kernel failed"); }
    }
    const double t_ext = now_s();
    RCU3(ws_alloc((void **)&d_f, len * sizeof(double)));
    RCU3(ws_alloc((void **)&d_out, len * sizeof(double)));
    RCU3(ws_alloc((void **)&d_w, 4 * nodes * sizeof(double)));
    RCU3(ws_alloc((void **)&d_tet, 4 * nodes * sizeof(int64_t)));
    RCU3(ws_alloc((void **)&d_flag, sizeof(int)));
    // patch radius: displacement + the reach of a lattice circumsphere in normalised cells + margin; doubled while the check fails
    double hn[3], diag = 0;
    for (int a = 0; a < 3; ++a) { hn[a] = P.ax[a].delta / P.ax[a].rg; diag += hn[a] * hn[a]; }
    diag = 0.5 * sqrt(diag);
    int r[3];
    for (int a = 0; a < 3; ++a) {
        r[a] = (int)ceil(P.ax[a].maxdisp) + 2 + (int)fmin(ceil(diag / hn[a]) - 1, 12.0);
        if (D->radius > 0) r[a] = D->radius;
        if (r[a] > P.ax[a].n) r[a] = P.ax[a].n;
    }
    dim3 block(64, 1, 1), grid((P.ax[0].n + 63) / 64, P.ax[1].n, P.ax[2].n);
    int flag = 1;
    for (int attempt = 0; attempt < 3 && flag; ++attempt) {
        for (int a = 0; a < 3; ++a) P.ax[a].r = r[a];
        RCU3(cudaMemset(d_flag, 0, sizeof(int)));
        // rows of nodes per block: the more nodes share a staged patch, the less staging per node and the more warps per
        // SM; 8 rows (256 threads) when that patch fits, else 4
        auto patch_of = [&](int by) { return (size_t)(R3_BX + 2 * r[0]) * (by + 2 * r[1]) * (2 * r[2] + 1) * 3 * sizeof(double); };
        const int by = (patch_of(8) <= 200 * 1024 && !getenv("LFB200_REMAP_BY4")) ? 8 : 4;
        const size_t patch_bytes = patch_of(by);
        if (patch_bytes <= 200 * 1024 && !getenv("LFB200_REMAP_GLOBAL")) {
            dim3 gs((P.ax[0].n + R3_BX - 1) / R3_BX, (P.ax[1].n + by - 1) / by, P.ax[2].n);
            if (by == 8) {
                RCU3(cudaFuncSetAttribute(lfb_remap3d_locate_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_bytes));
                lfb_remap3d_locate_smem<8><<<gs, R3_BX * 8, patch_bytes>>>(P, d_xi[0], d_xi[1], d_xi[2], d_tet, d_w, d_flag);
            } else {
                RCU3(cudaFuncSetAttribute(lfb_remap3d_locate_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)patch_bytes));
                lfb_remap3d_locate_smem<4><<<gs, R3_BX * 4, patch_bytes>>>(P, d_xi[0], d_xi[1], d_xi[2], d_tet, d_w, d_flag);
            }
        } else {
            lfb_remap3d_locate<<<grid, block>>>(P, d_xi[0], d_xi[1], d_xi[2], d_tet, d_w, d_flag);
        }
        RCU3(cudaGetLastError());
        RCU3(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) for (int a = 0; a < 3; ++a) { r[a] = r[a] * 2; if (r[a] > P.ax[a].n) r[a] = P.ax[a].n; }
    }
    if (flag) { cleanup(); RFAIL(LFB_ERR_STATE, "remap patch too small even at radius (%d, %d, %d): displacements too large for the local search", P.ax[0].r, P.ax[1].r, P.ax[2].r); }
    const double t_loc = now_s();

    // edge planes of the Bounded axes: triangle + weights per plane, shared by all fields
    struct Plane { RemapParams P2; int64_t base; int64_t off; int n0, n1; };
    std::vector<Plane> planes;
    int64_t plane_nodes = 0;
    for (int q = 0; q < 3; ++q) {
        if (D->topology[q] != LFB_BOUNDED) continue;
        const int o0 = q == 0 ? 1 : 0, o1 = q == 2 ? 1 : 2;
        for (int edge : {0, D->size[q] - 1}) {
            Plane pl;
            memset(&pl, 0, sizeof(pl));
            int plo[3] = {0, 0, 0}, phi[3] = {D->size[0], D->size[1], D->size[2]};
            plo[q] = edge; phi[q] = edge + 1;
            const int oa[2] = {o0, o1};
            for (int s2 = 0; s2 < 2; ++s2) {
                const int a = oa[s2];
                RemapParams &R = pl.P2;
                R.n[s2] = D->size[a]; R.H[s2] = D->halo[a]; R.e[s2] = (int)ext[a];
                R.periodic[s2] = D->topology[a] == LFB_PERIODIC; R.has_map[s2] = P.ax[a].has_map;
                R.origin[s2] = D->origin[a]; R.delta[s2] = D->spacing[a]; R.L[s2] = P.ax[a].L; R.stride[s2] = st[a];
                R.pad[s2] = P.ax[a].pad;
                if (cloud_extent(D, d_xi[a], d_part, a, plo, phi, st, &R.mn[s2], &R.rg[s2], &R.maxdisp[s2])) { cleanup(); RFAIL(LFB_ERR_ARG, "non-finite displacement"); }
            }
            pl.base = (int64_t)(edge + D->halo[q]) * st[q];
            pl.off = plane_nodes; pl.n0 = pl.P2.n[0]; pl.n1 = pl.P2.n[1];
            plane_nodes += (int64_t)pl.n0 * pl.n1;
            planes.push_back(pl);
        }
    }
    if (!planes.empty()) {
        RCU3(ws_alloc((void **)&d_pw, 3 * plane_nodes * sizeof(double)));
        RCU3(ws_alloc((void **)&d_ptri, 3 * plane_nodes * sizeof(int64_t)));
    }
    int axis_of_plane = 0;
    (void)axis_of_plane;
    for (size_t ip = 0; ip < planes.size(); ++ip) {
        Plane &pl = planes[ip];
        RemapParams &R = pl.P2;
        // which 3-D axes are the plane's two axes (for the xi pointers)
        int oa[2] = {-1, -1};
        for (int a = 0, s2 = 0; a < 3 && s2 < 2; ++a) if (st[a] == R.stride[s2]) oa[s2++] = a;
        const double h0 = R.delta[0] / R.rg[0], h1 = R.delta[1] / R.rg[1];
        int rr[2];
        rr[0] = (int)ceil(R.maxdisp[0]) + 3 + (h0 < h1 ? (int)fmin(ceil(h1 / h0) - 1, 12.0) : 0);
        rr[1] = (int)ceil(R.maxdisp[1]) + 3 + (h1 < h0 ? (int)fmin(ceil(h0 / h1) - 1, 12.0) : 0);
        dim3 g2((R.n[0] + 63) / 64, R.n[1], 1);
        int f2 = 1;
        for (int attempt = 0; attempt < 4 && f2; ++attempt) {
            R.r[0] = rr[0] > R.n[0] ? R.n[0] : rr[0]; R.r[1] = rr[1] > R.n[1] ? R.n[1] : rr[1];
            RCU3(cudaMemset(d_flag, 0, sizeof(int)));
            lfb_remap2d_locate<<<g2, block>>>(R, d_xi[oa[0]] ? d_xi[oa[0]] + pl.base : nullptr, d_xi[oa[1]] ? d_xi[oa[1]] + pl.base : nullptr,
                                              d_ptri + 3 * pl.off, d_pw + 3 * pl.off, d_flag);
            RCU3(cudaGetLastError());
            RCU3(cudaMemcpy(&f2, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
            if (f2) { rr[0] *= 2; rr[1] *= 2; }
        }
        if (f2) { cleanup(); RFAIL(LFB_ERR_STATE, "remap patch too small on an edge plane"); }
    }
    for (int v = 0; v < n_fields; ++v) {
        RCU3(cudaMemcpy(d_f, fields[v], len * sizeof(double), cudaMemcpyDefault));
        RCU3(cudaMemset(d_out, 0, len * sizeof(double)));                      // halos stay zero (post:748)
        lfb_remap3d_apply<<<grid, block>>>(P, d_tet, d_w, d_f, d_out);
        RCU3(cudaGetLastError());
        for (Plane &pl : planes) {                                              // bounded dims in axis order (post:641)
            dim3 g2((pl.n0 + 63) / 64, pl.n1, 1);
            lfb_remap2d_apply<<<g2, block>>>(pl.P2, d_ptri + 3 * pl.off, d_pw + 3 * pl.off, d_f + pl.base, d_out + pl.base);
            RCU3(cudaGetLastError());
        }
        RCU3(cudaMemcpy(out[v], d_out, len * sizeof(double), cudaMemcpyDefault));
    }
    cleanup();
#undef RCU3
    if (timing)
        fprintf(stderr, "lfb_remap_to_mean 3-D: %lld nodes, radius (%d, %d, %d): extents %.3f s, upload + tetrahedron search %.3f s, "
                        "edge planes + %d fields %.3f s\n", (long long)nodes, P.ax[0].r, P.ax[1].r, P.ax[2].r, t_ext - t_start, t_loc - t_ext,
                n_fields, now_s() - t_loc);
    return LFB_OK;
}

extern "C" int lfb_remap_to_mean(const lfb_remap_desc *D, const double *const *xi, int n_fields, const double *const *fields,
                                 double *const *out)
{
    if (!D || !xi || !fields || !out || n_fields < 1) RFAIL(LFB_ERR_ARG, "bad argument");
    int ax[3], nt = 0;
    for (int a = 0; a < 3; ++a) if (D->topology[a] != LFB_FLAT) ax[nt++] = a;
    if (nt == 3) return remap3(D, xi, n_fields, fields, out);
    if (nt != 2) RFAIL(LFB_ERR_ARG, "the device remap handles grids with two or three non-Flat axes (got %d)", nt);
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
    double *d_xi[2] = {nullptr, nullptr}, *d_f = nullptr, *d_out = nullptr, *d_w = nullptr, *d_part = nullptr;
    int64_t *d_tri = nullptr;
    int *d_flag = nullptr;
    const int64_t nodes = (int64_t)P.n[0] * P.n[1];
    auto cleanup = [&]() { cudaFree(d_xi[0]); cudaFree(d_xi[1]); cudaFree(d_f); cudaFree(d_out); cudaFree(d_w); cudaFree(d_tri); cudaFree(d_flag); cudaFree(d_part); };
    for (int q = 0; q < 2; ++q)
        if (P.has_map[q]) {
            RCU(cudaMalloc(&d_xi[q], len * sizeof(double)));
            RCU(cudaMemcpy(d_xi[q], xi[ax[q]], len * sizeof(double), cudaMemcpyDefault));      // host or device pointer
        }
    // normalisation extents of the padded cloud (post:527-628) and the largest displacements, reduced on the device
    RCU(cudaMalloc(&d_part, 3 * EXT_BLOCKS * sizeof(double)));
    {
        const int lo3[3] = {0, 0, 0}, hi3[3] = {D->size[0], D->size[1], D->size[2]};
        for (int q = 0; q < 2; ++q) {
            const int ce = cloud_extent(D, d_xi[q], d_part, ax[q], lo3, hi3, st, &P.mn[q], &P.rg[q], &P.maxdisp[q]);
            if (ce) { cleanup(); RFAIL(ce == -1 ? LFB_ERR_ARG : LFB_ERR_CUDA, ce == -1 ? "non-finite displacement" : "extent kernel failed"); }
            P.pad[q] = D->npad * P.delta[q];
        }
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
        RCU(cudaMemcpy(d_f, fields[v], len * sizeof(double), cudaMemcpyDefault));
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
        RCU(cudaMemcpy(out[v], d_out, len * sizeof(double), cudaMemcpyDefault));
    }
    cleanup();
    return LFB_OK;
}
