"""Remap of filtered fields to the mean position (reference regrid_to_mean_position!,
src/Utils/post_processing_utils.jl:327-762).

`remap_to_mean_device` -> lfb_remap_to_mean (csrc/lfb_remap.cu) remaps on the GPU: for every grid node the
Delaunay triangle (two non-Flat axes) or tetrahedron (three) that contains it is constructed locally by a walk
over a patch of lattice neighbours, followed by barycentric interpolation and the reference's edge fix on
Bounded axes; the search is shared by all variables of an output time.  `remap_to_mean` is the host form of the
same step with the calls the reference itself makes (scipy.interpolate.LinearNDInterpolator through PythonCall,
post:8-13, 634, 685); the product path (`post.regrid_to_mean_position`) uses the device form.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
from scipy.interpolate import LinearNDInterpolator

from . import _lib


def centre_nodes(N, H, x0, delta):
    """Cell-centre coordinates with halos: xnodes(grid, Center(), with_halos=true) (post:249-251)."""
    return x0 + delta * (np.arange(1 - H, N + H + 1) - 0.5)


def remap_to_mean(fields: Dict[str, np.ndarray], xi: Dict[str, np.ndarray], N, H, topology, origin, delta,
                  npad: int = 5) -> Dict[str, np.ndarray]:
    """fields: {name: parent array with halos}; xi: {"u"/"v"/"w": parent xi array} for the velocities
    that have maps.  origin[ax] = first face coordinate, delta[ax] = spacing.  Returns
    {name + "_at_mean": parent array} with halos zeroed (post:748)."""
    axes = [ax for ax in range(3) if topology[ax] != "Flat"]        # true_dims (post:412-416)
    full = tuple(N[ax] + 2 * H[ax] for ax in range(3))
    inner = tuple(slice(H[ax], H[ax] + N[ax]) if H[ax] else slice(None) for ax in range(3))   # _remove_halos
    L = [N[ax] * delta[ax] for ax in range(3)]
    nodes = [centre_nodes(N[ax], H[ax], origin[ax], delta[ax]) for ax in range(3)]
    mesh = []
    for ax in range(3):
        shape = [1, 1, 1]
        shape[ax] = full[ax]
        mesh.append(nodes[ax].reshape(shape) + np.zeros(full))
    vec = lambda a: a.reshape(-1, order="F")

    cols = []
    for ax in axes:
        name = "uvw"[ax]
        Xi = mesh[ax] + xi[name] if name in xi else mesh[ax].copy()     # post:419,437
        Xi = Xi[inner].copy()
        if name in xi and topology[ax] == "Periodic":                     # post:427-429
            Xi -= np.floor((Xi - origin[ax]) / L[ax]) * L[ax]
        idx = np.indices(Xi.shape)[ax] + 1.0                              # 1-based index column
        cols += [vec(Xi), vec(idx)]
    names = list(fields.keys())
    for nm in names:
        cols.append(vec(np.asarray(fields[nm], dtype=np.float64)[inner]))
    D = np.stack(cols, axis=1)

    col = 0
    for ax in axes:                                                       # post:527-611
        if topology[ax] != "Periodic":
            continue   # reference: column_number only advances inside periodic branches (x then y)
        lo, hi = origin[ax], origin[ax] + L[ax]
        pad = npad * delta[ax]
        X = D[:, col]
        m_hi = (X > hi - pad) & (X < hi)
        m_lo = (X < lo + pad) & (X > lo)
        sel = m_hi | m_lo
        E = D[sel].copy()
        xr = E[:, col].copy()
        xr[(xr > hi - pad) & (xr < hi)] -= L[ax]
        xr[(xr < lo + pad) & (xr > lo)] += L[ax]
        E[:, col] = xr
        D = np.vstack([D, E])
        col += 2
    nd = len(axes)
    coords = D[:, 0:2 * nd:2]
    vals = D[:, 2 * nd:]
    cmin = coords.min(axis=0)
    crng = coords.max(axis=0) - cmin
    cn = (coords - cmin) / crng
    mesh_n = [(mesh[ax] - cmin[i]) / crng[i] for i, ax in enumerate(axes)]

    out = {}
    for iv, nm in enumerate(names):
        interp = LinearNDInterpolator(cn, vals[:, iv])
        res = np.array(interp(*mesh_n))
        for ib, ax in enumerate(axes):                                    # post:641-745
            if topology[ax] != "Bounded":
                continue
            for edge in (1, N[ax]):
                eh = edge + H[ax] - 1                                     # 0-based parent index
                cut = D[D[:, 2 * ib + 1].astype(int) == edge]
                cut = np.hstack([cut[:, :2 * ib], cut[:, 2 * ib + 2:]])
                cut = cut[np.argsort(cut[:, 0], kind="stable")]
                ccoords = cut[:, 0:2 * (nd - 1):2]
                cvals = cut[:, 2 * (nd - 1) + iv]
                other = [a for a in axes if a != ax]
                sl = [slice(None)] * 3
                sl[ax] = eh
                if nd == 2:
                    o = other[0]
                    line = mesh[o][tuple(sl)].reshape(-1)
                    res_line = np.interp(line, ccoords[:, 0], cvals)
                    res[tuple(sl)] = res_line.reshape(res[tuple(sl)].shape)
                elif nd == 3:
                    mn = ccoords.min(axis=0)
                    rg = ccoords.max(axis=0) - mn
                    it2 = LinearNDInterpolator((ccoords - mn) / rg, cvals)
                    # NB the reference indexes coord_mins_cut by the ORIGINAL axis number (post:682-683,
                    # 709-710, 734-735), which is out of range for a 2-column array except for the
                    # bounded-z case; we normalise each remaining axis by its own column.
                    m2 = [(mesh[o][tuple(sl)] - mn[i]) / rg[i] for i, o in enumerate(other)]
                    res[tuple(sl)] = np.array(it2(*m2))
        for ax in range(3):                                               # _fill_halos!(…, 0.0), post:748
            if H[ax]:
                s = [slice(None)] * 3
                s[ax] = slice(0, H[ax]); res[tuple(s)] = 0.0
                s[ax] = slice(full[ax] - H[ax], full[ax]); res[tuple(s)] = 0.0
        out[nm + "_at_mean"] = res
    return out


def remap_to_mean_device(fields: Dict[str, np.ndarray], xi: Dict[str, np.ndarray], N, H, topology, origin, delta,
                         npad: int = 5, device: int = 0, radius: int = 0) -> Dict[str, np.ndarray]:
    """Same contract as remap_to_mean, computed by liblfb200 (two or three non-Flat axes)."""
    import ctypes as C
    L = _lib.load()
    d = _lib.RemapDesc()
    shape = tuple(N[ax] + 2 * H[ax] for ax in range(3))
    for ax in range(3):
        d.size[ax], d.halo[ax] = int(N[ax]), int(H[ax])
        d.topology[ax] = _lib.TOPOLOGY[topology[ax]]
        d.origin[ax], d.spacing[ax] = float(origin[ax]), float(delta[ax])
        d.has_map[ax] = int("uvw"[ax] in xi)
    d.npad, d.device, d.radius = int(npad), int(device), int(radius)
    keep = []

    def ptr(a):
        a = _lib.as_parent(a, shape, np.float64)
        keep.append(a)
        return a.ctypes.data

    xi_ptrs = (C.c_void_p * 3)(*[ptr(xi["uvw"[ax]]) if "uvw"[ax] in xi else None for ax in range(3)])
    names = list(fields.keys())
    f_ptrs = (C.c_void_p * len(names))(*[ptr(fields[n]) for n in names])
    outs = [np.empty(shape, dtype=np.float64, order="F") for _ in names]
    o_ptrs = (C.c_void_p * len(names))(*[o.ctypes.data for o in outs])
    _lib.check(L.lfb_remap_to_mean(C.byref(d), xi_ptrs, len(names), f_ptrs, o_ptrs))
    return {n + "_at_mean": o for n, o in zip(names, outs)}
