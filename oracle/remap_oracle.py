"""Remap-to-mean-position ORACLE (test infrastructure, NOT product code).

numpy/scipy restatement of regrid_to_mean_position! (reference
src/Utils/post_processing_utils.jl:327-762).  The reference itself calls
scipy.interpolate.LinearNDInterpolator / numpy.interp through PythonCall (post:634,670,685), so
the same scipy calls made here are the reference's own arithmetic for this step.  Pinned against
the golden `*_at_mean` fields in tests/test_oracle_golden.py.

Immersed boundaries are not covered (the reference only masks and warns there, post:356-362).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
from scipy.interpolate import LinearNDInterpolator


def centre_nodes(N, H, x0, delta):
    """Cell-centre coordinates with halos: xnodes(grid, Center(), with_halos=true) (post:249-251)."""
    return x0 + delta * (np.arange(1 - H, N + H + 1) - 0.5)


def remap_to_mean(fields: Dict[str, np.ndarray], xi: Dict[str, np.ndarray], N, H, topology, origin, delta,
                  npad: int = 5, pad_columns: str = "reference") -> Dict[str, np.ndarray]:
    """fields: {name: parent array with halos}; xi: {"u"/"v"/"w": parent xi array} for the velocities
    that have maps.  origin[ax] = first face coordinate, delta[ax] = spacing.  Returns
    {name + "_at_mean": parent array} with halos zeroed (post:748).

    pad_columns: "reference" follows post:523-611 literally: the coordinate column that receives the periodic padding
    only advances inside the periodic branches, so when a Bounded axis precedes a Periodic one the Periodic axis pads
    the WRONG coordinate column (a latent bug of the reference; no example or test of it has such a grid).
    "documented" pads every Periodic axis's own coordinate, which is what the reference's comments and docs describe
    and what liblfb200 computes; the two coincide whenever no Bounded axis precedes a Periodic one."""
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
    for ib, ax in enumerate(axes):                                        # post:527-611
        if topology[ax] != "Periodic":
            continue   # reference: column_number only advances inside periodic branches (x then y)
        if pad_columns == "documented":
            col = 2 * ib
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


def eulerian_filter(series: Sequence[np.ndarray], times, a, b, c, d, single_exp=False):
    """compute_Eulerian_filter! (post:1324-1380) with get_weight_function (post:1170-1197), direction "both"."""
    times = np.asarray(times, dtype=np.float64)
    out = []
    for t in times:
        at = np.abs(times - t)
        if single_exp:
            G = a[0] * np.exp(-c[0] * at)
        else:
            G = np.zeros_like(times)
            for ai, bi, ci, di in zip(a, b, c, d):
                G = G + (ai * np.cos(di * at) + bi * np.sin(di * at)) * np.exp(-ci * at)
        norm = np.sum(G[:-1] * np.diff(times))
        mean = np.zeros_like(np.asarray(series[0], dtype=np.float64))
        dt = times[1] - times[0]
        for j in range(len(times)):
            dt = times[j + 1] - times[j] if j < len(times) - 1 else dt
            mean = mean + G[j] * np.asarray(series[j], dtype=np.float64) * dt
        out.append(mean / norm)
    return out
