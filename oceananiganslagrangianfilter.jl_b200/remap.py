"""Remap of filtered fields to the mean position (reference regrid_to_mean_position!,
src/Utils/post_processing_utils.jl:327-762).

`remap_to_mean_device` -> lfb_remap_to_mean (csrc/lfb_remap.cu) remaps on the GPU: for every grid node the
Delaunay triangle (two non-Flat axes) or tetrahedron (three) that contains it is constructed locally by a walk
over a patch of lattice neighbours, followed by barycentric interpolation and the reference's edge fix on
Bounded axes; the search is shared by all variables of an output time.  There is no host form in the product: the
scipy restatement the reference itself calls (LinearNDInterpolator through PythonCall, post:8-13, 634, 685) lives
in oracle/remap_oracle.py as the checker.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
from . import _lib


def centre_nodes(N, H, x0, delta):
    """Cell-centre coordinates with halos: xnodes(grid, Center(), with_halos=true) (post:249-251)."""
    return x0 + delta * (np.arange(1 - H, N + H + 1) - 0.5)


def remap_to_mean_device(fields: Dict[str, object], xi: Dict[str, object], N, H, topology, origin, delta,
                         npad: int = 5, device: int = 0, radius: int = 0, immersed=None,
                         pointers: bool = False, alloc=None) -> Dict[str, np.ndarray]:
    """fields: {name: parent array with halos}; xi: {"u"/"v"/"w": parent xi array} for the velocities that have maps.
    origin[ax] = first face coordinate, delta[ax] = spacing.  Returns {name + "_at_mean": parent array} with halos
    zeroed (post:748), computed by liblfb200 (two or three non-Flat axes).

    immersed: boolean parent-shaped mask of the cells `_mask_immersed` blanks (post:285-298, 418-421, 508-511): their
    displacement is set to 0 and their value to NaN before the interpolation.
    pointers=True: the values of `fields` / `xi` are raw DEVICE pointers (FP64 parent arrays on `device`, e.g. from
    lfb_kept_info) instead of host arrays.  alloc(shape, dtype): allocator of the result arrays (default numpy.empty,
    Fortran order; page-locked memory makes the copies run at the full PCIe rate)."""
    import ctypes as C
    L = _lib.load()
    if immersed is not None:
        if pointers:
            raise ValueError("immersed masking works on host arrays")
        m = np.asarray(immersed, dtype=bool)
        xi = {k: np.where(m, 0.0, np.asarray(v, dtype=np.float64)) for k, v in xi.items()}
        fields = {k: np.where(m, np.nan, np.asarray(v, dtype=np.float64)) for k, v in fields.items()}
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
        if pointers:
            return int(a)
        a = _lib.as_parent(a, shape, np.float64)
        keep.append(a)
        return a.ctypes.data

    xi_ptrs = (C.c_void_p * 3)(*[ptr(xi["uvw"[ax]]) if "uvw"[ax] in xi else None for ax in range(3)])
    names = list(fields.keys())
    f_ptrs = (C.c_void_p * len(names))(*[ptr(fields[n]) for n in names])
    outs = [np.empty(shape, dtype=np.float64, order="F") if alloc is None else alloc(shape, np.float64) for _ in names]
    o_ptrs = (C.c_void_p * len(names))(*[o.ctypes.data for o in outs])
    _lib.check(L.lfb_remap_to_mean(C.byref(d), xi_ptrs, len(names), f_ptrs, o_ptrs))
    return {n + "_at_mean": o for n, o in zip(names, outs)}


def remap_release():
    """Return the cached work arrays of the 3-D remap to the device."""
    _lib.check(_lib.load().lfb_remap_release())


def remap_kept_to_mean(config, model, held: Dict[str, int], alloc=None) -> Dict[str, np.ndarray]:
    """regrid_to_mean_position! for ONE output time whose combined fields are kept on the device (ids of GLOBAL
    FP64 parent arrays on this rank): the device remap reads them in place."""
    from .post import _immersed_mask, _regrid_names
    g = config.grid
    if len(g.nonflat_axes) < 2:
        raise ValueError("regridding to the mean position needs at least two non-Flat axes")
    label = config.label
    names = _regrid_names(config)
    vels = [vel for vel in config.velocity_names if "xi_" + vel + label in held]
    mask = _immersed_mask(g)
    if mask is not None:                      # the immersed blanking is applied to host copies
        return remap_to_mean_device({n: model.kept_fetch(held[n]) for n in names},
                                    {vel: model.kept_fetch(held["xi_" + vel + label]) for vel in vels},
                                    g.N, g.H, g.topology, g.origin, g.spacing, npad=config.npad,
                                    device=getattr(config, "device", 0), immersed=mask, alloc=alloc)
    ptr = lambda kid: model.kept_info(kid, pointer=True)[2]
    return remap_to_mean_device({n: ptr(held[n]) for n in names}, {vel: ptr(held["xi_" + vel + label]) for vel in vels},
                                g.N, g.H, g.topology, g.origin, g.spacing, npad=config.npad,
                                device=getattr(config, "device", 0), pointers=True, alloc=alloc)
