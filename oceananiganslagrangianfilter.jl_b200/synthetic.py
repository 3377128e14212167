"""Synthetic input series for parity tests and the benchmark (analytic, smooth,
sign-mixed, divergent velocities with w = 0 on the z walls, a blob and a front as filtered variables,
snapshots rounded to Float32 like the reference's files).

`make_snapshot` works with numpy (tests, CPU baseline) and with torch on the GPU (benchmark sizes) and
returns Oceananigans parent arrays: halos included, Julia (x, y, z) order, velocities on their faces.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np

from .grids import Bounded, Flat, RectilinearGrid


def _coords(grid: RectilinearGrid, axis: int, face: bool) -> np.ndarray:
    N, H, d, o = grid.N[axis], grid.H[axis], grid.spacing[axis], grid.origin[axis]
    if grid.topology[axis] == Flat:
        return np.zeros(1)
    if face:
        n = N + 1 if grid.topology[axis] == Bounded else N
        return o + d * np.arange(-H, n + H)
    return o + d * (np.arange(-H, N + H) + 0.5)


def make_snapshot(grid: RectilinearGrid, t: float, T_period: float, var_names: Sequence[str] = ("T", "b"),
                  velocity_names: Sequence[str] = ("u", "v", "w"), amplitude: float = 1.0, xp=np, device=None,
                  dtype=np.float32, global_grid: RectilinearGrid = None) -> Dict[str, "np.ndarray"]:
    """All fields of one snapshot at time t.  With xp=torch the arrays are torch tensors on `device`
    of shape (ez, ey, ex) (C order), i.e. the same memory as a Fortran (ex, ey, ez) parent array."""
    is_torch = xp is not np
    two_pi = 2 * math.pi
    ref = global_grid or grid           # a y-slab evaluates the GLOBAL analytic fields on its own rows
    Lx, Ly, Lz = ref.Lx, ref.Ly, ref.Lz
    ox, oy, oz = ref.origin
    ph = two_pi * t / T_period

    def ax(axis, face):
        c = _coords(grid, axis, face)
        if is_torch:
            c = xp.as_tensor(c, dtype=xp.float64, device=device)
            shape = [1, 1, 1]
            shape[2 - axis] = c.numel()
            return c.reshape(shape)
        shape = [1, 1, 1]
        shape[axis] = c.size
        return c.reshape(shape)

    sin, cos, exp, tanh = xp.sin, xp.cos, xp.exp, xp.tanh

    def rel(c, o, L):      # coordinate mapped to [0, 1) over the domain
        return (c - o) / L

    out = {}
    flat = [grid.topology[a] == Flat for a in range(3)]

    def X(face): return rel(ax(0, face), ox, Lx)
    def Y(face): return rel(ax(1, face), oy, Ly)
    def Z(face): return rel(ax(2, face), oz, Lz)

    zprof = (lambda z: cos(math.pi * z)) if not flat[2] else (lambda z: z * 0 + 1.0)
    A = amplitude
    if "u" in velocity_names:
        x, y, z = X(True), Y(False), Z(False)
        out["u"] = A * (sin(two_pi * x) * cos(two_pi * y) * zprof(z) * math.cos(ph)
                        + 0.5 * cos(2 * two_pi * y + 1.0 + 0.3 * ph) * (1 + 0 * x) * (1 + 0 * z))
    if "v" in velocity_names:
        x, y, z = X(False), Y(True), Z(False)
        out["v"] = A * (-cos(two_pi * x) * sin(two_pi * y) * zprof(z) * math.cos(ph)
                        + 0.5 * sin(2 * two_pi * x + 2.0 - 0.2 * ph) * (1 + 0 * y) * (1 + 0 * z))
    if "w" in velocity_names:
        x, y, z = X(False), Y(False), Z(True)
        wz = sin(math.pi * z) if grid.topology[2] == Bounded else sin(two_pi * z)
        out["w"] = 0.2 * A * wz * sin(two_pi * x + ph) * cos(two_pi * y)
    x, y, z = X(False), Y(False), Z(False)
    shapes = (1 + 0 * x) * (1 + 0 * y) * (1 + 0 * z)
    for q, name in enumerate(var_names):
        if q % 2 == 0:      # a blob drifting in x, stratified in z
            xs = x - 0.5 - 0.1 * math.sin(ph)
            r2 = (xs * xs) * (0 if flat[0] else 1) + ((y - 0.5) ** 2) * (0 if flat[1] else 1)
            out[name] = (exp(-r2 / (0.125 ** 2)) * (1.0 + 0.5 * z) + 0.05 * q) * shapes
        else:               # a sharp periodic front, meandering in time
            arg = sin(two_pi * (y if not flat[1] else x)) + 0.3 * sin(two_pi * x + ph)
            out[name] = (tanh(8.0 * arg) + z + 0.05 * q) * shapes
    for k in list(out.keys()):
        v = out[k]
        if is_torch:
            out[k] = v.to(xp.float32).contiguous() if dtype == np.float32 else v.contiguous()
        else:
            v = np.asarray(v, dtype=np.float64)
            if dtype == np.float32:
                v = v.astype(np.float32)
            out[k] = np.asfortranarray(v)
    return out


def make_series(grid: RectilinearGrid, times: Sequence[float], **kw):
    """InputData-style dict: name -> list of parent arrays, one per time (numpy)."""
    T_period = kw.pop("T_period", float(times[-1] - times[0]) or 1.0)
    fields: Dict[str, list] = {}
    for t in times:
        snap = make_snapshot(grid, float(t), T_period, **kw)
        for k, v in snap.items():
            fields.setdefault(k, []).append(v)
    return fields
