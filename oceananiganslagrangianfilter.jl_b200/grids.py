"""Regular rectilinear grid description (the subset of Oceananigans' ``RectilinearGrid`` the filter uses).

The reference takes the grid from ``FieldTimeSeries(...).grid`` (OfflineLagrangianFilter.jl:425-426) or the
``grid`` keyword; its JLD2 serialisation is a Julia struct this host does not decode, so callers pass the
grid explicitly, with the same keywords Oceananigans uses::

    RectilinearGrid(size=(10, 10), x=(-5000, 5000), z=(-100, 0), topology=("Periodic", "Flat", "Bounded"))
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

Periodic, Bounded, Flat = "Periodic", "Bounded", "Flat"
_TOPOLOGIES = (Periodic, Bounded, Flat)


@dataclass(frozen=True)
class RectilinearGrid:
    Nx: int
    Ny: int
    Nz: int
    Hx: int
    Hy: int
    Hz: int
    topology: Tuple[str, str, str]
    origin: Tuple[float, float, float]     # first face coordinate per axis (0 on Flat axes)
    Lx: float
    Ly: float
    Lz: float

    def __init__(self, size: Sequence[int], topology: Sequence[str] = (Periodic, Periodic, Bounded),
                 x: Optional[Tuple[float, float]] = None, y: Optional[Tuple[float, float]] = None,
                 z: Optional[Tuple[float, float]] = None, extent: Optional[Sequence[float]] = None,
                 halo: Optional[Sequence[int]] = None):
        topology = tuple(topology)
        if len(topology) != 3 or any(t not in _TOPOLOGIES for t in topology):
            raise ValueError(f"topology must be three of {_TOPOLOGIES}, got {topology}")
        nonflat = [ax for ax in range(3) if topology[ax] != Flat]
        size = tuple(int(s) for s in size)
        if len(size) != len(nonflat):
            raise ValueError(f"size must have one entry per non-Flat dimension ({len(nonflat)}), got {size}")
        if halo is None:
            halo = (3,) * len(nonflat)          # Oceananigans' default halo
        halo = tuple(int(h) for h in halo)
        if len(halo) != len(nonflat):
            raise ValueError("halo must have one entry per non-Flat dimension")
        bounds = {0: x, 1: y, 2: z}
        if extent is not None:
            extent = tuple(float(e) for e in extent)
            if len(extent) != len(nonflat):
                raise ValueError("extent must have one entry per non-Flat dimension")
            for n, ax in enumerate(nonflat):
                if bounds[ax] is not None:
                    raise ValueError("give either extent or x/y/z, not both")
                # Oceananigans: extent puts x,y in (0, L) and z in (-Lz, 0)
                bounds[ax] = (-extent[n], 0.0) if ax == 2 else (0.0, extent[n])
        N, H, L, O = [1, 1, 1], [0, 0, 0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]
        for n, ax in enumerate(nonflat):
            if bounds[ax] is None:
                raise ValueError(f"{'xyz'[ax]} interval (or extent) required for the non-Flat {'xyz'[ax]} axis")
            lo, hi = float(bounds[ax][0]), float(bounds[ax][1])
            if not hi > lo:
                raise ValueError(f"{'xyz'[ax]} interval must be increasing")
            N[ax], H[ax], L[ax], O[ax] = size[n], halo[n], hi - lo, lo
            if size[n] < 1:
                raise ValueError("grid sizes must be positive")
        for ax in range(3):
            if topology[ax] == Flat and bounds[ax] is not None:
                raise ValueError(f"Flat axis {'xyz'[ax]} takes no interval")
        s = object.__setattr__
        s(self, "Nx", N[0]); s(self, "Ny", N[1]); s(self, "Nz", N[2])
        s(self, "Hx", H[0]); s(self, "Hy", H[1]); s(self, "Hz", H[2])
        s(self, "topology", topology); s(self, "origin", tuple(O))
        s(self, "Lx", L[0]); s(self, "Ly", L[1]); s(self, "Lz", L[2])

    # ------------------------------------------------------------------ derived
    @property
    def N(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def H(self):
        return (self.Hx, self.Hy, self.Hz)

    @property
    def L(self):
        return (self.Lx, self.Ly, self.Lz)

    @property
    def spacing(self):
        """(dx, dy, dz); 1.0 on Flat axes, as Oceananigans' operators treat them."""
        return tuple(1.0 if self.topology[ax] == Flat else self.L[ax] / self.N[ax] for ax in range(3))

    @property
    def nonflat_axes(self):
        return [ax for ax in range(3) if self.topology[ax] != Flat]

    def center_shape(self):
        return tuple(self.N[ax] + 2 * self.H[ax] for ax in range(3))

    def face_shape(self, axis: int):
        """Parent shape of a field on the faces of `axis`: one extra entry on a Bounded axis (on a y-slab only
        the last rank holds the wall face)."""
        s = list(self.center_shape())
        rank, nranks = getattr(self, "_slab", (0, 1))
        if self.topology[axis] == Bounded and not (axis == 1 and rank != nranks - 1):
            s[axis] += 1
        return tuple(s)

    def nodes(self, axis: int, with_halos: bool = True) -> np.ndarray:
        """Cell-centre coordinates (xnodes/ynodes/znodes(grid, Center(); with_halos))."""
        N, H = self.N[axis], self.H[axis] if with_halos else 0
        d = self.spacing[axis]
        return self.origin[axis] + d * (np.arange(1 - H, N + H + 1) - 0.5)

    def slab(self, rank: int, nranks: int) -> "RectilinearGrid":
        """The y-slab of rank `rank` out of `nranks` equal slabs (multi-GPU decomposition)."""
        if nranks == 1:
            return self
        if self.topology[1] == Flat or self.Ny % nranks:
            raise ValueError(f"Ny={self.Ny} cannot be split into {nranks} equal y-slabs")
        ny = self.Ny // nranks
        dy = self.Ly / self.Ny
        g = object.__new__(RectilinearGrid)
        for k, v in self.__dict__.items():
            object.__setattr__(g, k, v)
        object.__setattr__(g, "Ny", ny)
        object.__setattr__(g, "Ly", dy * ny)
        o = list(self.origin)
        o[1] = self.origin[1] + dy * ny * rank
        object.__setattr__(g, "origin", tuple(o))
        object.__setattr__(g, "_slab", (rank, nranks))
        return g
