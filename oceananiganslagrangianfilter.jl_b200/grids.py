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


# ---------------------------------------------------------------------------------------------- immersed boundary
class GridFittedBottom:
    """Oceananigans' `GridFittedBottom(bottom_height)`: whole cells whose centre lies at or below the bottom are
    immersed.  `bottom` is a number, a callable of the horizontal non-Flat coordinates (`bottom(x)` / `bottom(x, y)`)
    or an (Nx, Ny) array of interior values."""

    def __init__(self, bottom):
        self.bottom = bottom


class PartialCellBottom(GridFittedBottom):
    """Oceananigans' `PartialCellBottom(bottom_height; minimum_fractional_cell_height = 0.2)` as used by the
    reference's lee-wave example (examples/offline_filter_lee_wave.jl:53): a cell is immersed when its TOP face lies
    at or below the bottom, and the bottom is capped so that the lowest wet cell keeps at least the minimum fraction
    of its height.  The partial heights enter the filter PDE through the face areas and cell volumes of the flux
    divergence (`cell_heights`)."""

    def __init__(self, bottom, minimum_fractional_cell_height: float = 0.2):
        super().__init__(bottom)
        self.minimum_fractional_cell_height = float(minimum_fractional_cell_height)


class ImmersedBoundaryGrid:
    """`ImmersedBoundaryGrid(underlying_grid, immersed_boundary)`: the underlying regular grid plus a bottom.  It
    answers everything RectilinearGrid answers (attribute access is forwarded), and evaluates on the host what the
    device needs: the immersed-cell flags (`immersed_flags`) and, for partial cells, the cell heights."""

    def __init__(self, underlying_grid: RectilinearGrid, immersed_boundary: GridFittedBottom):
        if underlying_grid.topology[2] == Flat:
            raise ValueError("an immersed bottom needs a non-Flat z axis")
        self.underlying_grid = underlying_grid
        self.immersed_boundary = immersed_boundary
        self._parent = None

    def __getattr__(self, name):
        if name in ("underlying_grid", "immersed_boundary", "_parent"):
            raise AttributeError(name)
        return getattr(self.underlying_grid, name)

    def slab(self, rank: int, nranks: int) -> "ImmersedBoundaryGrid":
        """The y-slab of rank `rank`.  Its flags / cell heights are cut out of the GLOBAL arrays, so that the halo rows
        of a slab hold the neighbour slabs' values (and the periodic images of the global grid at its outer edges)."""
        if nranks == 1:
            return self
        g = ImmersedBoundaryGrid(self.underlying_grid.slab(rank, nranks), self.immersed_boundary)
        g._parent = (self, rank, nranks)
        return g

    def _cut(self, name: str):
        """`name`() of the global grid, cut down to this slab (None stays None)."""
        parent, rank, nranks = self._parent
        a = getattr(parent, name)()
        if a is None:
            return None
        ny, H = parent.underlying_grid.Ny // nranks, parent.underlying_grid.Hy
        return np.asfortranarray(a[:, rank * ny:(rank + 1) * ny + 2 * H, :])

    # ------------------------------------------------------------------ host evaluation
    def bottom_height(self) -> np.ndarray:
        """Bottom height on the interior horizontal cells, shape (Nx, Ny), as given (before any capping)."""
        g, b = self.underlying_grid, self.immersed_boundary.bottom
        if callable(b):
            hax = [ax for ax in (0, 1) if g.topology[ax] != Flat]
            coords = [g.nodes(ax, with_halos=False) for ax in hax]
            mesh = np.meshgrid(*coords, indexing="ij") if coords else []
            vals = np.vectorize(lambda *xs: float(b(*xs)))(*mesh) if coords else np.asarray(float(b()))
            shape = [1, 1]
            for n, ax in enumerate(hax):
                shape[ax] = len(coords[n])
            return np.asarray(vals, dtype=np.float64).reshape(shape)
        if np.isscalar(b):
            return np.full((g.Nx, g.Ny), float(b))
        a = np.asarray(b, dtype=np.float64).reshape(g.Nx, g.Ny)
        return a

    def _fill_periodic(self, a: np.ndarray) -> np.ndarray:
        g = self.underlying_grid
        for ax in range(3):
            H, N = g.H[ax], g.N[ax]
            if H == 0 or g.topology[ax] != Periodic:
                continue
            lo = [slice(None)] * 3; hi = [slice(None)] * 3; slo = [slice(None)] * 3; shi = [slice(None)] * 3
            lo[ax] = slice(0, H); slo[ax] = slice(N, N + H)
            hi[ax] = slice(N + H, N + 2 * H); shi[ax] = slice(H, 2 * H)
            a[tuple(lo)] = a[tuple(slo)]
            a[tuple(hi)] = a[tuple(shi)]
        return a

    def _adjusted_bottom(self):
        """(zb, z_faces): the bottom height the immersed test uses, per interior column."""
        g, ib = self.underlying_grid, self.immersed_boundary
        zb = self.bottom_height()
        dz = g.spacing[2]
        zf = g.origin[2] + dz * np.arange(g.Nz + 1)                 # faces k = 1 .. Nz+1
        if isinstance(ib, PartialCellBottom):
            eps = ib.minimum_fractional_cell_height
            zb = zb.copy()
            for k in range(g.Nz):
                bottom_cell = (zf[k] <= zb) & (zf[k + 1] >= zb)
                zb = np.where(bottom_cell, np.minimum(zf[k + 1] - eps * dz, zb), zb)
        return zb, zf

    def immersed_flags(self) -> np.ndarray:
        """Parent array (centre, halos included) of 0 / 1 flags: 1 inside the immersed boundary.  Periodic halos
        carry the periodic images, the halos of Bounded axes stay 0 (no stencil test reads them)."""
        if getattr(self, "_parent", None):
            return self._cut("immersed_flags")
        g, ib = self.underlying_grid, self.immersed_boundary
        zb, zf = self._adjusted_bottom()
        if isinstance(ib, PartialCellBottom):
            ztest = zf[1:]                                             # top face of cell k
        else:
            ztest = 0.5 * (zf[:-1] + zf[1:])                           # cell centre
        inner = (ztest[None, None, :] <= zb[:, :, None]).astype(np.float64)
        out = np.zeros(g.center_shape(), dtype=np.float64, order="F")
        out[g.Hx:g.Hx + g.Nx, g.Hy:g.Hy + g.Ny, g.Hz:g.Hz + g.Nz] = inner
        return np.asfortranarray(self._fill_periodic(out))

    def cell_heights(self):
        """Parent array of the cell heights dz(i,j,k) for PartialCellBottom (None for whole cells): the lowest wet
        cell of a column reaches from the (capped) bottom to its top face."""
        if getattr(self, "_parent", None):
            return self._cut("cell_heights")
        g, ib = self.underlying_grid, self.immersed_boundary
        if not isinstance(ib, PartialCellBottom):
            return None
        zb, zf = self._adjusted_bottom()
        dz = g.spacing[2]
        bottom_cell = (zf[None, None, :-1] <= zb[:, :, None]) & (zf[None, None, 1:] >= zb[:, :, None])
        inner = np.where(bottom_cell, zf[None, None, 1:] - zb[:, :, None], dz)
        out = np.full(g.center_shape(), dz, dtype=np.float64, order="F")
        out[g.Hx:g.Hx + g.Nx, g.Hy:g.Hy + g.Ny, g.Hz:g.Hz + g.Nz] = inner
        return np.asfortranarray(self._fill_periodic(out))

    def below_bottom_mask(self) -> np.ndarray:
        """Boolean parent-shaped array of the cells the reference's `_mask_immersed` blanks before the remap
        (post:285-298): z_centre < bottom_height (the stored, i.e. capped / grid-fitted, bottom).  For whole cells
        this is the set of immersed cells; with partial cells it also holds the wet bottom cells whose centre lies
        below the bottom."""
        if getattr(self, "_parent", None):
            return self._cut("below_bottom_mask")
        g, ib = self.underlying_grid, self.immersed_boundary
        if not isinstance(ib, PartialCellBottom):
            return self.immersed_flags() != 0
        zb, zf = self._adjusted_bottom()
        zc = 0.5 * (zf[:-1] + zf[1:])
        inner = zc[None, None, :] < zb[:, :, None]
        out = np.zeros(g.center_shape(), dtype=bool, order="F")
        out[g.Hx:g.Hx + g.Nx, g.Hy:g.Hy + g.Ny, g.Hz:g.Hz + g.Nz] = inner
        return out
