"""`OfflineFilterConfig` / `OnlineFilterConfig`: host-side mirrors of the reference's configuration
structs and their validating keyword constructors
(reference src/OfflineLagrangianFilter/OfflineLagrangianFilter.jl:83-517,
src/OnlineLagrangianFilter/OnlineLagrangianFilter.jl:20-275).

Field names, defaults, derived values and error conditions follow the reference.  Differences that
come from the host language: Julia `@info`/`@warn` become `logging` records on the "lfb200" logger,
`error(...)` becomes `ValueError` (or `FileNotFoundError`), the `architecture`/`backend` fields become
`device` and `n_slots` (resident snapshot slots; `InMemory(4)` -> 4), and the grid must be passed
explicitly (see grids.py).  Input data can be a JLD2 file written by Oceananigans' JLD2Writer or an
in-memory `InputData`.
"""
from __future__ import annotations

import inspect
import logging
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .filter_utils import (FilterParams, normalisation, set_offline_BW2_filter_params,
                           set_online_BW_filter_params, validate_filter_params)
from .grids import Flat, RectilinearGrid

log = logging.getLogger("lfb200")


class _Scheme:
    """Marker for an Oceananigans advection scheme (`OfflineFilterConfig.advection`, OfflineLagrangianFilter.jl:139,246).
    `scheme_name` selects the reconstruction and its buffer-scheme chain in liblfb200 (LFB_ADVECTION_*)."""
    _orders: Tuple[int, ...] = ()
    _prefix = ""

    def __init__(self, order: int):
        if order not in self._orders:
            raise ValueError(f"{type(self).__name__}(order={order}) is not available; liblfb200 builds orders {self._orders}")
        self.order = int(order)

    @property
    def scheme_name(self) -> str:
        return f"{self._prefix}{self.order}"

    @property
    def required_halo(self) -> int:
        """Cells the stencil reaches on each side of a face (Oceananigans' required_halo_size)."""
        return (self.order + 1) // 2

    def __repr__(self):
        return f"{type(self).__name__}(order={self.order})"

    def __eq__(self, other):
        return type(self) is type(other) and self.order == other.order

    def __hash__(self):
        return hash((type(self).__name__, self.order))


class WENO(_Scheme):
    """`WENO(order=5)` (the reference's default: WENO5-Z with buffer schemes WENO3-Z -> Centered2) or `WENO(order=3)`."""
    _orders, _prefix = (3, 5), "WENO"

    def __init__(self, order: int = 5):
        super().__init__(order)


class Centered(_Scheme):
    """`Centered(order=2)` / `Centered(order=4)` (buffer scheme Centered2)."""
    _orders, _prefix = (2, 4), "Centered"

    def __init__(self, order: int = 2):
        super().__init__(order)


class UpwindBiased(_Scheme):
    """`UpwindBiased(order=1 | 3 | 5)` (buffer schemes of the next lower odd order)."""
    _orders, _prefix = (1, 3, 5), "UpwindBiased"

    def __init__(self, order: int = 3):
        super().__init__(order)


@dataclass
class InputData:
    """In-memory stand-in for the original simulation's JLD2 output: `times` plus, per name, one
    parent array (halos included, Julia (x,y,z) order) per time."""
    times: np.ndarray
    fields: Dict[str, Sequence[np.ndarray]]

    def __post_init__(self):
        self.times = np.asarray(self.times, dtype=np.float64)
        for k, v in self.fields.items():
            if len(v) != len(self.times):
                raise ValueError(f"field {k!r} has {len(v)} snapshots for {len(self.times)} times")

    def has(self, name):
        return name in self.fields

    def snapshot(self, name, n):
        return self.fields[name][n]


class _JLD2Input:
    def __init__(self, path):
        from .jld2 import JLD2File
        self.file = JLD2File(path)
        self.iterations = self.file.iterations("t")
        self.times = self.file.times()

    def has(self, name):
        return f"timeseries/{name}" in self.file

    def snapshot(self, name, n):
        return self.file[f"timeseries/{name}/{self.iterations[n]}"]


def _check_relaxation(boundary_relaxation, relax_timescale, mask_params, mask_func, grid):
    """OfflineLagrangianFilter.jl:462-485 / OnlineLagrangianFilter.jl (same block)."""
    if not boundary_relaxation:
        return
    if relax_timescale is None:
        raise ValueError("A relax_timescale must be set if boundary_relaxation = true")
    if mask_params is None:
        log.warning("mask_params = nothing with boundary_relaxation = true. The mask function probably needs parameters.")
    if mask_func is None:
        raise ValueError("A spatial mask_func must be set if boundary_relaxation = true")
    n_args = len(inspect.signature(mask_func).parameters) - 1      # last argument is mask_params
    n_nonflat = sum(1 for t in grid.topology if t != Flat)
    if n_args != n_nonflat:
        raise ValueError("mask_func has the wrong number of arguments")


def _resolve_filter_params(filter_params, N, freq_c, setter, what):
    if filter_params is not None and (N is not None or freq_c is not None):
        raise ValueError("Specify either filter_params or N and freq_c, not both.")
    if filter_params is None and (N is None or freq_c is None):
        raise ValueError("Must specify either filter_params or both N and freq_c.")
    if filter_params is None:
        fp = setter(N=N, freq_c=freq_c)
        log.info("Setting filter parameters to use %s, order %s, cutoff frequency %s", what, N, freq_c)
        return fp
    return validate_filter_params(filter_params)


@dataclass
class OfflineFilterConfig:
    original_data_filename: Optional[str]
    var_names_to_filter: Tuple[str, ...]
    velocity_names: Tuple[str, ...]
    T_start: float
    T_end: float
    T: float
    T_out: float
    filter_params: FilterParams
    Δt: float
    map_to_mean: bool
    forward_output_filename: str
    backward_output_filename: str
    output_filename: str
    npad: int
    compute_mean_velocities: bool
    delete_intermediate_files: bool
    compute_Eulerian_filter: bool
    output_netcdf: bool
    output_original_data: bool
    advection: Optional[_Scheme]
    grid: RectilinearGrid
    label: str
    boundary_relaxation: bool
    relax_timescale: Optional[float]
    mask_params: Optional[dict]
    mask_func: Optional[Callable]
    # host-language replacements for `architecture` / `backend`
    device: int = 0
    n_slots: int = 4
    strict: bool = False
    kernel: str = "auto"
    timestepper: str = "RungeKutta3"       # LagrangianFilter(...; timestepper) (lagrangian_filter.jl:48,85)
    data: Optional[InputData] = field(default=None, repr=False)

    def __init__(self, *, var_names_to_filter, velocity_names, original_data_filename=None, data=None,
                 T_start=None, T_end=None, T=None, T_out=None, N=None, freq_c=None, filter_params=None, Δt=None,
                 dt=None, map_to_mean=True, forward_output_filename="forward_output.jld2",
                 backward_output_filename="backward_output.jld2", output_filename="filtered_output.jld2", npad=5,
                 compute_mean_velocities=True, delete_intermediate_files=True, compute_Eulerian_filter=False,
                 output_netcdf=False, output_original_data=True, advection="default", grid=None, label="",
                 boundary_relaxation=False, relax_timescale=None, mask_params=None, mask_func=None,
                 device=0, n_slots=4, strict=False, kernel="auto", timestepper="RungeKutta3"):
        if timestepper not in ("RungeKutta3", "QuasiAdamsBashforth2"):
            raise ValueError("timestepper must be 'RungeKutta3' or 'QuasiAdamsBashforth2'")
        if Δt is None:
            Δt = dt
        if isinstance(advection, str) and advection == "default":
            advection = WENO()
        if advection is not None and not isinstance(advection, _Scheme):
            raise ValueError("advection must be WENO(...), Centered(...), UpwindBiased(...) or None")
        var_names_to_filter = tuple(var_names_to_filter)
        velocity_names = tuple(velocity_names)
        if data is None:
            if original_data_filename is None:
                raise ValueError("original_data_filename (or data=InputData(...)) is required")
            if not os.path.isfile(original_data_filename):
                raise FileNotFoundError(f"Source file not found: {original_data_filename}")
        for vel in ("u", "v", "w"):
            if vel in var_names_to_filter:
                raise ValueError(f"Velocity variable '{vel}' cannot be in 'var_names_to_filter'. If you want to filter "
                                 "a velocity that you will also be advecting with, include it in 'velocity_names'.")
        if "t" in var_names_to_filter:
            raise ValueError("Time variable 't' cannot be in 'var_names_to_filter'.")
        for vel in velocity_names:
            if vel not in ("u", "v", "w"):
                # the reference's forcing looks velocities up by name in model.velocities (utils:544,564)
                raise ValueError(f"velocity_names entries must be 'u', 'v' or 'w', got {vel!r}")
        if map_to_mean:
            log.info("Advection for Lagrangian filtering will be performed using only velocities %s - any other "
                     "velocity components will be zero by default. Maps for regridding to mean position will be "
                     "computed corresponding to velocities: %s.", velocity_names, velocity_names)
        else:
            log.info("Advection for Lagrangian filtering will be performed using only velocities %s - any other "
                     "velocity components will be zero by default.", velocity_names)
        if compute_mean_velocities:
            log.info("Mean velocities corresponding to %s will be computed.", velocity_names)

        source = data if data is not None else _JLD2Input(original_data_filename)
        for var in var_names_to_filter + velocity_names:
            if not source.has(var):
                raise ValueError(f"Variable '{var}' not found in original data file.")
        times = np.asarray(source.times, dtype=np.float64)
        given = (T_start is not None) + (T_end is not None) + (T is not None)
        if given == 3:
            if T_start + T != T_end:
                raise ValueError("Inconsistent time specifications: T_start + T != T_end")
            if T_start > T_end:     # the reference has the typo `T-start` here (OfflineLagrangianFilter.jl:306)
                raise ValueError("Inconsistent time specifications: T_start > T_end")
        elif T_start is None:
            if T_end is not None and T is not None:
                T_start = T_end - T
            elif T_end is not None:
                T_start = times[0]; T = T_end - T_start
            elif T is not None:
                T_start = times[0]; T_end = T_start + T
            else:
                T_start = times[0]; T_end = times[-1]; T = T_end - T_start
        elif T_end is None:
            if T is None:
                T_end = times[-1]; T = T_end - T_start
            else:
                T_end = T_start + T
        else:
            T = T_end - T_start
            if T < 0:
                raise ValueError("Inconsistent time specifications: T_start > T_end")
        if T_start < times[0] or T_start > times[-1]:
            raise ValueError(f"T_start={T_start} is outside the range of the original data: [{times[0]}, {times[-1]}].")
        if T_end < times[0] or T_end > times[-1]:
            raise ValueError(f"T_end={T_end} is outside the range of the original data: [{times[0]}, {times[-1]}].")
        log.info("Filter interval will be from T_start=%s to T_end=%s, duration T=%s", T_start, T_end, T)
        if T_out is None:
            T_out = times[1] - times[0]
            log.info("T_out not set. Setting T_out = %s", T_out)
        if Δt is None:
            Δt = T_out / 10
            log.info("Δt (filter simulation timestep) not set. Setting Δt = %s, but be careful", Δt)

        fp = _resolve_filter_params(filter_params, N, freq_c, set_offline_BW2_filter_params, "Butterworth squared")
        if fp["N_coeffs"] == 0.5:
            if not math.isclose(fp["a1"] * 2, fp["c1"], rel_tol=1.5e-8):
                log.warning("Filter coefficients are not normalised: 2*a1=%s != c1=%s. You can continue, but you should "
                            "consider setting `map_to_mean=false` as the map may be meaningless.", 2 * fp["a1"], fp["c1"])
        else:
            s = normalisation(fp)
            if not math.isclose(s, 0.5, rel_tol=1.5e-8):
                log.warning("Filter coefficients are not normalised: %s != 0.5. You can continue, but you should consider "
                            "setting `map_to_mean=false` as the map may be meaningless.", s)
        if grid is None:
            raise ValueError("grid=RectilinearGrid(...) is required: this host does not decode the Julia-serialised "
                             "grid stored in the JLD2 file")
        if hasattr(grid, "immersed_boundary") and map_to_mean:
            # OfflineLagrangianFilter.jl:428-431
            log.warning("The final interpolation to mean position does not yet work well with immersed boundaries - "
                        "consider setting map_to_mean=false")
        if advection is None and map_to_mean:
            log.warning("Advection scheme is 'nothing' (Eulerian filter) so setting map_to_mean=false")
            map_to_mean = False
        elif advection is None:
            log.info("Advection scheme is 'nothing' so the Eulerian (not Lagrangian) filter will be computed.")
        if compute_Eulerian_filter and advection is None:
            log.warning("compute_Eulerian_filter=true and advection is 'nothing' - Eulerian filter will be computed "
                        "twice, so you should probably set compute_Eulerian_filter=false.")
            compute_Eulerian_filter = False
        _check_relaxation(boundary_relaxation, relax_timescale, mask_params, mask_func, grid)
        if n_slots < 2:
            raise ValueError("n_slots must be at least 2")

        self.original_data_filename = original_data_filename
        self.var_names_to_filter = var_names_to_filter
        self.velocity_names = velocity_names
        self.T_start, self.T_end, self.T = float(T_start), float(T_end), float(T)
        self.T_out = float(T_out)
        self.filter_params = fp
        self.Δt = float(Δt)
        self.map_to_mean = bool(map_to_mean)
        self.forward_output_filename = forward_output_filename
        self.backward_output_filename = backward_output_filename
        self.output_filename = output_filename
        self.npad = int(npad)
        self.compute_mean_velocities = bool(compute_mean_velocities)
        self.delete_intermediate_files = bool(delete_intermediate_files)
        self.compute_Eulerian_filter = bool(compute_Eulerian_filter)
        self.output_netcdf = bool(output_netcdf)
        self.output_original_data = bool(output_original_data)
        self.advection = advection
        self.grid = grid
        self.label = label
        self.boundary_relaxation = bool(boundary_relaxation)
        self.relax_timescale = relax_timescale
        self.mask_params = mask_params
        self.mask_func = mask_func
        self.device, self.n_slots, self.strict, self.kernel = int(device), int(n_slots), bool(strict), kernel
        self.timestepper = timestepper
        self.data = data
        self._source = source

    @property
    def dt(self):
        return self.Δt

    def source(self):
        return self._source


@dataclass
class OnlineFilterConfig:
    """Mirror of OnlineLagrangianFilter.jl:20-34.  The online filter tracers live inside the user's own
    Oceananigans model and are stepped by it, so this host only keeps the
    configuration object and the name / forcing constructors; there is no device path behind it."""
    grid: RectilinearGrid
    var_names_to_filter: Tuple[str, ...]
    velocity_names: Tuple[str, ...]
    filter_params: FilterParams
    map_to_mean: bool
    compute_mean_velocities: bool
    output_filename: str
    npad: int
    label: str
    boundary_relaxation: bool
    relax_timescale: Optional[float]
    mask_params: Optional[dict]
    mask_func: Optional[Callable]

    def __init__(self, *, grid, var_names_to_filter, velocity_names, N=None, freq_c=None, filter_params=None,
                 map_to_mean=True, compute_mean_velocities=True, output_filename="online_filtered_output.jld2", npad=5,
                 label="", boundary_relaxation=False, relax_timescale=None, mask_params=None, mask_func=None):
        var_names_to_filter = tuple(var_names_to_filter)
        velocity_names = tuple(velocity_names)
        for vel in ("u", "v", "w"):
            if vel in var_names_to_filter:
                raise ValueError(f"Velocity variable '{vel}' cannot be in 'var_names_to_filter'.")
        fp = _resolve_filter_params(filter_params, N, freq_c, set_online_BW_filter_params, "Butterworth")
        _check_relaxation(boundary_relaxation, relax_timescale, mask_params, mask_func, grid)
        self.grid = grid
        self.var_names_to_filter, self.velocity_names = var_names_to_filter, velocity_names
        self.filter_params = fp
        self.map_to_mean, self.compute_mean_velocities = bool(map_to_mean), bool(compute_mean_velocities)
        self.output_filename, self.npad, self.label = output_filename, int(npad), label
        self.boundary_relaxation, self.relax_timescale = bool(boundary_relaxation), relax_timescale
        self.mask_params, self.mask_func = mask_params, mask_func
