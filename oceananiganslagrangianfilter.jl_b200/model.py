"""`LagrangianFilter`: the device-resident model behind the C ABI.

Stands in for the reference's `LagrangianFilter <: AbstractModel` (reference
src/OfflineLagrangianFilter/lagrangian_filter.jl:23-180): tracers = the filter variables, prescribed
velocities and original variables = snapshots resident on the GPU, RK3 time stepping.  Every method
is a thin call into liblfb200.so; nothing here computes on field data.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check
from .filter_utils import FilterParams, create_filtered_vars, tracer_layout
from .grids import ImmersedBoundaryGrid, RectilinearGrid

_KERNELS = {"auto": _lib.KERNEL_AUTO, "generic": _lib.KERNEL_GENERIC, "ztma": _lib.KERNEL_ZTMA}


class LagrangianFilter:
    def __init__(self, grid: RectilinearGrid, *, var_names_to_filter: Sequence[str], velocity_names: Sequence[str],
                 filter_params: FilterParams, with_maps: bool = True, advection=True, n_slots: int = 4,
                 device: int = 0, strict: bool = False, kernel: str = "auto", label: str = "",
                 relax_timescale: Optional[float] = None, mask: Optional[np.ndarray] = None, config=None,
                 timestepper: str = "RungeKutta3", chi: float = 0.1):
        """advection: True (the default WENO(order=5)), False / None (advection = nothing), a config.WENO / Centered /
        UpwindBiased marker or a scheme name of _lib.ADVECTION; timestepper: "RungeKutta3" | "QuasiAdamsBashforth2"
        (lagrangian_filter.jl:48,85); grid: RectilinearGrid or ImmersedBoundaryGrid."""
        self.L = _lib.load()
        self.grid = grid
        self.var_names = tuple(var_names_to_filter)
        self.velocity_names = tuple(velocity_names)
        self.velocities_present = [v for v in ("u", "v", "w") if v in self.velocity_names]
        self.filter_params = filter_params
        self.with_maps = bool(with_maps)
        d = _lib.Desc()
        for ax in range(3):
            d.size[ax], d.halo[ax] = grid.N[ax], grid.H[ax]
            d.topology[ax] = _lib.TOPOLOGY[grid.topology[ax]]
            d.spacing[ax] = grid.spacing[ax]
        d.n_vars = len(self.var_names)
        d.velocity_mask = sum(1 << ax for ax, v in enumerate("uvw") if v in self.velocity_names)
        d.n_coeffs = 0 if filter_params.single_exponential else filter_params.n_sets
        a, b, c, dd = filter_params.coefficient_arrays()
        for i in range(filter_params.n_sets):
            d.a[i], d.b[i], d.c[i], d.d[i] = a[i], b[i], c[i], dd[i]
        d.with_maps = int(self.with_maps)
        if advection is True:
            advection = "WENO5"
        elif advection is False:
            advection = None
        elif hasattr(advection, "scheme_name"):
            advection = advection.scheme_name
        if advection not in _lib.ADVECTION:
            raise ValueError(f"unknown advection scheme {advection!r}; one of {sorted(k for k in _lib.ADVECTION if k)}")
        d.advection = _lib.ADVECTION[advection]
        if timestepper not in _lib.TIMESTEPPER:
            raise ValueError(f"timestepper must be one of {sorted(_lib.TIMESTEPPER)}")
        d.timestepper, d.chi = _lib.TIMESTEPPER[timestepper], float(chi)
        immersed = isinstance(grid, ImmersedBoundaryGrid)
        d.immersed = int(immersed)
        d.n_slots, d.device, d.strict, d.kernel = int(n_slots), int(device), int(bool(strict)), _KERNELS[kernel]
        d.relaxation = int(relax_timescale is not None)
        d.relax_timescale = float(relax_timescale or 0.0)
        self.n_slots = int(n_slots)
        self.rank, self.nranks = 0, 1
        self._in_flight, self._pinned = [], []
        h = C.c_void_p()
        check(self.L.lfb_create(C.byref(d), C.byref(h)))
        self.h = h
        self.center_shape = grid.center_shape()
        # reference tracer names -> device tracer index
        cfg = config or _NameConfig(self.var_names, self.velocity_names, filter_params, with_maps, label)
        self.tracer_names = create_filtered_vars(cfg)
        self._tracer_index: Dict[str, int] = {}
        for name, (f, p, s) in tracer_layout(cfg).items():
            self._tracer_index[name] = check(self.L.lfb_tracer_index(self.h, f, p, s))
        if mask is not None:
            m = _lib.as_parent(mask, self.center_shape)
            check(self.L.lfb_set_mask(self.h, m.ctypes.data))
        if immersed:
            flags = _lib.as_parent(grid.immersed_flags(), self.center_shape)
            heights = grid.cell_heights()
            hp = _lib.as_parent(heights, self.center_shape) if heights is not None else None
            check(self.L.lfb_set_immersed(self.h, flags.ctypes.data, hp.ctypes.data if hp is not None else None))

    # ------------------------------------------------------------------ lifecycle
    def free_pinned(self, since: int = 0):
        """Releases the page-locked arrays handed out by pinned_empty (they must not be used afterwards): all of
        them, or those allocated after the first `since` ones."""
        for p in self._pinned[since:]:
            self.L.lfb_host_free(p)
        self._pinned = self._pinned[:since]

    @property
    def n_pinned(self) -> int:
        return len(self._pinned)

    def close(self):
        if getattr(self, "h", None):
            self.L.lfb_wait_transfers(self.h)
            self.free_pinned()
            self.L.lfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def sync(self):
        check(self.L.lfb_sync(self.h))

    # ------------------------------------------------------------------ input series
    def field_id(self, name: str) -> int:
        if name in ("u", "v", "w"):
            return "uvw".index(name)
        return _lib.FIELD_VAR0 + self.var_names.index(name)

    def field_shape(self, name: str):
        return self.grid.face_shape("uvw".index(name)) if name in ("u", "v", "w") else self.center_shape

    def upload_snapshot(self, slot: int, name: str, parent: np.ndarray, time: float, wait: bool = True):
        """One field of one input snapshot (update_input_data! source data, utils:1036-1058).  Float32
        arrays (the reference's on-disk precision) are sent as Float32 and widened on the device.
        With wait=False the call returns once the copy is queued; the array is kept alive until
        wait_transfers() and the copy overlaps with queued steps if it is page-locked (pinned_empty)."""
        shape = self.field_shape(name)
        f32 = np.asarray(parent).dtype == np.float32
        a = _lib.as_parent(parent, shape, np.float32 if f32 else np.float64)
        if not wait:
            self._in_flight.append(a)
            check(self.L.lfb_upload_snapshot_async(self.h, slot, self.field_id(name), a.ctypes.data, int(f32), float(time)))
        elif f32:
            check(self.L.lfb_upload_snapshot_f32(self.h, slot, self.field_id(name), a.ctypes.data, float(time)))
        else:
            check(self.L.lfb_upload_snapshot(self.h, slot, self.field_id(name), a.ctypes.data, float(time)))

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """Fortran-ordered array in page-locked host memory (for asynchronous uploads / outputs)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        ptr = C.c_void_p()
        check(self.L.lfb_host_alloc(max(1, n * dtype.itemsize), C.byref(ptr)))
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order="F")
        self._pinned.append(ptr)
        return arr

    def output_async(self, kind: int, field: int, out: np.ndarray):
        """Queue an output into `out` (Fortran order, parent shape, float32 or float64); valid after wait_transfers()."""
        if out.shape != self.center_shape or not out.flags.f_contiguous or out.dtype not in (np.float32, np.float64):
            raise ValueError("output array must be a Fortran-ordered float32/float64 array of the parent shape")
        self._in_flight.append(out)
        check(self.L.lfb_output_async(self.h, kind, field, out.ctypes.data, int(out.dtype == np.float32)))
        return out

    def wait_transfers(self):
        check(self.L.lfb_wait_transfers(self.h))
        self._in_flight.clear()

    def set_slot_time(self, slot: int, time: float):
        check(self.L.lfb_set_slot_time(self.h, slot, float(time)))

    def invalidate_slot(self, slot: int):
        check(self.L.lfb_invalidate_slot(self.h, slot))

    def set_direction(self, direction: int):
        check(self.L.lfb_set_direction(self.h, int(direction)))

    # ------------------------------------------------------------------ state
    def init_from_data(self):
        check(self.L.lfb_init_from_data(self.h))

    def negate_maps(self):
        check(self.L.lfb_negate_maps(self.h))

    def reset_clock(self):
        check(self.L.lfb_reset_clock(self.h))

    @property
    def time(self) -> float:
        return self.L.lfb_time(self.h)

    @property
    def iteration(self) -> int:
        return self.L.lfb_iteration(self.h)

    @property
    def n_tracers(self) -> int:
        return self.L.lfb_n_tracers(self.h)

    def tracer_index(self, name: str) -> int:
        return self._tracer_index[name]

    def tracer(self, name_or_index) -> np.ndarray:
        t = name_or_index if isinstance(name_or_index, int) else self._tracer_index[name_or_index]
        out = np.empty(self.center_shape, dtype=np.float64, order="F")
        check(self.L.lfb_get_tracer(self.h, t, out.ctypes.data))
        return out

    def set_tracer(self, name_or_index, parent):
        t = name_or_index if isinstance(name_or_index, int) else self._tracer_index[name_or_index]
        a = _lib.as_parent(parent, self.center_shape)
        check(self.L.lfb_set_tracer(self.h, t, a.ctypes.data))

    def tendency(self, name_or_index) -> np.ndarray:
        t = name_or_index if isinstance(name_or_index, int) else self._tracer_index[name_or_index]
        out = np.empty(self.center_shape, dtype=np.float64, order="F")
        check(self.L.lfb_get_tendency(self.h, t, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ stepping
    def step(self, dt: float):
        check(self.L.lfb_step(self.h, float(dt)))

    def run_until(self, t_stop: float, dt: float, align_intervals: Sequence[float] = ()):
        arr = (C.c_double * max(1, len(align_intervals)))(*[float(x) for x in align_intervals])
        check(self.L.lfb_run_until(self.h, float(t_stop), float(dt), len(align_intervals), arr))

    # ------------------------------------------------------------------ outputs
    def output(self, kind: int, field: int, dtype=np.float64) -> np.ndarray:
        out = np.empty(self.center_shape, dtype=dtype, order="F")
        if dtype == np.float32:
            check(self.L.lfb_output_f32(self.h, kind, field, out.ctypes.data))
        else:
            check(self.L.lfb_output(self.h, kind, field, out.ctypes.data))
        return out

    def combine(self, fwd, bwd_lo, bwd_hi, w_hi: float, sign: float) -> np.ndarray:
        """fwd + sign * lerp(bwd_lo, bwd_hi; w_hi) on the device (post:135-164)."""
        f = np.asfortranarray(fwd, dtype=np.float64)
        lo = np.asfortranarray(bwd_lo, dtype=np.float64)
        hi = lo if bwd_hi is None else np.asfortranarray(bwd_hi, dtype=np.float64)
        out = np.empty(f.shape, dtype=np.float64, order="F")
        check(self.L.lfb_combine(self.h, f.size, f.ctypes.data, lo.ctypes.data, hi.ctypes.data, float(w_hi),
                                 float(sign), out.ctypes.data))
        return out

    # ------------------------------------------------------------------ device-resident outputs
    def output_keep(self, kind: int, field: int, dtype=np.float32) -> int:
        """Evaluate an output now and keep it on the device (rounded to Float32 like the reference's JLD2Writer when
        dtype is float32); returns its id."""
        i = C.c_int32()
        check(self.L.lfb_output_keep(self.h, kind, field, int(np.dtype(dtype) == np.float32), C.byref(i)))
        return i.value

    def kept_combine(self, fwd: int, bwd_lo: int, bwd_hi: Optional[int], w_hi: float, sign: float) -> int:
        i = C.c_int32()
        check(self.L.lfb_kept_combine(self.h, fwd, bwd_lo, bwd_lo if bwd_hi is None else bwd_hi, float(w_hi), float(sign),
                                      C.byref(i)))
        return i.value

    def kept_info(self, kid: int, pointer: bool = False):
        """(dtype, parent shape[, device pointer]) of a kept output."""
        f32, ext, ptr = C.c_int32(), (C.c_int32 * 3)(), C.c_void_p()
        check(self.L.lfb_kept_info(self.h, kid, C.byref(f32), C.byref(ext), C.byref(ptr) if pointer else None))
        dtype = np.float32 if f32.value else np.float64
        return (dtype, tuple(ext), ptr.value) if pointer else (dtype, tuple(ext))

    def kept_fetch(self, kid: int, alloc=None, wait: bool = True) -> np.ndarray:
        """Host copy of a kept output; `alloc(shape, dtype)` supplies the (Fortran-ordered) destination array.  With
        wait=False the copy is only queued (download stream): the array is valid after wait_transfers()."""
        dtype, shape = self.kept_info(kid)
        out = np.empty(shape, dtype=dtype, order="F") if alloc is None else alloc(shape, dtype)
        if out.shape != shape or out.dtype != dtype or not out.flags.f_contiguous:
            raise ValueError("alloc must return a Fortran-ordered array of the requested shape and dtype")
        if wait:
            check(self.L.lfb_kept_fetch(self.h, kid, out.ctypes.data))
        else:
            self._in_flight.append(out)
            check(self.L.lfb_kept_fetch_async(self.h, kid, out.ctypes.data))
        return out

    def kept_release(self, kid: int):
        check(self.L.lfb_kept_release(self.h, kid))

    def kept_reserve(self, count: int, dtype=np.float32, shape=None):
        """Allocate `count` blocks for kept outputs of this dtype (parent shape) ahead of time."""
        n = int(np.prod(shape or self.center_shape)) * np.dtype(dtype).itemsize
        check(self.L.lfb_kept_reserve(self.h, n, int(count)))

    def kept_trim(self):
        """Return the memory of released kept outputs (pooled for reuse) to the device."""
        check(self.L.lfb_kept_trim(self.h))

    def kept_gather(self, kid: int, root: int) -> int:
        """Collective: the global parent array of a kept output on rank `root` (its id there, -1 elsewhere)."""
        g = C.c_int32()
        check(self.L.lfb_kept_gather(self.h, kid, int(root), C.byref(g)))
        return g.value

    def mem_info(self):
        f, t = C.c_size_t(), C.c_size_t()
        check(self.L.lfb_mem_info(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    # ------------------------------------------------------------------ multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(self.L.lfb_comm_init(self.h, buf, int(rank), int(nranks)))
        self.rank, self.nranks = int(rank), int(nranks)

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        check(_lib.load().lfb_comm_unique_id(buf))
        return bytes(buf)

    # ------------------------------------------------------------------ introspection
    @property
    def kernel_launches(self) -> int:
        return self.L.lfb_kernel_launches(self.h)

    def stage_time_ms(self, reset: bool = False):
        ms, n = C.c_double(), C.c_int64()
        check(self.L.lfb_stage_time_ms(self.h, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timer_start(self):
        check(self.L.lfb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        check(self.L.lfb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    @property
    def kernel_name(self) -> str:
        return self.L.lfb_kernel_name(self.h).decode()


class _NameConfig:
    """Just enough of a config for the naming helpers."""

    def __init__(self, var_names, velocity_names, filter_params, with_maps, label):
        self.var_names_to_filter = tuple(var_names)
        self.velocity_names = tuple(velocity_names)
        self.filter_params = filter_params
        self.map_to_mean = bool(with_maps)
        self.compute_mean_velocities = bool(with_maps)
        self.label = label
