"""ctypes binding of liblfb200.so (the C ABI in include/lfb200.h).

This is the Python twin of the `ccall` stubs a Julia maintainer would add.  There is
no fallback: if the shared library is missing or cannot be loaded, importing the compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LFB200_LIB") or os.path.join(HERE, "liblfb200.so")      # override: kernel experiments only

MAX_COEFFS, MAX_VARS = 8, 8
PERIODIC, BOUNDED, FLAT = 0, 1, 2
TOPOLOGY = {"Periodic": PERIODIC, "Bounded": BOUNDED, "Flat": FLAT}
ADVECTION_NONE, ADVECTION_WENO5 = 0, 1
# scheme name -> LFB_ADVECTION_* (config.WENO / Centered / UpwindBiased carry these names)
ADVECTION = {None: 0, "WENO5": 1, "Centered2": 2, "Centered4": 3, "UpwindBiased1": 4, "UpwindBiased3": 5,
             "UpwindBiased5": 6, "WENO3": 7}
TIMESTEPPER = {"RungeKutta3": 0, "QuasiAdamsBashforth2": 1}
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_ZTMA = 0, 1, 4
FIELD_U, FIELD_V, FIELD_W, FIELD_VAR0 = 0, 1, 2, 3
OUT_FILTERED_VAR, OUT_XI, OUT_MEAN_VEL = 0, 1, 2
ERR_NAMES = {-1: "LFB_ERR_ARG", -2: "LFB_ERR_CUDA", -3: "LFB_ERR_STATE", -4: "LFB_ERR_NO_DATA", -5: "LFB_ERR_COMM"}

# every symbol include/lfb200.h declares (tests check the library exports all of them)
EXPORTS = [
    "lfb_version", "lfb_last_error", "lfb_create", "lfb_destroy", "lfb_sync", "lfb_n_tracers", "lfb_tracer_index",
    "lfb_parent_len", "lfb_upload_snapshot", "lfb_upload_snapshot_f32", "lfb_upload_snapshot_async", "lfb_invalidate_slot", "lfb_set_slot_time",
    "lfb_set_direction", "lfb_set_mask", "lfb_set_immersed", "lfb_init_from_data", "lfb_negate_maps", "lfb_reset_clock", "lfb_time",
    "lfb_iteration", "lfb_get_tracer", "lfb_set_tracer", "lfb_get_tendency", "lfb_step", "lfb_run_until",
    "lfb_output", "lfb_output_f32", "lfb_output_async", "lfb_wait_transfers", "lfb_host_alloc", "lfb_host_free",
    "lfb_combine", "lfb_output_keep", "lfb_kept_combine", "lfb_kept_fetch", "lfb_kept_fetch_async", "lfb_kept_info", "lfb_kept_release", "lfb_kept_trim", "lfb_kept_reserve",
    "lfb_kept_gather", "lfb_mem_info", "lfb_remap_to_mean", "lfb_remap_release", "lfb_comm_unique_id", "lfb_comm_init", "lfb_kernel_launches",
    "lfb_stage_time_ms", "lfb_timer_start", "lfb_timer_stop", "lfb_kernel_name", "lfb_device_ptr", "lfb_layout",
]


class LfbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("halo", C.c_int32 * 3), ("topology", C.c_int32 * 3),
                ("spacing", C.c_double * 3), ("n_vars", C.c_int32), ("velocity_mask", C.c_int32),
                ("n_coeffs", C.c_int32),
                ("a", C.c_double * MAX_COEFFS), ("b", C.c_double * MAX_COEFFS),
                ("c", C.c_double * MAX_COEFFS), ("d", C.c_double * MAX_COEFFS),
                ("with_maps", C.c_int32), ("advection", C.c_int32), ("n_slots", C.c_int32), ("device", C.c_int32),
                ("strict", C.c_int32), ("kernel", C.c_int32), ("relaxation", C.c_int32),
                ("relax_timescale", C.c_double), ("timestepper", C.c_int32), ("immersed", C.c_int32),
                ("chi", C.c_double), ("reserved", C.c_int32 * 4)]


class RemapDesc(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("halo", C.c_int32 * 3), ("topology", C.c_int32 * 3),
                ("origin", C.c_double * 3), ("spacing", C.c_double * 3), ("has_map", C.c_int32 * 3),
                ("npad", C.c_int32), ("device", C.c_int32), ("radius", C.c_int32), ("reserved", C.c_int32 * 4)]


_lib = None


def load():
    """Load liblfb200.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). lfb200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "lfb_version": (i32, []), "lfb_last_error": (C.c_char_p, []),
        "lfb_create": (i32, [C.POINTER(Desc), C.POINTER(vp)]), "lfb_destroy": (i32, [vp]), "lfb_sync": (i32, [vp]),
        "lfb_n_tracers": (i32, [vp]), "lfb_tracer_index": (i32, [vp, i32, i32, i32]), "lfb_parent_len": (i64, [vp, i32]),
        "lfb_upload_snapshot": (i32, [vp, i32, i32, vp, f64]), "lfb_upload_snapshot_f32": (i32, [vp, i32, i32, vp, f64]),
        "lfb_upload_snapshot_async": (i32, [vp, i32, i32, vp, i32, f64]),
        "lfb_output_async": (i32, [vp, i32, i32, vp, i32]), "lfb_wait_transfers": (i32, [vp]),
        "lfb_host_alloc": (i32, [C.c_size_t, C.POINTER(vp)]), "lfb_host_free": (i32, [vp]),
        "lfb_invalidate_slot": (i32, [vp, i32]), "lfb_set_slot_time": (i32, [vp, i32, f64]),
        "lfb_set_direction": (i32, [vp, i32]), "lfb_set_mask": (i32, [vp, vp]), "lfb_set_immersed": (i32, [vp, vp, vp]),
        "lfb_init_from_data": (i32, [vp]), "lfb_negate_maps": (i32, [vp]), "lfb_reset_clock": (i32, [vp]),
        "lfb_time": (f64, [vp]), "lfb_iteration": (i64, [vp]),
        "lfb_get_tracer": (i32, [vp, i32, vp]), "lfb_set_tracer": (i32, [vp, i32, vp]),
        "lfb_get_tendency": (i32, [vp, i32, vp]),
        "lfb_step": (i32, [vp, f64]), "lfb_run_until": (i32, [vp, f64, f64, i32, C.POINTER(f64)]),
        "lfb_output": (i32, [vp, i32, i32, vp]), "lfb_output_f32": (i32, [vp, i32, i32, vp]),
        "lfb_combine": (i32, [vp, i64, vp, vp, vp, f64, f64, vp]),
        "lfb_output_keep": (i32, [vp, i32, i32, i32, C.POINTER(i32)]),
        "lfb_kept_combine": (i32, [vp, i32, i32, i32, f64, f64, C.POINTER(i32)]),
        "lfb_kept_fetch": (i32, [vp, i32, vp]), "lfb_kept_fetch_async": (i32, [vp, i32, vp]),
        "lfb_kept_info": (i32, [vp, i32, C.POINTER(i32), C.POINTER(i32 * 3), C.POINTER(vp)]),
        "lfb_kept_release": (i32, [vp, i32]), "lfb_kept_trim": (i32, [vp]), "lfb_kept_reserve": (i32, [vp, C.c_size_t, i32]), "lfb_kept_gather": (i32, [vp, i32, i32, C.POINTER(i32)]),
        "lfb_mem_info": (i32, [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "lfb_remap_to_mean": (i32, [C.POINTER(RemapDesc), C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(vp)]),
        "lfb_remap_release": (i32, []),
        "lfb_comm_unique_id": (i32, [vp]), "lfb_comm_init": (i32, [vp, vp, i32, i32]),
        "lfb_kernel_launches": (i64, [vp]), "lfb_stage_time_ms": (i32, [vp, i32, C.POINTER(f64), C.POINTER(i64)]),
        "lfb_timer_start": (i32, [vp]), "lfb_timer_stop": (i32, [vp, C.POINTER(f64)]),
        "lfb_kernel_name": (C.c_char_p, [vp]), "lfb_device_ptr": (vp, [vp, i32, i32]),
        "lfb_layout": (i32, [vp] + [C.POINTER(i64)] * 4),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def check(rc):
    if rc < 0:
        raise LfbError(rc, load().lfb_last_error().decode())
    return rc


def as_parent(arr, shape, dtype=np.float64):
    """Fortran-ordered contiguous view/copy with the parent shape the ABI expects."""
    a = np.asarray(arr)
    if a.shape != tuple(shape):
        raise ValueError(f"parent array has shape {a.shape}, expected {tuple(shape)}")
    return np.asfortranarray(a, dtype=dtype)
