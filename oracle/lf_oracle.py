"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Wraps oracle/lf_oracle.c and adds, in plain Python, the host-level pieces of the reference's
offline pipeline the checker needs: the Simulation's aligned step sequence (SURVEY.md App. A.2,
reference run_offline_lagrangian_filter.jl:60-80), the forward/backward driver
(run_offline_lagrangian_filter.jl:25-105) and the forward+backward combination
(src/Utils/post_processing_utils.jl:45-170).  It deliberately imports nothing from the product
package so that it stays an independent check.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblf_oracle.so")

TOPO = {"Periodic": 0, "Bounded": 1, "Flat": 2}
MAXP, MAXV = 8, 8


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a few seconds)."""
    src = os.path.join(_HERE, "lf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "MAKEFLAGS": ""})
    return _LIB_PATH


class _Problem(C.Structure):
    _fields_ = [("N", C.c_int * 3), ("H", C.c_int * 3), ("topo", C.c_int * 3), ("delta", C.c_double * 3),
                ("n_vars", C.c_int), ("vel_present", C.c_int * 3), ("n_pairs", C.c_int), ("single_exp", C.c_int),
                ("with_maps", C.c_int), ("advect", C.c_int),
                ("a", C.c_double * MAXP), ("b", C.c_double * MAXP), ("c", C.c_double * MAXP), ("d", C.c_double * MAXP),
                ("relax", C.c_int), ("relax_timescale", C.c_double),
                ("timestepper", C.c_int), ("chi", C.c_double), ("immersed", C.c_int), ("weights_f32", C.c_int)]


ADVECTION = {None: 0, "WENO5": 1, "Centered2": 2, "Centered4": 3, "UpwindBiased1": 4, "UpwindBiased3": 5,
             "UpwindBiased5": 6, "WENO3": 7}


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.lfo_create.restype = C.c_void_p
        L.lfo_create.argtypes = [C.POINTER(_Problem)]
        L.lfo_destroy.argtypes = [C.c_void_p]
        L.lfo_n_tracers.argtypes = [C.c_void_p]
        L.lfo_center_len.argtypes = [C.c_void_p]
        L.lfo_center_len.restype = C.c_long
        L.lfo_vel_len.argtypes = [C.c_void_p, C.c_int]
        L.lfo_vel_len.restype = C.c_long
        L.lfo_time.argtypes = [C.c_void_p]
        L.lfo_time.restype = C.c_double
        L.lfo_reset_clock.argtypes = [C.c_void_p]
        PP = C.POINTER(C.c_void_p)
        L.lfo_load_series.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int,
                                      PP, PP, PP, C.POINTER(PP)]
        L.lfo_set_mask.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_set_immersed.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_set_cell_heights.argtypes = [C.c_void_p, C.c_void_p]
        L.lfo_update_state.argtypes = [C.c_void_p]
        L.lfo_init_from_data.argtypes = [C.c_void_p]
        L.lfo_negate_maps.argtypes = [C.c_void_p]
        L.lfo_time_step.argtypes = [C.c_void_p, C.c_double]
        L.lfo_get_tracer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfo_set_tracer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfo_get_tendency.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lfo_output.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.lfo_bw2_params.argtypes = [C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 4
        L.lfo_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def set_threads(n: int = 0) -> int:
    """Set the OpenMP thread count of the oracle (n <= 0: leave it); returns the count in effect."""
    return int(lib().lfo_set_threads(int(n)))


def bw2_params(N: int, freq_c: float) -> Dict[str, float]:
    """set_offline_BW2_filter_params (reference utils:251-283) as a dict a1,b1,c1,d1,...,N_coeffs."""
    arr = [(C.c_double * MAXP)() for _ in range(4)]
    nc = lib().lfo_bw2_params(N, freq_c, *arr)
    if nc < 0:
        raise ValueError("N must be a non-negative even integer, or 1 for a single exponential.")
    if nc == 0:
        return {"a1": arr[0][0], "c1": arr[2][0], "N_coeffs": 0.5}
    out: Dict[str, float] = {}
    for i in range(nc):
        for name, a in zip("abcd", arr):
            out[f"{name}{i + 1}"] = a[i]
    out["N_coeffs"] = nc
    return out


def step_sequence(T: float, dt: float, T_out: float, progress_interval: Optional[float] = None) -> List[float]:
    """Oceananigans Simulation time-step alignment: every step is clipped so that
    the clock lands on the next multiple of the output interval, of the progress-callback interval
    (T/10 in run_offline_lagrangian_filter.jl:74) and on the stop time."""
    t, seq = 0.0, []
    while t < T:
        cands = [dt, (math.floor(t / T_out) + 1) * T_out - t, T - t]
        if progress_interval:
            cands.append((math.floor(t / progress_interval) + 1) * progress_interval - t)
        step = min(cands)
        seq.append(step)
        t = t + step
    return seq


class Oracle:
    """One LagrangianFilter model + its input series, CPU reference numerics."""

    def __init__(self, N, H, topology, delta, n_vars, velocities="uvw", filter_params=None, with_maps=True,
                 advect=True, relax_timescale=None, mask=None, timestepper="RungeKutta3", chi=0.1, immersed=None,
                 weights_f32=False, cell_heights=None):
        """cell_heights: parent array of dz(i,j,k) (PartialCellBottom) or None.  advect: True (the default WENO(order=5)), False / None (advection = nothing) or a scheme name from
        ADVECTION; timestepper: "RungeKutta3" | "QuasiAdamsBashforth2"; immersed: parent array of 0 / 1 flags."""
        P = _Problem()
        nc = filter_params["N_coeffs"]
        single = nc == 0.5
        npairs = 1 if single else int(nc)
        for ax in range(3):
            P.N[ax], P.H[ax] = int(N[ax]), int(H[ax])
            P.topo[ax] = TOPO[topology[ax]]
            P.delta[ax] = float(delta[ax])
            P.vel_present[ax] = int("uvw"[ax] in velocities)
        P.n_vars, P.n_pairs, P.single_exp = n_vars, npairs, int(single)
        P.with_maps = int(with_maps)
        P.advect = ADVECTION[advect] if (advect is None or isinstance(advect, str)) else int(bool(advect))
        P.timestepper = {"RungeKutta3": 0, "QuasiAdamsBashforth2": 1}[timestepper]
        P.chi = float(chi)
        P.weights_f32 = int(bool(weights_f32))
        for i in range(npairs):
            P.a[i] = filter_params[f"a{i + 1}"]
            P.c[i] = filter_params[f"c{i + 1}"]
            P.b[i] = 0.0 if single else filter_params[f"b{i + 1}"]
            P.d[i] = 0.0 if single else filter_params[f"d{i + 1}"]
        P.relax = int(relax_timescale is not None)
        P.relax_timescale = float(relax_timescale or 0.0)
        self.P = P
        self.L = lib()
        self.h = C.c_void_p(self.L.lfo_create(C.byref(P)))
        self.N, self.H, self.topology = tuple(N), tuple(H), tuple(topology)
        self.n_vars, self.velocities = n_vars, velocities
        self.center_shape = tuple(N[ax] + 2 * H[ax] for ax in range(3))
        self.n_tracers = self.L.lfo_n_tracers(self.h)
        self._keep = []
        if mask is not None:
            m = np.asfortranarray(mask, dtype=np.float64)
            assert m.shape == self.center_shape
            self.L.lfo_set_mask(self.h, m.ctypes.data)
        if immersed is not None:
            m = np.asfortranarray(immersed, dtype=np.float64)
            assert m.shape == self.center_shape
            self.L.lfo_set_immersed(self.h, m.ctypes.data)
        if cell_heights is not None:
            m = np.asfortranarray(cell_heights, dtype=np.float64)
            assert m.shape == self.center_shape
            self.L.lfo_set_cell_heights(self.h, m.ctypes.data)

    def close(self):
        if self.h:
            self.L.lfo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def vel_shape(self, ax):
        s = list(self.center_shape)
        if self.topology[ax] == "Bounded":
            s[ax] += 1
        return tuple(s)

    def load_series(self, times, vel: Dict[str, Sequence[np.ndarray]], vars_: Sequence[Sequence[np.ndarray]],
                    T_start: float, T_end: float, backward: bool):
        """create_input_data_on_disk + load_data (reference utils:97-215) for one direction."""
        n = len(times)
        tarr = (C.c_double * n)(*[float(t) for t in times])
        keep = []

        def ptrs(arrs, shape):
            out = (C.c_void_p * n)()
            for i, a in enumerate(arrs):
                a = np.asfortranarray(a, dtype=np.float64)
                assert a.shape == shape, (a.shape, shape)
                keep.append(a)
                out[i] = a.ctypes.data
            return out

        vp = []
        for ax, name in enumerate("uvw"):
            vp.append(ptrs(vel[name], self.vel_shape(ax)) if name in self.velocities else (C.c_void_p * n)())
        PP = C.POINTER(C.c_void_p)
        var_ptrs = (PP * max(1, self.n_vars))()
        var_keep = []
        for q in range(self.n_vars):
            p = ptrs(vars_[q], self.center_shape)
            var_keep.append(p)
            var_ptrs[q] = C.cast(p, PP)
        self.L.lfo_load_series(self.h, n, tarr, float(T_start), float(T_end), int(backward),
                               C.cast(vp[0], PP), C.cast(vp[1], PP), C.cast(vp[2], PP), var_ptrs)

    def init_from_data(self):
        self.L.lfo_init_from_data(self.h)

    def update_state(self):
        self.L.lfo_update_state(self.h)

    def step(self, dt):
        self.L.lfo_time_step(self.h, float(dt))

    def negate_maps(self):
        self.L.lfo_negate_maps(self.h)

    def reset_clock(self):
        self.L.lfo_reset_clock(self.h)

    @property
    def time(self):
        return self.L.lfo_time(self.h)

    def tracer(self, t):
        out = np.empty(self.center_shape, dtype=np.float64, order="F")
        self.L.lfo_get_tracer(self.h, t, out.ctypes.data)
        return out

    def tendency(self, t):
        out = np.empty(self.center_shape, dtype=np.float64, order="F")
        self.L.lfo_get_tendency(self.h, t, out.ctypes.data)
        return out

    def set_tracer(self, t, arr):
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.center_shape
        self.L.lfo_set_tracer(self.h, t, a.ctypes.data)

    def output(self, kind, field):
        out = np.empty(self.center_shape, dtype=np.float64, order="F")
        self.L.lfo_output(self.h, kind, field, out.ctypes.data)
        return out


def run_offline(times, vel, vars_, var_names, N, H, topology, delta, filter_params, dt, T_out,
                T_start=None, T_end=None, velocities="uw", map_to_mean=True, compute_mean_velocities=True,
                advect=True, float32_outputs=True, relax_timescale=None, mask=None, return_state=False,
                immersed=None, weights_f32=False, cell_heights=None, timestepper="RungeKutta3", chi=0.1):
    """Oracle restatement of run_offline_Lagrangian_filter up to and including
    sum_forward_backward_contributions! (reference run_offline_lagrangian_filter.jl:25-105).

    Returns {name: [array per output time]} plus "t".  With float32_outputs=True each pass's outputs
    are rounded to Float32 before the combination, as the reference's JLD2Writer does
    (SURVEY.md App. A.8)."""
    times = np.asarray(times, dtype=np.float64)
    T_start = times[0] if T_start is None else T_start
    T_end = times[-1] if T_end is None else T_end
    T = T_end - T_start
    sel = [i for i, t in enumerate(times) if T_start <= t <= T_end]
    tsel = times[sel]
    vel_s = {k: [v[i] for i in sel] for k, v in vel.items()}
    vars_s = [[v[i] for i in sel] for v in vars_]
    with_maps = map_to_mean or compute_mean_velocities
    o = Oracle(N, H, topology, delta, len(var_names), velocities, filter_params, with_maps, advect,
               relax_timescale, mask, immersed=immersed, weights_f32=weights_f32, cell_heights=cell_heights,
               timestepper=timestepper, chi=chi)
    vel_names = [n for n in "uvw" if n in velocities]

    def collect():
        out = {}
        for q, name in enumerate(var_names):
            out[name + ("_Lagrangian_filtered" if advect else "_Eulerian_filtered")] = o.output(0, q)
        if map_to_mean:
            for m, vn in enumerate(vel_names):
                out["xi_" + vn] = o.output(1, m)
        if compute_mean_velocities:
            for m, vn in enumerate(vel_names):
                out[vn + ("_Lagrangian_filtered" if advect else "_Eulerian_filtered")] = o.output(2, m)
        return out

    def run_pass():
        o.update_state()
        outs = {0.0: collect()}
        for step in step_sequence(T, dt, T_out, T / 10):
            o.step(step)
            t = o.time
            if abs(t / T_out - round(t / T_out)) < 1e-12:
                outs[round(t / T_out) * T_out] = collect()
        return outs

    o.load_series(tsel, vel_s, vars_s, T_start, T_end, backward=False)
    o.init_from_data()
    fwd = run_pass()
    o.load_series(tsel, vel_s, vars_s, T_start, T_end, backward=True)
    o.negate_maps()
    o.reset_clock()
    bwd = run_pass()

    rnd = (lambda x: x.astype(np.float32).astype(np.float64)) if float32_outputs else (lambda x: x)
    bt = np.array(sorted(bwd.keys()))

    def bwd_at(name, t):
        # FieldTimeSeries linear interpolation of the backward output (post:138-146)
        if t <= bt[0]:
            return rnd(bwd[bt[0]][name])
        if t >= bt[-1]:
            return rnd(bwd[bt[-1]][name])
        n1 = int(np.searchsorted(bt, t, side="right") - 1)
        if bt[n1] == t:
            return rnd(bwd[bt[n1]][name])
        w = (t - bt[n1]) / (bt[n1 + 1] - bt[n1])
        return rnd(bwd[bt[n1 + 1]][name]) * w + rnd(bwd[bt[n1]][name]) * (1 - w)

    result = {"t": []}
    mean_vel_names = {vn + ("_Lagrangian_filtered" if advect else "_Eulerian_filtered") for vn in vel_names}
    for t in sorted(fwd.keys()):
        result["t"].append(t)
        for name, arr in fwd[t].items():
            b = bwd_at(name, T - t)
            comb = rnd(arr) - b if name in mean_vel_names else rnd(arr) + b
            result.setdefault(name, []).append(comb)
    if return_state:
        result["_final_state"] = [o.tracer(i) for i in range(o.n_tracers)]
    o.close()
    return result
