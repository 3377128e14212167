"""Post-processing of the two filter passes: host-side mirror of reference
src/Utils/post_processing_utils.jl.

  sum_forward_backward_contributions   post:45-170    device kernel through lfb_combine
  regrid_to_mean_position              post:327-762   remap of filtered fields to the mean position
  compute_Eulerian_filter              post:1324-1380 host utility (weighted sums of the output series)
  get_weight_function                  post:1170-1197
  get_offline/online_frequency_response post:1221-1294
  compute_time_shift                   post:1396-1428

The reference works file-to-file on JLD2; here the same transforms act on an in-memory
`FilterOutput` (name -> one parent array per output time) which can be saved afterwards.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence

import numpy as np

log = logging.getLogger("lfb200")


class FilterOutput:
    """Combined filter output: `times`, `iterations` (the reference's JLD2 keys) and per variable a
    list with one parent array (halos included, Julia (x,y,z) order) per output time."""

    def __init__(self, times, iterations):
        self.times = [float(t) for t in times]
        self.iterations = [int(i) for i in iterations]
        self.fields: Dict[str, List[np.ndarray]] = {}
        self.stats: Dict[str, object] = {}
        self.extra: Dict[str, np.ndarray] = {}

    def __getitem__(self, name):
        return self.fields[name]

    def __contains__(self, name):
        return name in self.fields

    def keys(self):
        return self.fields.keys()

    def stacked(self, name) -> np.ndarray:
        return np.stack(self.fields[name], axis=0)

    def add_original_data(self, config, source, indices, file_times):
        """Copies the unfiltered fields at the output times (the reference's writer outputs
        model.auxiliary_fields / velocities, utils:1001-1008; post:118-131)."""
        ft = np.asarray(file_times)[indices] - config.T_start
        for name in tuple(config.var_names_to_filter) + tuple(config.velocity_names):
            series = []
            for t in self.times:
                n1 = int(np.searchsorted(ft, t, side="right") - 1)
                n1 = min(max(n1, 0), len(ft) - 1)
                a1 = np.asarray(source.snapshot(name, indices[n1]))
                if ft[n1] == t or n1 == len(ft) - 1:
                    series.append(np.array(a1, dtype=np.float32))
                else:
                    a2 = np.asarray(source.snapshot(name, indices[n1 + 1]))
                    w = (t - ft[n1]) / (ft[n1 + 1] - ft[n1])
                    series.append((a2.astype(np.float64) * w + a1.astype(np.float64) * (1 - w)).astype(np.float32))
            self.fields[name] = series

    def save(self, filename: str):
        """Writes `<stem>.npz` with `timeseries/<name>/<iteration>`-style keys (the JLD2 group layout of
        the reference flattened into npz keys)."""
        stem = filename[:-5] if filename.endswith(".jld2") else filename
        payload = {"timeseries/t": np.array(self.times), "iterations": np.array(self.iterations)}
        for name, series in self.fields.items():
            for it, arr in zip(self.iterations, series):
                payload[f"timeseries/{name}/{it}"] = arr
        for k, v in self.extra.items():
            payload[k] = v
        np.savez(stem + ".npz", **payload)
        return stem + ".npz"


def _identifier(config) -> str:
    return "_Eulerian_filtered" if getattr(config, "advection", True) is None else "_Lagrangian_filtered"


def _backward_bracket(btimes: np.ndarray, tb: float):
    """Index n1 and weight w of the backward series' linear interpolation at pass time tb (post:138-146)."""
    if tb <= btimes[0]:
        return 0, 0.0
    if tb >= btimes[-1]:
        return len(btimes) - 1, 0.0
    n1 = int(np.searchsorted(btimes, tb, side="right") - 1)
    w = 0.0 if btimes[n1] == tb else (tb - btimes[n1]) / (btimes[n1 + 1] - btimes[n1])
    return n1, w


def sum_forward_backward_contributions(config, fwd: Dict[float, Dict[str, object]], bwd: Dict[float, Dict[str, object]],
                                       iterations: Dict[float, int], model=None, rank: int = 0, world: int = 1,
                                       regrid: bool = False) -> FilterOutput:
    """out(t_k) = fwd(t_k) + bwd(T - t_k) for filtered variables and maps, fwd - bwd for mean velocities;
    the backward series is interpolated linearly in time (reference post:135-164).  The arithmetic runs
    on the device in FP64 on the (possibly Float32-rounded) pass outputs.

    The pass outputs are either host arrays (-> lfb_combine) or ids of arrays kept on the device
    (-> lfb_kept_combine).  In the second form, which is the only one for world > 1, output time k is then gathered
    across the y-slabs on rank k mod world, remapped to the mean position there when `regrid` (the remap is
    sharded by output time) and fetched to the host once; every rank returns the output times it owns."""
    if model is None:
        raise ValueError("sum_forward_backward_contributions needs the device model (lfb_combine); "
                         "there is no CPU path")
    from .distributed import owner_of_time
    T = config.T
    label = config.label
    ident = _identifier(config)
    mean_vel_names = {vel + label + ident for vel in config.velocity_names} if config.compute_mean_velocities else set()
    ftimes = sorted(fwd.keys())
    btimes = np.array(sorted(bwd.keys()))
    on_device = isinstance(next(iter(fwd[ftimes[0]].values())), (int, np.integer))
    if world > 1 and not on_device:
        raise ValueError("the multi-GPU pipeline combines pass outputs kept on the device")
    mine = [k for k in range(len(ftimes)) if owner_of_time(k, world) == rank]
    out = FilterOutput([ftimes[k] for k in mine], [iterations[ftimes[k]] for k in mine])
    if not on_device:
        for t in ftimes:
            n1, w = _backward_bracket(btimes, T - t)
            for name, arr in fwd[t].items():
                lo = bwd[btimes[n1]][name]
                hi = bwd[btimes[n1 + 1]][name] if w != 0.0 else None
                sign = -1.0 if name in mean_vel_names else 1.0
                out.fields.setdefault(name, []).append(model.combine(arr, lo, hi, w, sign))
        log.info("Combined forward and backward contributions")
        if regrid:
            regrid_to_mean_position(config, out)
        return out

    # ---- device form ----
    # which backward outputs are still needed by later forward times (they are released as soon as possible)
    last_use = {}
    plan = []
    for k, t in enumerate(ftimes):
        n1, w = _backward_bracket(btimes, T - t)
        plan.append((n1, w))
        last_use[n1] = k
        if w != 0.0:
            last_use[n1 + 1] = k
    names = list(fwd[ftimes[0]].keys())
    for nb in range(len(btimes)):                      # backward outputs no forward time refers to
        if nb not in last_use:
            for name in names:
                model.kept_release(bwd[btimes[nb]][name])
    for k, t in enumerate(ftimes):
        n1, w = plan[k]
        owner = owner_of_time(k, world)
        held = {}
        for name in names:
            sign = -1.0 if name in mean_vel_names else 1.0
            cid = model.kept_combine(fwd[t][name], bwd[btimes[n1]][name], bwd[btimes[n1 + 1]][name] if w != 0.0 else None,
                                     w, sign)
            model.kept_release(fwd[t][name])
            gid = model.kept_gather(cid, owner)            # collective; with one rank gid == cid
            if gid != cid:
                model.kept_release(cid)
            if owner == rank:
                held[name] = gid
        for nb in [n for n, last in last_use.items() if last == k]:
            for name in names:
                model.kept_release(bwd[btimes[nb]][name])
        if owner != rank:
            continue
        if regrid:
            from .remap import remap_kept_to_mean
            for name, arr in remap_kept_to_mean(config, model, held).items():
                out.fields.setdefault(name, []).append(arr)
        for name in names:
            out.fields.setdefault(name, []).append(model.kept_fetch(held[name]))
            model.kept_release(held[name])
    log.info("Combined forward and backward contributions")
    if regrid:
        log.info("Wrote regridded data to new variables with _at_mean suffix")
    return out


# ---------------------------------------------------------------------------------------------- remap
def _regrid_names(config, extra_vars_to_regrid: Sequence[str] = ()):
    """Variables regrid_to_mean_position! remaps (post:340-348)."""
    label = config.label
    names = list(config.var_names_to_filter)
    if config.compute_mean_velocities:
        names += list(config.velocity_names)
    names = [n + label + "_Lagrangian_filtered" for n in names]
    for n in extra_vars_to_regrid:
        if n not in names:
            names.append(n)
    return names


def regrid_to_mean_position(config, out: FilterOutput, extra_vars_to_regrid: Sequence[str] = ()):
    """Interpolates every `<var>_Lagrangian_filtered` field from the displaced positions x + xi to the
    grid nodes (generalised Lagrangian mean position), adding `<name>_at_mean` (reference post:327-762).
    Runs on the device (lfb_remap_to_mean); there is no host path."""
    from .remap import remap_to_mean_device
    if getattr(config, "advection", True) is None:
        raise ValueError("Regridding to mean position is not meaningful for Eulerian filtering")
    label = config.label
    names = _regrid_names(config, extra_vars_to_regrid)
    g = config.grid
    if len(g.nonflat_axes) < 2:
        raise ValueError("regridding to the mean position needs at least two non-Flat axes (the reference's "
                         "LinearNDInterpolator does not exist in one dimension either)")
    for k in range(len(out.times)):
        xi = {vel: out.fields["xi_" + vel + label][k] for vel in config.velocity_names
              if "xi_" + vel + label in out.fields}
        res = remap_to_mean_device({n: out.fields[n][k] for n in names}, xi, g.N, g.H, g.topology, g.origin, g.spacing,
                                   npad=config.npad, device=getattr(config, "device", 0),
                                   immersed=_immersed_mask(g))
        for n in names:
            out.fields.setdefault(n + "_at_mean", []).append(res[n + "_at_mean"])
    log.info("Wrote regridded data to new variables with _at_mean suffix")


def _immersed_mask(grid):
    """Cells _mask_immersed (post:285-298) blanks before the remap: z_centre < bottom_height, halos included."""
    from .grids import ImmersedBoundaryGrid
    if not isinstance(grid, ImmersedBoundaryGrid):
        return None
    return grid.below_bottom_mask()


# ---------------------------------------------------------------------------------------------- Eulerian
def get_weight_function(t, tref, filter_params, direction: str = "both") -> np.ndarray:
    """Impulse response of the filter at times `t` relative to `tref` (reference post:1170-1197)."""
    t = np.asarray(t, dtype=np.float64)
    at = np.abs(t - tref)
    if filter_params["N_coeffs"] == 0.5:
        G = filter_params["a1"] * np.exp(-filter_params["c1"] * at)
    else:
        G = np.zeros_like(t)
        for i in range(1, int(filter_params["N_coeffs"]) + 1):
            a, b, c, d = (filter_params[f"{k}{i}"] for k in "abcd")
            G = G + (a * np.cos(d * at) + b * np.sin(d * at)) * np.exp(-c * at)
    if direction == "forward":
        G[t > tref] = 0
    elif direction == "backward":
        G[t < tref] = 0
    elif direction != "both":
        raise ValueError("Direction must be 'forward', 'backward' or 'both'")
    return G


def get_offline_frequency_response(freq, filter_params) -> np.ndarray:
    """reference post:1221-1245"""
    freq = np.asarray(freq, dtype=np.float64)
    if filter_params["N_coeffs"] == 0.5:
        a1, c1 = filter_params["a1"], filter_params["c1"]
        return (2.0 * a1 * c1) / (c1 ** 2 + freq ** 2)
    G = np.zeros_like(freq)
    for i in range(1, int(filter_params["N_coeffs"]) + 1):
        a, b, c, d = (filter_params[f"{k}{i}"] for k in "abcd")
        G = G + (a * c + b * (d + freq)) / (c ** 2 + (d + freq) ** 2) + (a * c + b * (d - freq)) / (c ** 2 + (d - freq) ** 2)
    return G


def get_online_frequency_response(freq, filter_params) -> np.ndarray:
    """reference post:1270-1294"""
    freq = np.asarray(freq, dtype=np.float64)
    if filter_params["N_coeffs"] == 0.5:
        return filter_params["a1"] / (filter_params["c1"] + 1j * freq)
    G = np.zeros_like(freq, dtype=np.complex128)
    for i in range(1, int(filter_params["N_coeffs"]) + 1):
        a, b, c, d = (filter_params[f"{k}{i}"] for k in "abcd")
        G = G + 0.5 * ((a + 1j * b) / (c + 1j * (d + freq)) + (a - 1j * b) / (c + 1j * (-d + freq)))
    return G


def compute_Eulerian_filter(config, out: FilterOutput):
    """Direct convolution of the OUTPUT series with the filter's weight function, for comparison with
    the Lagrangian filter (reference post:1324-1380)."""
    names = list(config.var_names_to_filter)
    if config.compute_mean_velocities:
        names += list(config.velocity_names)
    times = np.asarray(out.times)
    for name in names:
        if name not in out.fields:
            raise KeyError(f"{name!r} is not in the output (output_original_data=false?)")
        series = [np.asarray(a, dtype=np.float64) for a in out.fields[name]]
        res = []
        for t in times:
            G = get_weight_function(times, t, config.filter_params, "both")
            norm = np.sum(G[:-1] * np.diff(times))
            mean = np.zeros_like(series[0])
            dt = times[1] - times[0]
            for j in range(len(times)):
                dt = times[j + 1] - times[j] if j < len(times) - 1 else dt
                mean = mean + G[j] * series[j] * dt
            res.append(mean / norm)
        out.fields[name + config.label + "_Eulerian_filtered"] = res


def compute_time_shift(config, out: FilterOutput):
    """reference post:1396-1428: shifted time axis t - time_shift of a one-sided filter."""
    fp = config.filter_params
    if fp["N_coeffs"] == 0.5:
        shift = 1 / fp["c1"]
    else:
        shift = 0.0
        for i in range(1, int(fp["N_coeffs"]) + 1):
            a, b, c, d = (fp[f"{k}{i}"] for k in "abcd")
            shift += (a * c ** 2 + 2 * b * c * d - a * d ** 2) / (c ** 2 + d ** 2) ** 2
    out.extra["timeseries/t_shifted"] = np.asarray(out.times) - shift
    return shift
