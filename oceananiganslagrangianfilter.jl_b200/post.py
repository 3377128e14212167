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
        self.slab = None          # (rank, world): the arrays are this rank's y-slabs of every output time

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
        """Writes the combined output with the reference's on-disk layout (SURVEY.md Appendix C;
        post_processing_utils.jl:395-404, 751-755, 1417-1424): `timeseries/t/<iteration>` scalars and
        `timeseries/<name>/<iteration>` arrays with halos, in the element type they have here (Float64 for the combined
        and remapped fields, Float32 for the copied original data), as a JLD2 file (`*.jld2`, written by
        jld2.JLD2Writer, readable by JLD2.jl's `jldopen` and by jld2.JLD2File) or, for any other suffix, as an `.npz`
        archive with the same keys.  Returns the path written."""
        payload = {}
        for it, t in zip(self.iterations, self.times):
            payload[f"timeseries/t/{it}"] = float(t)
        for name, series in self.fields.items():
            for it, arr in zip(self.iterations, series):
                payload[f"timeseries/{name}/{it}"] = arr
        for k, v in self.extra.items():
            if np.ndim(v) == 1 and len(v) == len(self.iterations):      # per-output-time scalars, e.g. timeseries/t_shifted
                for it, x in zip(self.iterations, v):
                    payload[f"{k}/{it}"] = float(x)
            else:
                payload[k] = v
        if filename.endswith(".jld2"):
            from .jld2 import JLD2Writer
            with JLD2Writer(filename) as w:
                for k, v in payload.items():
                    w.write(k, v)
            return filename
        stem = filename[:-4] if filename.endswith(".npz") else filename
        np.savez(stem + ".npz", **{k: np.asarray(v) for k, v in payload.items()})
        return stem + ".npz"

    @classmethod
    def load(cls, filename: str) -> "FilterOutput":
        """Reads a combined output file written by `save` (or by the reference) back: every `timeseries/<name>` group
        whose entries are plain arrays."""
        from .jld2 import JLD2File, JLD2FormatError
        f = JLD2File(filename)
        its = f.iterations("t")
        out = cls([f[f"timeseries/t/{it}"] for it in its], its)
        for name in f.keys("timeseries"):
            if name == "t":
                continue
            try:
                first = f[f"timeseries/{name}/{its[0]}"]
            except (KeyError, JLD2FormatError):
                continue
            if np.ndim(first) == 0:
                out.extra[f"timeseries/{name}"] = np.array([f[f"timeseries/{name}/{it}"] for it in its])
            else:
                out.fields[name] = [f[f"timeseries/{name}/{it}"] for it in its]
        return out


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


class ForwardBackwardCombiner:
    """sum_forward_backward_contributions! (post:45-170), incremental: out(t_k) = fwd(t_k) + bwd(T - t_k) for filtered
    variables and maps, fwd - bwd for mean velocities, the backward series interpolated linearly in time
    (post:135-164).  The arithmetic runs on the device in FP64 on the (possibly Float32-rounded) pass outputs.

    The forward outputs are given up front; the backward outputs are fed one output time after the other
    (`feed(tb, outputs)`, in increasing pass time), and every forward time is combined as soon as the backward
    outputs bracketing T - t_k exist, so that -- with pass outputs kept on the device -- the combination and the
    download of the combined fields overlap with the rest of the backward pass.

    Pass outputs are either host arrays (-> lfb_combine) or ids of arrays kept on the device (-> lfb_kept_combine;
    the only form for world > 1).  In the device form the combined arrays are fetched to the host once, into arrays
    from `alloc(shape, dtype)` (default numpy.empty, Fortran order; pass an allocator of page-locked memory for
    full-speed asynchronous copies).  With `regrid`, output time k is first gathered across the y-slabs on rank
    k mod world and remapped to the mean position there (the remap is sharded by output time): every rank ends up
    with the output times it owns as GLOBAL arrays.  Without it nothing is gathered: every rank ends up with its
    y-slab of every output time (`out.slab = (rank, world)`)."""

    def __init__(self, config, fwd: Dict[float, Dict[str, object]], iterations: Dict[float, int], model, btimes=None,
                 rank: int = 0, world: int = 1, regrid: bool = False, alloc=None):
        if model is None:
            raise ValueError("sum_forward_backward_contributions needs the device model (lfb_combine); "
                             "there is no CPU path")
        from .distributed import owner_of_time
        self.config, self.model, self.fwd = config, model, fwd
        self.rank, self.world, self.regrid, self.alloc = rank, world, regrid, alloc
        label, ident = config.label, _identifier(config)
        self.mean_vel_names = ({vel + label + ident for vel in config.velocity_names}
                               if config.compute_mean_velocities else set())
        self.ftimes = sorted(fwd.keys())
        # the backward pass takes the same aligned steps, hence writes at the same pass times
        self.btimes = np.array(sorted(btimes) if btimes is not None else self.ftimes)
        self.names = list(fwd[self.ftimes[0]].keys())
        self.on_device = isinstance(fwd[self.ftimes[0]][self.names[0]], (int, np.integer))
        if world > 1 and not self.on_device:
            raise ValueError("the multi-GPU pipeline combines pass outputs kept on the device")
        self.by_time = regrid or world == 1              # else: by slab
        self.owner = [owner_of_time(k, world) for k in range(len(self.ftimes))]
        self.mine = [k for k in range(len(self.ftimes)) if not self.by_time or self.owner[k] == rank]
        self.out = FilterOutput([self.ftimes[k] for k in self.mine], [iterations[self.ftimes[k]] for k in self.mine])
        if world > 1 and not self.by_time:
            self.out.slab = (rank, world)
        self.plan = [_backward_bracket(self.btimes, config.T - t) for t in self.ftimes]
        self.need = [n1 + (1 if w != 0.0 else 0) for n1, w in self.plan]        # last backward index time k waits for
        self.last_use = {}
        for k, (n1, w) in enumerate(self.plan):
            for nb in ([n1, n1 + 1] if w != 0.0 else [n1]):
                self.last_use[nb] = max(self.last_use.get(nb, -1), k)
        self.bwd: Dict[int, Dict[str, object]] = {}
        self.results: Dict[int, Dict[str, np.ndarray]] = {}
        self.done = set()
        self.deferred: List[int] = []                     # combined arrays whose asynchronous fetch may be pending
        self.pending: Dict[int, Dict[str, int]] = {}      # multi-GPU: gathered output times waiting for their remap

    def feed(self, tb: float, outputs: Dict[str, object]):
        nb = int(np.argmin(np.abs(self.btimes - tb)))
        if abs(self.btimes[nb] - tb) > 1e-9 * max(1.0, abs(tb)):
            raise ValueError(f"unexpected backward output time {tb}")
        self.bwd[nb] = outputs
        if self.on_device and nb not in self.last_use:
            for name in self.names:
                self.model.kept_release(outputs[name])
        have = max(self.bwd) if len(self.bwd) == max(self.bwd) + 1 else -1      # contiguous from 0
        # collectives (gather) must be issued in the same order on every rank: forward times in decreasing order of t,
        # i.e. in the order their backward partners appear
        for k in sorted(range(len(self.ftimes)), key=lambda q: self.need[q]):
            if k not in self.done and self.need[k] <= have:
                self._combine(k)

    def _combine(self, k: int):
        model, t = self.model, self.ftimes[k]
        n1, w = self.plan[k]
        self.done.add(k)
        if not self.on_device:
            res = {}
            for name in self.names:
                sign = -1.0 if name in self.mean_vel_names else 1.0
                res[name] = model.combine(self.fwd[t][name], self.bwd[n1][name], self.bwd[n1 + 1][name] if w != 0.0 else None,
                                          w, sign)
            self.results[k] = res
            return
        for kid in self.deferred:                          # the previous output time's copies are long done
            model.kept_release(kid)
        self.deferred = []
        held = {}
        for name in self.names:
            sign = -1.0 if name in self.mean_vel_names else 1.0
            cid = model.kept_combine(self.fwd[t][name], self.bwd[n1][name], self.bwd[n1 + 1][name] if w != 0.0 else None,
                                     w, sign)
            model.kept_release(self.fwd[t][name])
            if not self.by_time:
                held[name] = cid
                continue
            gid = model.kept_gather(cid, self.owner[k])    # collective; with one rank gid == cid
            if gid != cid:
                model.kept_release(cid)
            if self.owner[k] == self.rank:
                held[name] = gid
        for nb in [n for n, last in self.last_use.items() if last == k]:
            for name in self.names:
                model.kept_release(self.bwd[nb][name])
        if self.by_time and self.owner[k] != self.rank:
            return
        if self.world > 1 and self.regrid:
            # the owner remaps after the pass: a remap here would keep every other rank waiting at its next halo exchange,
            # and the owners of different output times would remap one after the other instead of side by side
            self.pending[k] = held
            return
        self.results[k] = self._remap_and_fetch(held)

    def _remap_and_fetch(self, held: Dict[str, int]) -> Dict[str, np.ndarray]:
        res = {}
        if self.regrid:
            from .remap import remap_kept_to_mean
            res.update(remap_kept_to_mean(self.config, self.model, held, alloc=self.alloc))
        for name in self.names:
            res[name] = self.model.kept_fetch(held[name], alloc=self.alloc, wait=False)
            self.deferred.append(held[name])
        return res

    def finish(self) -> FilterOutput:
        if len(self.done) != len(self.ftimes):
            raise RuntimeError("backward outputs missing: not every forward time could be combined")
        for k in sorted(self.pending):                    # multi-GPU: every owner remaps its output times now, side by side
            self.results[k] = self._remap_and_fetch(self.pending[k])
        self.pending = {}
        if self.on_device:
            self.model.wait_transfers()
            for kid in self.deferred:
                self.model.kept_release(kid)
            self.deferred = []
        for k in self.mine:
            for name, arr in self.results[k].items():
                self.out.fields.setdefault(name, []).append(arr)
        log.info("Combined forward and backward contributions")
        if self.regrid and self.on_device:
            from .remap import remap_release
            remap_release()
            log.info("Wrote regridded data to new variables with _at_mean suffix")
        elif self.regrid:
            regrid_to_mean_position(self.config, self.out)
        return self.out


def sum_forward_backward_contributions(config, fwd: Dict[float, Dict[str, object]], bwd: Dict[float, Dict[str, object]],
                                       iterations: Dict[float, int], model=None, rank: int = 0, world: int = 1,
                                       regrid: bool = False, alloc=None) -> FilterOutput:
    """The combination of two complete passes (see ForwardBackwardCombiner)."""
    c = ForwardBackwardCombiner(config, fwd, iterations, model, btimes=list(bwd.keys()), rank=rank, world=world,
                                regrid=regrid, alloc=alloc)
    for tb in sorted(bwd.keys()):
        c.feed(tb, bwd[tb])
    return c.finish()


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
    if pads_documented_column(g.topology):
        log.warning("Regridding on a grid with a Bounded axis ahead of a Periodic one %s: the reference pads the wrong "
                    "coordinate column there (post_processing_utils.jl:523-611: the column index only advances inside the "
                    "periodic branches); lfb200 pads every Periodic axis's own coordinate, as the reference's comments and "
                    "documentation describe, so the *_at_mean fields differ from the reference's on this topology "
                    "(include/lfb200.h, lfb_remap_to_mean)", tuple(g.topology))
    for k in range(len(out.times)):
        xi = {vel: out.fields["xi_" + vel + label][k] for vel in config.velocity_names
              if "xi_" + vel + label in out.fields}
        res = remap_to_mean_device({n: out.fields[n][k] for n in names}, xi, g.N, g.H, g.topology, g.origin, g.spacing,
                                   npad=config.npad, device=getattr(config, "device", 0),
                                   immersed=_immersed_mask(g))
        for n in names:
            out.fields.setdefault(n + "_at_mean", []).append(res[n + "_at_mean"])
    from .remap import remap_release
    remap_release()
    log.info("Wrote regridded data to new variables with _at_mean suffix")


def pads_documented_column(topology) -> bool:
    """True where the device remap deliberately departs from the reference: among the non-Flat axes a Bounded one comes
    before a Periodic one, so the reference's padding loop (post:523-611) would pad the wrong coordinate column."""
    axes = [t for t in topology if t != "Flat"]
    return any(t == "Periodic" and "Bounded" in axes[:i] for i, t in enumerate(axes))


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
