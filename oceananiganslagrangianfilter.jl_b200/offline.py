"""`run_offline_Lagrangian_filter`: host orchestration of the offline filter.

Mirror of reference src/OfflineLagrangianFilter/run_offline_lagrangian_filter.jl:25-132 with the
model / Simulation block (:47-102) re-pointed at liblfb200.so:

    reference                                           here
    create_input_data_on_disk (forward, :28)         -> SnapshotFeeder.begin(+1): slots tagged t - T_start
    load_data / FieldTimeSeries InMemory(4) (:31)    -> SnapshotFeeder: n_slots resident snapshots, streamed
    LagrangianFilter(...) (:47)                      -> model.LagrangianFilter -> lfb_create
    initialise_filtered_vars_from_data (:52)         -> lfb_init_from_data
    Simulation + callbacks + JLD2Writer + run! (:60-83) -> aligned_time_steps + lfb_step + lfb_output
    create_input_data_on_disk (backward, :87)        -> SnapshotFeeder.begin(-1): re-tag T_end - t, negate on the fly
    change_sign_of_map_variables!, reset! (:90-93)   -> lfb_negate_maps, lfb_reset_clock
    sum_forward_backward_contributions! (:105)       -> post.sum_forward_backward_contributions (lfb_combine)
    regrid_to_mean_position! (:118-120)              -> post.regrid_to_mean_position
    compute_Eulerian_filter! (:123-125)              -> post.compute_Eulerian_filter

No file is rewritten between the passes: the backward pass re-uses the snapshots already resident on
the device (or streams them in reverse order when they do not all fit in the slots).
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .config import OfflineFilterConfig
from .model import LagrangianFilter
from .post import (FilterOutput, compute_Eulerian_filter, regrid_to_mean_position,
                   sum_forward_backward_contributions)

log = logging.getLogger("lfb200")


def aligned_time_steps(T: float, Δt: float, intervals) -> List[float]:
    """The step sizes Oceananigans' `run!` takes for `Simulation(model; Δt, stop_time=T)` with
    TimeInterval schedules `intervals` attached (output writer T_out and the progress callback T/10,
    reference run_offline_lagrangian_filter.jl:60,74,77-80): every step is clipped to land on the next
    multiple of each interval and on the stop time."""
    t, steps = 0.0, []
    while t < T:
        step = min([Δt, T - t] + [(math.floor(t / iv) + 1) * iv - t for iv in intervals if iv])
        if not step > 0:
            raise RuntimeError(f"time step alignment produced Δt={step} at t={t}")
        steps.append(step)
        t = t + step
    return steps


def _is_output_time(t: float, T_out: float) -> bool:
    return abs(t / T_out - round(t / T_out)) < 1e-12


class SnapshotFeeder:
    """Keeps the input snapshots a pass needs resident in the model's slots (the reference's
    `FieldTimeSeries(...; backend = InMemory(4))` sliding window, utils:211-212)."""

    def __init__(self, model: LagrangianFilter, source, indices, file_times, T_start, T_end):
        self.model, self.source = model, source
        self.indices = list(indices)                       # positions in the file, forward order
        self.file_times = np.asarray(file_times, dtype=np.float64)
        self.T_start, self.T_end = T_start, T_end
        self.names = list(model.velocities_present) + list(model.var_names)
        self.all_resident = len(self.indices) <= model.n_slots
        self.uploaded = False
        self.slot_of: Dict[int, int] = {}                  # pass position -> slot
        self.bytes_uploaded = 0

    def begin(self, direction: int):
        self.direction = direction
        n = len(self.indices)
        if direction > 0:
            self.order = list(range(n))
            self.pass_times = [self.file_times[self.indices[p]] - self.T_start for p in self.order]
        else:
            self.order = list(range(n - 1, -1, -1))
            self.pass_times = [self.T_end - self.file_times[self.indices[p]] for p in self.order]
        self.model.set_direction(direction)
        if self.all_resident:
            if not self.uploaded:
                for pos, p in enumerate(self.order):
                    self._upload(pos, p, slot=p)
                self.uploaded = True
            self.slot_of = {}
            for pos, p in enumerate(self.order):
                self.model.set_slot_time(p, self.pass_times[pos])       # re-tag only, data stays put
                self.slot_of[pos] = p
        else:
            for s in range(self.model.n_slots):
                self.model.invalidate_slot(s)
            self.slot_of = {}

    def _upload(self, pos: int, p: int, slot: int):
        n = self.indices[p]
        if not self.all_resident:
            self.model.wait_transfers()          # bound the number of host arrays kept alive for queued copies
        for name in self.names:
            arr = self.source.snapshot(name, n)
            self.model.upload_snapshot(slot, name, arr, self.pass_times[pos], wait=False)
            self.bytes_uploaded += np.asarray(arr).nbytes
        self.slot_of[pos] = slot

    def ensure(self, t_lo: float, t_hi: float):
        """Make the snapshots bracketing every time in [t_lo, t_hi] resident."""
        if self.all_resident:
            return
        times = self.pass_times
        lo = max(i for i, t in enumerate(times) if t <= t_lo) if times[0] <= t_lo else 0
        hi = min((i for i, t in enumerate(times) if t >= t_hi), default=len(times) - 1)
        if hi - lo + 1 > self.model.n_slots:
            raise RuntimeError(f"a step from t={t_lo} to t={t_hi} spans {hi - lo + 1} snapshots but only "
                               f"{self.model.n_slots} slots are resident; raise n_slots or lower Δt")
        want = list(range(lo, hi + 1))
        if hi + 1 < len(times) and len(want) < self.model.n_slots:
            want.append(hi + 1)                                           # prefetch the next one
        for pos in [q for q in self.slot_of if q < lo]:
            del self.slot_of[pos]
        used = set(self.slot_of.values())
        for pos in want:
            if pos in self.slot_of:
                continue
            free = [s for s in range(self.model.n_slots) if s not in used]
            slot = free[0]
            self._upload(pos, self.order[pos], slot)
            used.add(slot)


def _collect_outputs(model: LagrangianFilter, config: OfflineFilterConfig, dtype) -> Dict[str, np.ndarray]:
    """create_output_fields (utils:898-1012) evaluated now, with the reference's output names."""
    ident = "_Eulerian_filtered" if config.advection is None else "_Lagrangian_filtered"
    label = config.label
    out = {}

    def queue(kind, index):
        # asynchronous: the copy to the (page-locked) result array overlaps with the following time steps
        return model.output_async(kind, index, model.pinned_empty(model.center_shape, dtype))

    for q, v in enumerate(config.var_names_to_filter):
        out[v + label + ident] = queue(_lib.OUT_FILTERED_VAR, q)
    vels = model.velocities_present
    if config.map_to_mean:
        for m, vel in enumerate(vels):
            out["xi_" + vel + label] = queue(_lib.OUT_XI, m)
    if config.compute_mean_velocities:
        for m, vel in enumerate(vels):
            out[vel + label + ident] = queue(_lib.OUT_MEAN_VEL, m)
    return out


def _run_pass(model, feeder, config, dtype, progress=True):
    T, Δt, T_out = config.T, config.Δt, config.T_out
    feeder.ensure(0.0, 0.0)
    outputs = {0.0: _collect_outputs(model, config, dtype)}
    iterations = {0.0: 0}
    for n, step in enumerate(aligned_time_steps(T, Δt, [T_out, T / 10]), 1):
        t = model.time
        feeder.ensure(t, t + step)
        model.step(step)
        t = model.time
        if _is_output_time(t, T_out):
            tk = round(t / T_out) * T_out
            outputs[tk] = _collect_outputs(model, config, dtype)
            iterations[tk] = n
    model.wait_transfers()
    # results leave the page-locked staging arrays, which are released
    outputs = {t: {k: np.array(v, order="F") for k, v in d.items()} for t, d in outputs.items()}
    model.free_pinned()
    return outputs, iterations


def build_model(config: OfflineFilterConfig) -> LagrangianFilter:
    mask = None
    if config.boundary_relaxation:
        mask = evaluate_mask(config)
    return LagrangianFilter(config.grid, var_names_to_filter=config.var_names_to_filter,
                            velocity_names=config.velocity_names, filter_params=config.filter_params,
                            with_maps=config.map_to_mean or config.compute_mean_velocities,
                            advection=config.advection, n_slots=config.n_slots, device=config.device,
                            timestepper=config.timestepper,
                            strict=config.strict, kernel=config.kernel, label=config.label,
                            relax_timescale=config.relax_timescale if config.boundary_relaxation else None,
                            mask=mask, config=config)


def evaluate_mask(config) -> np.ndarray:
    """Pre-evaluates `mask_func(x..., mask_params)` at every cell centre, halos included (the reference
    evaluates it inside the forcing closure at the tracer location, utils:592-697)."""
    g = config.grid
    coords = [g.nodes(ax) for ax in g.nonflat_axes]
    mesh = np.meshgrid(*coords, indexing="ij")
    vals = np.vectorize(lambda *xs: float(config.mask_func(*xs, config.mask_params)))(*mesh)
    shape = [1, 1, 1]
    for n, ax in enumerate(g.nonflat_axes):
        shape[ax] = len(coords[n])
    return np.asfortranarray(np.asarray(vals, dtype=np.float64).reshape(shape))


def run_offline_Lagrangian_filter(config: OfflineFilterConfig, output_dtype=np.float32, save: bool = True,
                                  model: Optional[LagrangianFilter] = None, regrid: bool = True) -> FilterOutput:
    """Runs the offline Lagrangian filter configured by `config` on the GPU and returns the combined
    output (also written to `config.output_filename` unless save=False).

    `output_dtype=np.float32` reproduces the reference's JLD2Writer, which rounds every pass output to
    Float32 before the forward+backward sum; pass np.float64 to keep full precision."""
    source = config.source()
    file_times = np.asarray(source.times, dtype=np.float64)
    indices = [n for n, t in enumerate(file_times) if config.T_start <= t <= config.T_end]      # utils:124
    if len(indices) < 2:
        raise ValueError("fewer than two snapshots inside [T_start, T_end]")
    own_model = model is None
    if own_model:
        model = build_model(config)
        log.info("Created model")
    try:
        feeder = SnapshotFeeder(model, source, indices, file_times, config.T_start, config.T_end)
        # forward pass
        feeder.begin(+1)
        feeder.ensure(0.0, 0.0)
        model.init_from_data()
        log.info("Initialised filtered variables")
        fwd, iterations = _run_pass(model, feeder, config, output_dtype)
        # backward pass: same snapshots, reversed in time, velocities negated; maps change sign
        feeder.begin(-1)
        model.negate_maps()
        model.reset_clock()
        bwd, _ = _run_pass(model, feeder, config, output_dtype)
        out = sum_forward_backward_contributions(config, fwd, bwd, iterations, model=model)
        if config.output_original_data:
            out.add_original_data(config, source, indices, file_times)
        if config.map_to_mean and regrid:
            regrid_to_mean_position(config, out)
        if config.compute_Eulerian_filter:
            compute_Eulerian_filter(config, out)
        out.stats = {"h2d_bytes": feeder.bytes_uploaded, "kernel_launches": model.kernel_launches,
                     "kernel": model.kernel_name}
    finally:
        if own_model:
            model.close()
    if save and config.output_filename:
        out.save(config.output_filename)
    return out
