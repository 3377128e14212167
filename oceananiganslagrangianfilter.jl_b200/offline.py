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
from .post import (FilterOutput, ForwardBackwardCombiner, compute_Eulerian_filter, regrid_to_mean_position,
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
            arr = self.source.snapshot(name, n)          # this rank's slab of the snapshot (SlabSource)
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


def _output_specs(model: LagrangianFilter, config: OfflineFilterConfig):
    """create_output_fields (utils:898-1012): (reference output name, output kind, index) in writer order."""
    ident = "_Eulerian_filtered" if config.advection is None else "_Lagrangian_filtered"
    label = config.label
    specs = [(v + label + ident, _lib.OUT_FILTERED_VAR, q) for q, v in enumerate(config.var_names_to_filter)]
    vels = model.velocities_present
    if config.map_to_mean:
        specs += [("xi_" + vel + label, _lib.OUT_XI, m) for m, vel in enumerate(vels)]
    if config.compute_mean_velocities:
        specs += [(vel + label + ident, _lib.OUT_MEAN_VEL, m) for m, vel in enumerate(vels)]
    return specs


class _DeviceStore:
    """Pass outputs stay on the device as kept arrays (ids): no forward_output / backward_output round trip
    (the reference writes both files and reads them back, run_offline_lagrangian_filter.jl:77-80, post:45-170)."""
    on_device = True

    def __init__(self, model, dtype, n_per_time):
        self.model, self.dtype = model, dtype

    def collect(self, kind, index):
        return self.model.output_keep(kind, index, self.dtype)

    def finish_pass(self, outputs):
        return outputs


class _HostStore:
    """Pass outputs are copied to the host through a BOUNDED ring of page-locked buffers (two output times' worth,
    reused cyclically): the copies overlap with the following time steps, and the pinned footprint does not grow
    with the number of output times."""
    on_device = False

    def __init__(self, model, dtype, n_per_time):
        self.model, self.dtype = model, dtype
        self.mark = model.n_pinned                           # buffers allocated before this store are not ours
        self.ring = [model.pinned_empty(model.center_shape, dtype) for _ in range(2 * max(1, n_per_time))]
        self.free = list(range(len(self.ring)))
        self.pending = []                                   # (holder, ring slot)

    def collect(self, kind, index):
        if not self.free:
            self.drain()
        slot = self.free.pop(0)
        self.model.output_async(kind, index, self.ring[slot])
        holder = [None]
        self.pending.append((holder, slot))
        return holder

    def drain(self):
        self.model.wait_transfers()
        for holder, slot in self.pending:
            holder[0] = np.array(self.ring[slot], order="F")          # leaves the page-locked staging buffer
            self.free.append(slot)
        self.pending = []

    def finish_pass(self, outputs):
        self.drain()
        return {t: {k: h[0] for k, h in d.items()} for t, d in outputs.items()}

    def close(self):
        self.model.free_pinned(since=self.mark)


def _run_pass(model, feeder, config, store, on_output=None):
    """One pass of the filter simulation (run!(simulation), run_offline_lagrangian_filter.jl:83,102): aligned time
    steps, outputs every T_out.  `on_output(t, outputs)` sees every output time as soon as it is queued."""
    T, Δt, T_out = config.T, config.Δt, config.T_out
    specs = _output_specs(model, config)

    def collect(tk=0.0):
        d = {name: store.collect(kind, index) for name, kind, index in specs}
        if on_output is not None:
            on_output(tk, d)
        return d

    feeder.ensure(0.0, 0.0)
    outputs = {0.0: collect()}
    iterations = {0.0: 0}
    for n, step in enumerate(aligned_time_steps(T, Δt, [T_out, T / 10]), 1):
        t = model.time
        feeder.ensure(t, t + step)
        model.step(step)
        t = model.time
        if _is_output_time(t, T_out):
            tk = round(t / T_out) * T_out
            outputs[tk] = collect(tk)
            iterations[tk] = n
    if on_output is None:
        model.wait_transfers()
    return store.finish_pass(outputs), iterations


def build_model(config: OfflineFilterConfig, rank: int = 0, world: int = 1) -> LagrangianFilter:
    """The device model of this rank: the whole grid, or its y-slab (rank r owns rows [r Ny/R, (r+1) Ny/R))."""
    grid = config.grid.slab(rank, world)
    mask = None
    if config.boundary_relaxation:
        mask = evaluate_mask(config, grid)
    return LagrangianFilter(grid, var_names_to_filter=config.var_names_to_filter,
                            velocity_names=config.velocity_names, filter_params=config.filter_params,
                            with_maps=config.map_to_mean or config.compute_mean_velocities,
                            advection=config.advection, n_slots=config.n_slots, device=config.device,
                            timestepper=config.timestepper,
                            strict=config.strict, kernel=config.kernel, label=config.label,
                            relax_timescale=config.relax_timescale if config.boundary_relaxation else None,
                            mask=mask, config=config)


def evaluate_mask(config, grid=None) -> np.ndarray:
    """Pre-evaluates `mask_func(x..., mask_params)` at every cell centre, halos included (the reference
    evaluates it inside the forcing closure at the tracer location, utils:592-697)."""
    g = grid or config.grid
    coords = [g.nodes(ax) for ax in g.nonflat_axes]
    mesh = np.meshgrid(*coords, indexing="ij")
    vals = np.vectorize(lambda *xs: float(config.mask_func(*xs, config.mask_params)))(*mesh)
    shape = [1, 1, 1]
    for n, ax in enumerate(g.nonflat_axes):
        shape[ax] = len(coords[n])
    return np.asfortranarray(np.asarray(vals, dtype=np.float64).reshape(shape))


def _pick_store(model, config, dtype, n_times, prefer):
    """Device store when both passes' outputs (plus one output time of combined and gathered FP64 arrays) fit into
    the free device memory, else the bounded host ring."""
    specs = _output_specs(model, config)
    if prefer == "host":
        return _HostStore
    n = int(np.prod(model.center_shape))
    need = 2 * n_times * len(specs) * n * np.dtype(dtype).itemsize + 3 * len(specs) * n * 8 * max(1, model.nranks)
    free, _ = model.mem_info()
    if prefer == "device" or need < 0.85 * free:
        return _DeviceStore
    if model.nranks > 1:
        raise MemoryError(f"the pass outputs ({need / 2**30:.1f} GiB) do not fit next to the model on this GPU "
                          f"({free / 2**30:.1f} GiB free); the multi-GPU pipeline keeps them on the device: use more "
                          "GPUs or fewer output times")
    log.info("Pass outputs (%.1f GiB) exceed the free device memory (%.1f GiB): staging them on the host",
             need / 2**30, free / 2**30)
    return _HostStore


def run_offline_Lagrangian_filter(config: OfflineFilterConfig, output_dtype=np.float32, save: bool = True,
                                  model: Optional[LagrangianFilter] = None, regrid: bool = True,
                                  rank: Optional[int] = None, world: Optional[int] = None,
                                  output_store: str = "auto", gather_to_root: bool = True, local_source=None,
                                  alloc=None) -> FilterOutput:
    """Runs the offline Lagrangian filter configured by `config` on the GPU(s) and returns the combined
    output (also written to `config.output_filename` unless save=False).

    `output_dtype=np.float32` reproduces the reference's JLD2Writer, which rounds every pass output to
    Float32 before the forward+backward sum; pass np.float64 to keep full precision.

    Multi-GPU (one process per GPU, launched with torchrun; `rank` / `world` default to torch.distributed's):
    every rank owns a y-slab of the grid, both passes run on all ranks (the backward pass starts from the forward
    pass's final state, run_offline_lagrangian_filter.jl:86-102), the pass outputs stay on the device, are combined
    there slab by slab, gathered by output time (time k on rank k mod R) and remapped to the mean position on their
    owner: the remap is sharded by output time.  Rank 0 receives the complete output and
    writes the file; with gather_to_root=False every rank returns the output times it owns.

    `output_store`: "device" keeps the pass outputs on the device, "host" stages them on the host through a bounded
    page-locked ring, "auto" picks the device when they fit.  `local_source`: an input source that already yields this
    rank's y-slab of every snapshot (instead of cutting it out of the global arrays of `config`'s source).
    `alloc(shape, dtype)`: allocator of the host arrays the combined outputs are fetched into (device store)."""
    from . import distributed as D
    rank, world = D.rank_world(rank, world)
    source = config.source()
    file_times = np.asarray(source.times, dtype=np.float64)
    indices = [n for n, t in enumerate(file_times) if config.T_start <= t <= config.T_end]      # utils:124
    if len(indices) < 2:
        raise ValueError("fewer than two snapshots inside [T_start, T_end]")
    if file_times[indices[0]] != config.T_start or file_times[indices[-1]] != config.T_end:
        raise ValueError(f"T_start={config.T_start} and T_end={config.T_end} must coincide with snapshot times of the "
                         f"original data (nearest inside: {file_times[indices[0]]}, {file_times[indices[-1]]}): the "
                         "filter is initialised from the snapshot at T_start and the backward pass from the one at T_end")
    own_model = model is None
    if own_model:
        model = build_model(config, rank, world)
        D.init_model_comm(model, rank, world)
        log.info("Created model")
    elif model.nranks != world:
        raise ValueError(f"the model passed in belongs to a {model.nranks}-rank communicator, not {world}")
    store = None
    try:
        feeder = SnapshotFeeder(model, local_source or D.SlabSource(source, config.grid, rank, world), indices, file_times,
                                config.T_start, config.T_end)
        n_times = int(round(config.T / config.T_out)) + 1
        store = _pick_store(model, config, output_dtype, n_times, output_store)(model, output_dtype,
                                                                                  len(_output_specs(model, config)))
        import time as _time
        phases, t_mark = {}, _time.perf_counter()

        def phase(name, sync=True):
            nonlocal t_mark
            if sync:
                model.sync()
            now = _time.perf_counter()
            phases[name] = now - t_mark
            t_mark = now

        if store.on_device:
            # the memory of the kept outputs is taken now, not between queued time steps: one pass's outputs (the backward
            # pass reuses the forward pass's blocks as they are combined) and two output times of combined Float64 fields
            n_spec = len(_output_specs(model, config))
            model.kept_reserve((n_times + 1) * n_spec, output_dtype)
            model.kept_reserve(2 * n_spec, np.float64)
            if world > 1 and config.map_to_mean and regrid:
                # the global arrays of the output times this rank owns (gathered during the backward pass)
                g = config.grid
                n_owned = len([k for k in range(n_times) if D.owner_of_time(k, world) == rank])
                model.kept_reserve(n_owned * n_spec, np.float64, shape=g.center_shape())
        # forward pass
        feeder.begin(+1)
        feeder.ensure(0.0, 0.0)
        model.init_from_data()               # needs the first snapshot only: the other uploads overlap with the steps
        log.info("Initialised filtered variables")
        phase("queue_uploads_and_init_s", sync=False)      # the uploads keep running under the first time steps
        fwd, iterations = _run_pass(model, feeder, config, store)
        phase("forward_pass_s")
        # backward pass: same snapshots, reversed in time, velocities negated; maps change sign
        feeder.begin(-1)
        model.negate_maps()
        model.reset_clock()
        do_regrid = config.map_to_mean and regrid
        if store.on_device:
            # every forward time is combined (and its download queued) as soon as its backward partner exists
            combiner = ForwardBackwardCombiner(config, fwd, iterations, model, rank=rank, world=world, regrid=do_regrid,
                                               alloc=alloc)
            _run_pass(model, feeder, config, store, on_output=combiner.feed)
            phase("backward_pass_with_combination_s")
            out = combiner.finish()
            phase("last_downloads_s")
        else:
            bwd, _ = _run_pass(model, feeder, config, store)
            phase("backward_pass_s")
            out = sum_forward_backward_contributions(config, fwd, bwd, iterations, model=model, rank=rank, world=world,
                                                     regrid=do_regrid, alloc=alloc)
            phase("combination_and_remap_s")
        stats = {"phases": phases,"h2d_bytes": feeder.bytes_uploaded, "kernel_launches": model.kernel_launches,
                 "kernel": model.kernel_name, "output_store": "device" if store.on_device else "host",
                 "rank": rank, "world": world}
    finally:
        if store is not None and hasattr(store, "close"):
            store.close()
        if own_model:
            model.close()                # frees the pooled blocks of the kept outputs too
        # a model passed in keeps its pool of output blocks for the next run (model.kept_trim() returns the memory)
    if world > 1 and gather_to_root:
        out = D.merge_outputs_on_root(out, rank, world, config.grid)
        if rank != 0:
            out.stats = stats
            return out
    complete = world == 1 or gather_to_root
    if complete:
        if config.output_original_data:
            out.add_original_data(config, source, indices, file_times)
        if config.compute_Eulerian_filter:
            compute_Eulerian_filter(config, out)
    out.stats = stats
    if save and config.output_filename and complete:
        out.save(config.output_filename)
    return out
