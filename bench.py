#!/usr/bin/env python
"""Benchmark of the filter-PDE hot path (BASELINE.json metric: filter cell-updates/s, tracer x cell / s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size NX NY NZ]

One "step" = one full RK3 time step (3 fused stage launches) of every filter tracer on the synthetic
3-D 1024x1024x128 FP64 grid (Periodic, Periodic, Bounded), N=2 Butterworth, filtered variables (T, b),
velocities (u, v, w) => 10 tracers (BASELINE.json configs[4]).  With N > 1
(torchrun, one rank per GPU) the same grid is split into N y-slabs with an NCCL halo exchange per
stage: strong scaling, `value` = whole-job cell-updates / max-over-ranks device time.

  value      stepping only, inputs resident in HBM, CUDA events on the library's compute stream
  e2e        the same metric through the public API, run_offline_Lagrangian_filter(config), with HOST buffers:
             Float32 snapshot uploads from pinned host memory, forward + backward pass, device-side combination,
             combined Float64 fields fetched to pinned host memory
  roofline   dominant kernel (the fused stage kernel): algorithmic bytes / mean launch duration vs the
             measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU restatement of the reference numerics (oracle/) on a bounded
             sample of the same workload, all host threads (Julia + Oceananigans are not installed, so
             the reference itself cannot run; kind = "port")
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "filter cell-updates/s"
UNIT = "tracer*cell/s"
VAR_NAMES, VEL_NAMES = ("T", "b"), ("u", "v", "w")
N_TRACERS, N_INPUTS = 10, 5
T_S = 1.0                      # snapshot spacing
DT = T_S / 8                   # filter time step
N_SNAP = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lfb200")
    ap.add_argument("--size", type=int, nargs=3, default=[1024, 1024, 128])
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-size", type=int, nargs=3, default=[256, 128, 64])
    ap.add_argument("--remap", action="store_true",
                    help="time the remap to the mean position (post-processing, one output time) instead of the stepping")
    ap.add_argument("--remap-fields", type=int, default=5)
    ap.add_argument("--pipeline-remap", action="store_true",
                    help="also run the pipeline once WITH the remap to the mean position (sharded by output time) and report its wall time")
    return ap.parse_args()


def make_grid(size):
    import lfb200
    return lfb200.RectilinearGrid(size=tuple(size), extent=tuple(float(s) for s in size),
                                  topology=("Periodic", "Periodic", "Bounded"))


def filter_params():
    import lfb200
    return lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / (4 * T_S))


def algorithmic_bytes_per_cell_step():
    """8 (10 n_T + 6 n_in) bytes per cell per RK3 step."""
    return 8 * (10 * N_TRACERS + 6 * N_INPUTS)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 7:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads():
    """Host threads this process may use (its CPU affinity mask, else the core count)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_rate(size, steps, warmup):
    """cell-updates/s of the CPU restatement (oracle/lf_oracle.c, OpenMP) on a `size` sample of the workload with
    the same filter, tracers and synthetic fields.  The OpenMP thread count is set explicitly (launchers such as
    torchrun export OMP_NUM_THREADS=1) and the count actually in effect is returned."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lf_oracle as lo
    threads = lo.set_threads(host_threads())
    from lfb200.synthetic import make_series
    g = make_grid(size)
    times = [0.0, T_S, 2 * T_S]
    fields = make_series(g, times, var_names=VAR_NAMES, velocity_names=VEL_NAMES, T_period=4 * T_S)
    fp = filter_params()
    o = lo.Oracle(g.N, g.H, g.topology, g.spacing, len(VAR_NAMES), "".join(VEL_NAMES), fp, True, True)
    o.load_series(times, {k: [a.astype(np.float64) for a in fields[k]] for k in VEL_NAMES},
                  [[a.astype(np.float64) for a in fields[v]] for v in VAR_NAMES], 0.0, 2 * T_S, backward=False)
    o.init_from_data()
    for _ in range(warmup):
        o.step(DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT)
    dt = time.perf_counter() - t0
    o.close()
    cells = size[0] * size[1] * size[2]
    return cells * N_TRACERS * steps / dt, dt / steps * 1e3, threads


def run_reference(args, rank):
    if rank != 0:
        return
    size = args.cpu_size
    rate, ms, cores = cpu_port_rate(size, args.steps, args.warmup)
    sample = (f"{size[0]}x{size[1]}x{size[2]} sub-grid of the workload, {args.steps} RK3 steps after {args.warmup} warm-up, "
              f"OpenMP with omp_get_max_threads() = {cores}; CPU restatement of the reference numerics (Julia/Oceananigans absent)")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    nx, ny, nz = args.size
    return {"workload": f"synthetic 3D {nx}x{ny}x{nz} FP64 (Periodic,Periodic,Bounded), N=2 Butterworth offline filter, "
                        f"vars (T,b) + maps (u,v,w) = {N_TRACERS} tracers, WENO5-Z + RK3, dt = T_s/8",
            "grid": [nx, ny, nz], "tracers": N_TRACERS, "inputs": N_INPUTS, "snapshots": N_SNAP,
            "decomposition": f"{world} y-slab(s)", "l2": "inputs (>= 25 GB of fields per GPU at full size) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ remap
def run_remap(args, local):
    """regrid_to_mean_position! for ONE output time of the workload's grid (post:327-762): synthetic, smooth
    displacement maps of about half a cell, `--remap-fields` fields sharing one tetrahedron search; inputs resident on
    the device, result arrays on the host.  Prints its own JSON line (metric: remap nodes/s)."""
    import ctypes as C
    import torch
    import lfb200
    from lfb200 import _lib
    g = make_grid(args.size)
    dev = torch.device("cuda", local)
    shape = tuple(reversed(g.center_shape()))                 # torch (z, y, x) = Fortran (x, y, z)
    z, y, x = torch.meshgrid(*[torch.arange(n, device=dev, dtype=torch.float64) for n in shape], indexing="ij")
    two_pi = 2 * np.pi
    X, Y, Z = (x - g.Hx + 0.5) / g.Nx, (y - g.Hy + 0.5) / g.Ny, (z - g.Hz + 0.5) / g.Nz
    xi = [0.45 * g.spacing[0] * torch.sin(two_pi * (X + 2 * Y)) * torch.cos(two_pi * Z),
          0.40 * g.spacing[1] * torch.cos(two_pi * (2 * X - Y)) * torch.sin(two_pi * Z + 0.3),
          0.35 * g.spacing[2] * torch.sin(two_pi * (X + Y)) * torch.sin(np.pi * Z)]
    fields = [torch.sin(two_pi * (X * (q + 1) + Y)) + Z * q for q in range(args.remap_fields)]
    del x, y, z, X, Y, Z
    torch.cuda.synchronize()
    L = _lib.load()
    d = _lib.RemapDesc()
    for ax in range(3):
        d.size[ax], d.halo[ax] = g.N[ax], g.H[ax]
        d.topology[ax] = _lib.TOPOLOGY[g.topology[ax]]
        d.origin[ax], d.spacing[ax] = g.origin[ax], g.spacing[ax]
        d.has_map[ax] = 1
    d.npad, d.device = 5, local
    # results into page-locked host arrays, as the pipeline fetches them (pageable arrays halve the download rate)
    n_out = int(np.prod(g.center_shape()))
    pinned, outs = [], []
    for _ in fields:
        ptr = C.c_void_p()
        _lib.check(L.lfb_host_alloc(n_out * 8, C.byref(ptr)))
        pinned.append(ptr)
        outs.append(np.frombuffer((C.c_char * (n_out * 8)).from_address(ptr.value), dtype=np.float64,
                                  count=n_out).reshape(g.center_shape(), order="F"))
    xp = (C.c_void_p * 3)(*[t.data_ptr() for t in xi])
    fp_ = (C.c_void_p * len(fields))(*[t.data_ptr() for t in fields])
    op = (C.c_void_p * len(fields))(*[o.ctypes.data for o in outs])
    times = []
    for rep in range(1 + max(1, args.steps // 8)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.lfb_remap_to_mean(C.byref(d), xp, len(fields), fp_, op))
        times.append(time.perf_counter() - t0)
    best = min(times[1:]) if len(times) > 1 else times[0]
    nodes = g.Nx * g.Ny * g.Nz
    inner = outs[0][g.Hx:-g.Hx, g.Hy:-g.Hy, g.Hz:-g.Hz]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_node = (3 + 2 * len(fields)) * 8            # read 3 displacement maps once, per field read f and write f_at_mean
    achieved = nodes * bytes_per_node / best / 1e9
    line = {"metric": "remap nodes/s", "value": nodes / best,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_node": bytes_per_node,
                         "note": "the search is FP64-latency bound (17 % FP64 pipe, 31 % issue slots at 8 warps per SM, "
                                 "profiles/r02_ncu_remap3d_512x512x128.txt), nowhere near the HBM roof: several hundred candidate "
                                 "evaluations per node against 104 B of algorithmic traffic"}, "unit": "node/s (one output time, search shared by the fields)",
            "n_gpus": 1, "seconds_per_output_time": best, "fields": len(fields), "grid": list(args.size),
            "nan_fraction": float(np.isnan(inner).mean()), "higher_is_better": True, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"remap to the mean position of {len(fields)} fields on the synthetic 3D "
                                   f"{args.size[0]}x{args.size[1]}x{args.size[2]} grid (Periodic,Periodic,Bounded), displacement maps of "
                                   "about 0.45 cells, npad = 5; inputs on the device, results in page-locked host arrays"}}
    print(json.dumps(line), flush=True)
    del inner, outs
    for ptr in pinned:
        L.lfb_host_free(ptr)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import lfb200
    from lfb200.model import LagrangianFilter
    from lfb200.synthetic import make_snapshot
    from lfb200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lfb200 has no CPU fallback")
    torch.cuda.set_device(local)
    if args.remap:
        if rank == 0:
            run_remap(args, local)
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gglobal = make_grid(args.size)
    g = gglobal.slab(rank, world)
    fp = filter_params()
    n_slots = N_SNAP
    m = LagrangianFilter(g, var_names_to_filter=VAR_NAMES, velocity_names=VEL_NAMES, filter_params=fp, n_slots=n_slots,
                         device=local, kernel=args.kernel)
    if world > 1:
        uid = [LagrangianFilter.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        m.comm_init(uid[0], rank, world)

    # ---- synthetic series generated on the GPU, kept on the host as pinned Float32 (the on-disk precision)
    times = [n * T_S for n in range(N_SNAP)]
    names = list(VEL_NAMES) + list(VAR_NAMES)
    host = {}
    for s, t in enumerate(times):
        snap = make_snapshot(g, t, 4 * T_S, var_names=VAR_NAMES, velocity_names=VEL_NAMES, xp=torch,
                             device=torch.device("cuda", local), dtype=np.float32, global_grid=gglobal)
        for name in names:
            pinned = torch.empty(snap[name].shape, dtype=torch.float32, pin_memory=True)
            pinned.copy_(snap[name])
            host[(name, s)] = pinned.numpy().T          # Fortran (ex, ey, ez) view of the pinned buffer
        del snap
    torch.cuda.synchronize()

    def upload(slot, s):
        nbytes = 0
        for name in names:
            m.upload_snapshot(slot, name, host[(name, s)], times[s])
            nbytes += host[(name, s)].nbytes
        return nbytes

    for s in range(N_SNAP):
        upload(s, s)
    m.init_from_data()
    m.sync()

    def barrier():
        m.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        m.sync()

    # multi-GPU correctness the driver can see: the whole pipeline on small grids, slabs vs one GPU (bit-exact with the
    # strict kernel, <= 1e-12 with the TMA kernel), before anything is timed
    mgpu_parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from mgpu_pipeline_check import run_cases
        ok, lines = run_cases(rank, world, local, cases=(("Periodic", "generic", True, (24, 8 * world, 10), True),
                                                          ("Periodic", "auto", False, (40, 16 * world, 16), False)))
        mgpu_parity = {"ok": bool(ok), "cases": lines}

    cells_global = args.size[0] * args.size[1] * args.size[2]
    cells_local = g.Nx * g.Ny * g.Nz
    K, W = args.steps, args.warmup
    max_steps = int((times[-1] - 0.0) / DT)
    if K + W > max_steps:
        raise SystemExit(f"--steps + --warmup must be <= {max_steps} (the synthetic series covers {times[-1]} T_s)")

    # ---- value: stepping only, inputs resident ------------------------------------------------------
    for _ in range(W):
        m.step(DT)
    barrier()
    m.stage_time_ms(reset=True)
    launches0 = m.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m.timer_start()
    for _ in range(K):
        m.step(DT)
    ms = m.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = m.kernel_launches - launches0
    stage_ms, stage_n = m.stage_time_ms(reset=True)
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = cells_global * N_TRACERS * K / (ms_max * 1e-3)

    # ---- e2e: the whole offline pipeline through the public API, host buffers inside the timed region --------
    # run_offline_Lagrangian_filter(config): Float32 snapshots uploaded from page-locked host memory, forward pass,
    # backward pass (both with their per-output-time outputs kept on the device), device-side combination, and the
    # combined Float64 fields fetched into page-locked host arrays (every rank its y-slab).  The remap to the mean
    # position is post-processing and is timed separately (--remap).
    e2e = None
    if not args.no_e2e:
        from lfb200.offline import aligned_time_steps, run_offline_Lagrangian_filter

        class HostSeries:
            """The synthetic series as this rank's input source: slab parent arrays in page-locked host memory."""
            def __init__(self):
                self.times = np.asarray(times)

            def has(self, name):
                return name in names

            def snapshot(self, name, n):
                return host[(name, n)]

        src = HostSeries()
        cfg = lfb200.OfflineFilterConfig(data=src, grid=gglobal, var_names_to_filter=VAR_NAMES, velocity_names=VEL_NAMES,
                                         filter_params=fp, dt=DT, T_out=T_S, device=local, n_slots=n_slots,
                                         kernel=args.kernel, output_original_data=False)
        n_pipe_steps = 2 * len(aligned_time_steps(cfg.T, cfg.dt, [cfg.T_out, cfg.T / 10]))
        n_out = len(VAR_NAMES) + 2 * len(VEL_NAMES)
        n_times = int(round(cfg.T / cfg.T_out)) + 1

        class PinnedPool:
            """Destination arrays of the combined outputs, allocated before the timed region like any long-lived
            I/O buffer (a consumer would write them to disk)."""
            def __init__(self, count, shape, dtype):
                self.arrays = [m.pinned_empty(shape, dtype) for _ in range(count)]
                self.next, self.bytes = 0, 0

            def __call__(self, shape, dtype):
                if tuple(shape) != tuple(self.arrays[0].shape) or self.next >= len(self.arrays):
                    a = np.empty(shape, dtype=dtype, order="F")       # e.g. the gathered global arrays of the remap at N > 1
                    self.bytes += a.nbytes
                    return a
                a = self.arrays[self.next]
                self.next += 1
                self.bytes += a.nbytes
                return a

        n_at_mean = (len(VAR_NAMES) + len(VEL_NAMES)) if args.pipeline_remap else 0     # remapped fields per output time
        pinned_mark = m.n_pinned
        pool = PinnedPool((n_out + n_at_mean) * n_times, m.center_shape, np.float64)
        # device blocks of the kept pass outputs: taken before the timed region, like the model's own fields
        m.kept_reserve((n_times + 1) * n_out, np.float32)
        m.kept_reserve(2 * n_out, np.float64)
        barrier()
        t0 = time.perf_counter()
        out = run_offline_Lagrangian_filter(cfg, save=False, regrid=False, model=m, rank=rank, world=world,
                                            output_store="device", gather_to_root=False, local_source=src, alloc=pool)
        barrier()
        wall = time.perf_counter() - t0
        assert len(out.times) == n_times and out.stats["output_store"] == "device"
        t_w = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_w, op=dist.ReduceOp.MAX)
        e2e = {"value": cells_global * N_TRACERS * n_pipe_steps / float(t_w.item()), "unit": UNIT,
               "h2d_bytes_per_step": out.stats["h2d_bytes"] / n_pipe_steps, "d2h_bytes_per_step": pool.bytes / n_pipe_steps,
               "steps": n_pipe_steps, "wall_s": float(t_w.item()), "phases_rank0": out.stats["phases"],
               "note": "run_offline_Lagrangian_filter(config) on host buffers: Float32 snapshot uploads from page-locked host "
                       "memory, forward + backward pass with aligned time steps (outputs every T_out kept on the device), "
                       "device-side forward+backward combination, combined Float64 fields fetched to page-locked host "
                       "arrays (per rank: its y-slab); the device blocks of the kept outputs and the page-locked result "
                       "arrays are allocated before the timed region, like the model; wall clock between device syncs, max "
                       "over ranks; the remap to the mean position is not included"}
        del out
        if args.pipeline_remap:
            pool.next, pool.bytes = 0, 0
            if world > 1:
                # the owners of output times receive GLOBAL arrays: page-locked destinations for those (the slab arrays of
                # the run above are released first)
                from lfb200.distributed import owner_of_time
                m.free_pinned(since=pinned_mark)
                n_owned = len([k for k in range(n_times) if owner_of_time(k, world) == rank])
                pool = PinnedPool((n_out + n_at_mean) * n_owned, gglobal.center_shape(), np.float64) if n_owned else pool
                if not n_owned:
                    pool.arrays = [np.empty((1, 1, 1))]
            barrier()
            t0 = time.perf_counter()
            out = run_offline_Lagrangian_filter(cfg, save=False, regrid=True, model=m, rank=rank, world=world,
                                                output_store="device", gather_to_root=False, local_source=src, alloc=pool)
            barrier()
            t_w = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t_w, op=dist.ReduceOp.MAX)
            e2e["with_remap"] = {"wall_s": float(t_w.item()), "output_times": n_times, "phases_rank0": out.stats["phases"],
                                 "fields_at_mean": sum(1 for k in out.keys() if k.endswith("_at_mean")),
                                 "note": "the same pipeline with regrid_to_mean_position: output time k gathered on rank k mod N and "
                                         "remapped there (T, b and the three mean velocities); rank 0's share of the output times"}
            del out

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        stage_avg_ms = stage_ms / max(stage_n, 1)
        bytes_per_launch = cells_local * algorithmic_bytes_per_cell_step() / 3.0
        achieved = bytes_per_launch / (stage_avg_ms * 1e-3) / 1e9
        # DRAM bytes per launch of this kernel on this grid from the committed `ncu --set full` capture (it cannot be
        # measured from inside the bench); null for configurations that were not captured
        traffic = None
        try:
            for cap in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
                if cap["kernel"] == m.kernel_name and cap["grid"] == list(args.size) and cap["n_gpus"] == world:
                    traffic = cap["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": m.kernel_name, "launch_ms": stage_avg_ms, "peak_source": which,
                    "algorithmic_bytes_per_launch": bytes_per_launch,
                    "note": "FP64 WENO5-Z is bound by the warp schedulers' issue port (an FP64 instruction holds it for two "
                            "cycles) before it is HBM bound: with no other instruction at all the ceiling would be 61 % of "
                            "the HBM peak at 1965 MHz and 55 % at the ~1750 MHz the power cap allows; see DESIGN.md section 5"}
        cpu = None
        if not args.no_cpu and world == 1:
            rate, cms, cores = cpu_port_rate(args.cpu_size, 24, 2)
            cs = args.cpu_size
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cs[0]}x{cs[1]}x{cs[2]} sub-grid of the workload, 24 RK3 steps after 2 warm-up, OpenMP with "
                             f"omp_get_max_threads() = {cores} ({cms:.0f} ms/step)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu}
        if mgpu_parity is not None:
            line["mgpu_parity"] = mgpu_parity
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
