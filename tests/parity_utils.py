"""Helpers shared by the GPU parity tests: run the CPU oracle and the CUDA model on the same inputs."""
import numpy as np

import lf_oracle as lo
import lfb200
from lfb200.model import LagrangianFilter
from lfb200.synthetic import make_series


def rel_l2(a, b):
    d = np.linalg.norm(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    n = np.linalg.norm(np.asarray(b, dtype=np.float64))
    return d / n if n > 0 else d


def make_pair(grid, var_names, vel_names, fp, times, fields, *, strict=False, kernel="auto", with_maps=True,
              advect=True, relax_timescale=None, mask=None, n_slots=None, backward=False, T_end=None,
              timestepper="RungeKutta3", chi=0.1):
    """(oracle, model) loaded with the same series and initialised from data.  `advect`: True / False or a scheme
    name ("WENO3", "Centered4", "UpwindBiased5", ...); `grid` may be an ImmersedBoundaryGrid."""
    T_end = times[-1] if T_end is None else T_end
    immersed = cell_heights = None
    if isinstance(grid, lfb200.ImmersedBoundaryGrid):
        immersed, cell_heights = grid.immersed_flags(), grid.cell_heights()
    o = lo.Oracle(grid.N, grid.H, grid.topology, grid.spacing, len(var_names), "".join(vel_names), fp,
                  with_maps, advect, relax_timescale, mask, timestepper=timestepper, chi=chi, immersed=immersed,
                  cell_heights=cell_heights)
    o.load_series(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in vel_names},
                  [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in var_names], times[0], T_end,
                  backward=backward)
    m = LagrangianFilter(grid, var_names_to_filter=var_names, velocity_names=vel_names, filter_params=fp,
                         with_maps=with_maps, advection=advect, n_slots=n_slots or max(2, len(times)),
                         strict=strict, kernel=kernel, relax_timescale=relax_timescale, mask=mask,
                         timestepper=timestepper, chi=chi)
    for s, t in enumerate(times):
        tt = (T_end - t) if backward else (t - times[0])
        for name in tuple(vel_names) + tuple(var_names):
            m.upload_snapshot(s, name, fields[name][s], tt)
    m.set_direction(-1 if backward else 1)
    return o, m


def oracle_names(var_names, vel_names, fp, with_maps=True):
    """Reference tracer names in the oracle's index order (velocities in u, v, w order)."""
    cfg = type("C", (), dict(var_names_to_filter=tuple(var_names),
                             velocity_names=tuple(v for v in "uvw" if v in vel_names), filter_params=fp,
                             map_to_mean=with_maps, compute_mean_velocities=with_maps, label=""))()
    return lfb200.create_filtered_vars(cfg)


def compare_state(o, m, names, tol, exact=False):
    worst = 0.0
    for q, name in enumerate(names):
        ref, got = o.tracer(q), m.tracer(name)
        if exact:
            assert np.array_equal(ref, got), f"{name}: {np.abs(ref - got).max()} (bitwise mismatch)"
        else:
            e = rel_l2(got, ref)
            worst = max(worst, e)
            assert e <= tol, f"{name}: rel-L2 {e:.3e} > {tol:.1e}"
    return worst
