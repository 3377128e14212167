"""GPU parity of the numerics options beyond the reference's defaults: advection
schemes other than WENO(order=5), the QuasiAdamsBashforth2 time stepper and immersed boundaries (whole and partial
cells), all through the C ABI against the CPU oracle on the same inputs.

No reference vector pins these paths (the reference's tests run the defaults only): the oracle
restates Oceananigans' published schemes (oracle/lf_oracle.c: scheme_chain / face_value / ab2_time_step /
mask_immersed) and tests/test_oracle_numerics.py checks their defining properties on the CPU.

Tolerances: strict handles bit-exact; production handles rel-L2 <= 1e-12 per tracer after the steps taken.
"""
import numpy as np
import pytest

import lfb200
from lfb200.synthetic import make_series
from parity_utils import compare_state, make_pair, oracle_names

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12
SCHEMES = ["WENO3", "Centered2", "Centered4", "UpwindBiased1", "UpwindBiased3", "UpwindBiased5"]


def _grid(size, topology, halo=None):
    nonflat = [n for n, t in zip(size, topology) if t != "Flat"]
    return lfb200.RectilinearGrid(size=tuple(nonflat), extent=tuple(float(n) for n in nonflat), topology=topology,
                                  halo=halo)


def _run(g, var_names, vel_names, N, strict, steps=(0.125,) * 4, **kw):
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 4.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, strict=strict, **kw)
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in steps:
        o.step(dt); m.step(dt)
    worst = compare_state(o, m, names, TOL_STEP, exact=strict)
    assert m.time == o.time
    kernel = m.kernel_name
    m.close(); o.close()
    return worst, kernel


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_advection_schemes_vs_oracle(scheme, strict):
    """Every scheme OfflineFilterConfig.advection admits (OfflineLagrangianFilter.jl:139,246) with its buffer-scheme
    chain next to the walls: 2-D (P,F,B) like the reference's examples, and a 3-D box with walls on every axis."""
    _run(_grid((24, 1, 14), ("Periodic", "Flat", "Bounded")), ("b",), ("u", "w"), 2, strict, advect=scheme)
    _run(_grid((12, 9, 10), ("Bounded", "Bounded", "Bounded")), ("T", "b"), ("u", "v", "w"), 2, strict, advect=scheme)


@pytest.mark.parametrize("strict", [True, False])
def test_ab2_time_stepper_vs_oracle(strict):
    """QuasiAdamsBashforth2 (lagrangian_filter_ab2_step.jl:24-51): Euler at the first step and whenever dt changes,
    chi = 0.1 otherwise; generic kernel."""
    steps = (0.1, 0.1, 0.1, 0.05, 0.05, 0.1)
    _run(_grid((20, 1, 12), ("Periodic", "Flat", "Bounded")), ("b",), ("u", "w"), 2, strict, steps=steps,
         timestepper="QuasiAdamsBashforth2")
    _run(_grid((16, 8, 10), ("Periodic", "Periodic", "Bounded")), ("T",), ("u", "v", "w"), 1, strict, steps=steps,
         timestepper="QuasiAdamsBashforth2", chi=0.05, kernel="generic")


def test_ab2_on_the_production_kernel():
    """AB2 is one fused stage per step with other coefficients: the TMA z-march kernel runs it unchanged."""
    steps = (0.1, 0.1, 0.1, 0.05, 0.05, 0.1, 0.1)
    worst, kernel = _run(_grid((40, 12, 16), ("Periodic", "Periodic", "Bounded")), ("T", "b"), ("u", "v", "w"), 2, False,
                         steps=steps, timestepper="QuasiAdamsBashforth2")
    assert kernel == "ztma"


def _bump(g, height, width, x0, y_dep=False):
    zb = g.origin[2]
    if y_dep:
        return lambda x, y: zb + height * np.exp(-((x - x0) ** 2 + (y - 0.5 * g.Ly) ** 2) / (2 * width ** 2))
    return lambda x: zb + height * np.exp(-(x - x0) ** 2 / (2 * width ** 2))


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("bottom_kind", ["GridFittedBottom", "PartialCellBottom"])
def test_immersed_boundary_lee_wave_shape(bottom_kind, strict):
    """The lee-wave configuration's shape (examples/offline_filter_lee_wave.jl:49-53: (Periodic, Flat, Bounded),
    halo = (4, 4), a Gaussian bump, T and b filtered, N = 4): tracers zero inside the bump, no flux through its faces,
    buffer schemes next to it; whole cells and the example's partial cells."""
    g0 = lfb200.RectilinearGrid(size=(48, 20), x=(-24.0, 24.0), z=(-10.0, 0.0), topology=("Periodic", "Flat", "Bounded"),
                                halo=(4, 4))
    g = lfb200.ImmersedBoundaryGrid(g0, getattr(lfb200, bottom_kind)(_bump(g0, 4.3, 5.0, -6.0)))
    flags = g.immersed_flags()
    assert 0 < flags[4:-4, :, 4:-4].sum() < 0.3 * 48 * 20
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=("T", "b"), velocity_names=("u", "w"), T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=4, freq_c=2 * np.pi / 4.0)
    o, m = make_pair(g, ("T", "b"), ("u", "w"), fp, times, fields, strict=strict)
    names = oracle_names(("T", "b"), ("u", "w"), fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.125,) * 5:
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP, exact=strict)
    inside = flags[4:-4, :, 4:-4] != 0
    for name in names:
        assert np.all(m.tracer(name)[4:-4, :, 4:-4][inside] == 0.0)
    m.close(); o.close()


@pytest.mark.parametrize("strict", [True, False])
def test_immersed_boundary_3d(strict):
    """A 3-D bump (x and y dependent bottom) on a (Periodic, Periodic, Bounded) grid, partial cells, with the
    backward pass's negated velocities."""
    g0 = lfb200.RectilinearGrid(size=(20, 12, 12), extent=(20.0, 12.0, 12.0), topology=("Periodic", "Periodic", "Bounded"))
    g = lfb200.ImmersedBoundaryGrid(g0, lfb200.PartialCellBottom(_bump(g0, 5.2, 3.0, 8.0, y_dep=True)))
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=("b",), velocity_names=("u", "v", "w"), T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 4.0)
    o, m = make_pair(g, ("b",), ("u", "v", "w"), fp, times, fields, strict=strict, backward=True)
    assert m.kernel_name.startswith("generic")          # immersed grids run on the generic kernel
    names = oracle_names(("b",), ("u", "v", "w"), fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.125,) * 4:
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP, exact=strict)
    m.close(); o.close()


@pytest.mark.parametrize("size,topology,vel_names,N,backward", [
    ((40, 20, 16), ("Periodic", "Periodic", "Bounded"), ("u", "v", "w"), 2, False),      # a 3-D bump under a periodic ocean
    ((36, 18, 12), ("Periodic", "Bounded", "Bounded"), ("u", "v", "w"), 1, True),        # channel, single exponential, backward pass
    ((70, 9, 14), ("Bounded", "Periodic", "Bounded"), ("u", "v", "w"), 4, False),        # walls in x, partial tiles, two coefficient sets
    ((96, 1, 20), ("Periodic", "Flat", "Bounded"), ("u", "w"), 2, False),                # the lee-wave shape on the folded kernel
])
def test_immersed_boundary_on_the_tma_kernel(size, topology, vel_names, N, backward):
    """Whole-cell immersed grids (GridFittedBottom) on the TMA z-march kernel: the scheme of every face and the immersed
    flag of every cell travel as one more TMA tile of packed codes (lfb_build_face_codes), fluxes through faces that touch
    the boundary vanish, cells inside stay zero; against the CPU oracle on the same inputs."""
    nonflat = [n for n, t in zip(size, topology) if t != "Flat"]
    g0 = lfb200.RectilinearGrid(size=tuple(nonflat), extent=tuple(float(n) for n in nonflat), topology=topology)
    y_dep = topology[1] != "Flat"
    g = lfb200.ImmersedBoundaryGrid(g0, lfb200.GridFittedBottom(_bump(g0, 0.45 * g0.Lz, 0.12 * g0.Lx, g0.origin[0] + 0.4 * g0.Lx, y_dep=y_dep)))
    flags = g.immersed_flags()
    H = g.H
    inner = (slice(H[0], -H[0]), slice(H[1], -H[1]) if H[1] else slice(None), slice(H[2], -H[2]))
    assert 0 < flags[inner].sum() < 0.35 * flags[inner].size
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=("T", "b"), velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, ("T", "b"), vel_names, fp, times, fields, kernel="ztma", backward=backward)
    assert m.kernel_name == "ztma"
    names = oracle_names(("T", "b"), vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    for name in names[:2]:
        assert np.all(m.tracer(name)[inner][flags[inner] != 0] == 0.0)
    m.close(); o.close()


def test_offline_pipeline_with_options_vs_oracle():
    """run_offline_Lagrangian_filter end to end (forward + backward pass, combination) with the options beyond the
    defaults: an immersed lee-wave-shaped grid with partial cells (examples/offline_filter_lee_wave.jl:49-53), and AB2 with
    UpwindBiased(order=3); FP64 pass outputs against the oracle's pipeline, north-star tolerance 1e-10 per field."""
    import lf_oracle as lo
    g0 = lfb200.RectilinearGrid(size=(40, 16), x=(-20.0, 20.0), z=(-8.0, 0.0), topology=("Periodic", "Flat", "Bounded"),
                                halo=(4, 4))
    gi = lfb200.ImmersedBoundaryGrid(g0, lfb200.PartialCellBottom(_bump(g0, 3.3, 4.0, -5.0)))
    cases = [(gi, {}, dict(immersed=gi.immersed_flags(), cell_heights=gi.cell_heights())),
             (g0, dict(advection=lfb200.UpwindBiased(order=3), timestepper="QuasiAdamsBashforth2"),
              dict(advect="UpwindBiased3", timestepper="QuasiAdamsBashforth2"))]
    times = [0.0, 0.5, 1.0, 1.5, 2.0]
    for g, cfg_kw, oracle_kw in cases:
        fields = make_series(g, times, var_names=("T", "b"), velocity_names=("u", "w"), T_period=4.0, amplitude=0.6)
        data = lfb200.InputData(times, {k: [np.asarray(a, dtype=np.float64) for a in v] for k, v in fields.items()})
        cfg = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("T", "b"), velocity_names=("u", "w"), N=2,
                                         freq_c=2 * np.pi / 2.0, dt=0.1, T_out=0.5, n_slots=5, npad=3,
                                         output_original_data=False, **cfg_kw)
        out = lfb200.run_offline_Lagrangian_filter(cfg, output_dtype=np.float64, save=False)
        assert out.stats["kernel"] == "generic"
        fp = lo.bw2_params(2, 2 * np.pi / 2.0)
        ref = lo.run_offline(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in ("u", "w")},
                             [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in ("T", "b")], ["T", "b"], N=g.N, H=g.H,
                             topology=g.topology, delta=g.spacing, filter_params=fp, dt=0.1, T_out=0.5, velocities="uw",
                             float32_outputs=False, **oracle_kw)
        for name in ("T_Lagrangian_filtered", "b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered", "w_Lagrangian_filtered"):
            a, b = out.stacked(name), np.stack(ref[name])
            e = np.linalg.norm(a - b) / np.linalg.norm(b)
            assert e <= 1e-10, (name, e)
        at_mean = out.stacked("b_Lagrangian_filtered_at_mean")
        if g is gi:
            # the cells _mask_immersed blanks (and the nodes whose simplex touches them) come back as NaN, the rest finite
            blank = gi.below_bottom_mask()
            assert np.all(np.isnan(at_mean[:, blank])) and np.isfinite(at_mean[:, ~blank]).mean() > 0.7
        else:
            assert np.all(np.isfinite(at_mean))


def test_numerics_option_errors():
    g = _grid((16, 1, 12), ("Periodic", "Flat", "Bounded"))
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=1.0)
    with pytest.raises(ValueError):
        lfb200.LagrangianFilter(g, var_names_to_filter=("b",), velocity_names=("u", "w"), filter_params=fp, advection="WENO7")
    with pytest.raises(ValueError):
        lfb200.LagrangianFilter(g, var_names_to_filter=("b",), velocity_names=("u", "w"), filter_params=fp, timestepper="Euler")
    # the TMA kernel only carries the default scheme
    g3 = _grid((32, 16, 12), ("Periodic", "Periodic", "Bounded"))
    with pytest.raises(lfb200._lib.LfbError):
        lfb200.LagrangianFilter(g3, var_names_to_filter=("b",), velocity_names=("u", "v", "w"), filter_params=fp,
                                advection="Centered4", kernel="ztma")
    m = lfb200.LagrangianFilter(g3, var_names_to_filter=("b",), velocity_names=("u", "v", "w"), filter_params=fp,
                                advection=lfb200.UpwindBiased(order=5))
    assert m.kernel_name == "generic"
    m.close()
