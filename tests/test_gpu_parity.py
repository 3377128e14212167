"""GPU parity tests proper: the CUDA path behind the C ABI against the CPU oracle on the same seeded /
analytic inputs, and against the reference's golden file.

Tolerances (BASELINE.json north_star: FP64 relative L2 <= 1e-10 per filtered field after the full run):
  * strict handles (reference order of operations, no FMA contraction): bit-exact with the oracle;
  * production handles (difference-form WENO, FMA): rel-L2 <= 1e-12 per tracer after the steps taken,
    <= 1e-10 on the filtered fields after a full forward+backward run.
"""
import numpy as np
import pytest

import lf_oracle as lo
import lfb200
from conftest import GOLDEN_GRID, GOLDEN_RUN
from lfb200 import _lib
from lfb200.synthetic import make_series
from parity_utils import compare_state, make_pair, oracle_names, rel_l2

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12
TOL_RUN = 1e-10


def golden_grid():
    return lfb200.RectilinearGrid(size=(10, 10), x=(-5000, 5000), z=(-100, 0), topology=("Periodic", "Flat", "Bounded"))


def golden_series(golden_sim):
    nt = len(golden_sim["t"])
    return list(golden_sim["t"]), {k: [golden_sim[k][n] for n in range(nt)] for k in ("u", "w", "b")}


@pytest.mark.parametrize("strict", [True, False])
def test_golden_case_steps_vs_oracle(golden_sim, strict):
    """80 aligned RK3 steps of the golden case, forward pass, every tracer compared with the oracle."""
    g = golden_grid()
    times, fields = golden_series(golden_sim)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=GOLDEN_RUN["freq_c"])
    o, m = make_pair(g, ("b",), ("u", "w"), fp, times, fields, strict=strict, kernel="generic")
    names = oracle_names(("b",), ("u", "w"), fp)
    o.init_from_data(); m.init_from_data()
    compare_state(o, m, names, 0, exact=True)          # initialisation is exact in both flavours
    steps = lfb200.aligned_time_steps(86400.0, 1200.0, [3600.0, 8640.0])
    assert len(steps) == 80
    for n, dt in enumerate(steps):
        o.step(dt); m.step(dt)
        if n in (0, 1, 9, 39, 79):
            compare_state(o, m, names, TOL_STEP, exact=strict)
    assert m.time == o.time == 86400.0
    # tendencies of the final state
    for q, name in enumerate(names):
        o.update_state()
        if strict:
            assert np.array_equal(o.tendency(q)[3:-3, :, 3:-3], m.tendency(name)[3:-3, :, 3:-3])
        else:
            assert rel_l2(m.tendency(name)[3:-3, :, 3:-3], o.tendency(q)[3:-3, :, 3:-3]) < 1e-10
    m.close(); o.close()


def test_golden_full_offline_run_vs_reference_file(golden_sim, golden_offline, tmp_path):
    """The reference's own integration test (test/test_offline_integration.jl:1-32) through the public API:
    run_offline_Lagrangian_filter on reference_sim, compared with reference_offline_output."""
    times, fields = golden_series(golden_sim)
    cfg = lfb200.OfflineFilterConfig(data=lfb200.InputData(times, fields), grid=golden_grid(),
                                     var_names_to_filter=("b",), velocity_names=("u", "w"), Δt=1200.0, T_out=3600.0,
                                     N=2, freq_c=1e-4 / 4, compute_mean_velocities=True, compute_Eulerian_filter=True,
                                     output_filename="")
    out = lfb200.run_offline_Lagrangian_filter(cfg, save=False)
    assert out.iterations == list(golden_offline["iterations"])
    # the file written for this run has the golden file's `timeseries` layout: same variables, same iteration keys,
    # same shapes and element types (run_offline_lagrangian_filter.jl:77-80, post:395-404, 751-755)
    f = lfb200.JLD2File(out.save(str(tmp_path / "filtered_output.jld2")))
    golden_names = [k for k in golden_offline.files if k not in ("t", "iterations")]
    assert sorted(f.keys("timeseries")) == sorted(golden_names + ["t"])
    for name in golden_names:
        assert f.iterations(name) == list(golden_offline["iterations"])
        a = f[f"timeseries/{name}/{out.iterations[5]}"]
        assert a.shape == golden_offline[name][5].shape and a.dtype == golden_offline[name].dtype, name
    # the reference test: isapprox(test, ref; rtol=1e-6) on the u family (array norms, halos included)
    for name in ("u", "u_Lagrangian_filtered", "u_Lagrangian_filtered_at_mean", "u_Eulerian_filtered"):
        got, ref = out.stacked(name), golden_offline[name]
        ok = ~np.isnan(ref)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), name
        assert np.linalg.norm(got[ok] - ref[ok]) <= 1e-6 * np.linalg.norm(ref[ok]), name
    # much tighter than the reference's own tolerance: every pinned field to 1e-10 (production kernel,
    # Float32-rounded pass outputs as the reference writes them)
    for name in ("b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered", "w_Lagrangian_filtered",
                 "b_Lagrangian_filtered_at_mean", "w_Lagrangian_filtered_at_mean"):
        got, ref = out.stacked(name), golden_offline[name]
        ok = ~np.isnan(ref)
        assert rel_l2(got[ok], ref[ok]) <= 1e-7, (name, rel_l2(got[ok], ref[ok]))
    # known-answer values, SURVEY.md App. A.10 (t = 43200 s, cells (1,6),(2,6),(3,6),(1,1))
    n = 12
    want = [-2.838078944478184e-05, -7.992282553459518e-05, -9.990392209147103e-05, -5.282615893520415e-05]
    got = [out["b_Lagrangian_filtered"][n][i + 2, 0, k + 2] for (i, k) in [(1, 6), (2, 6), (3, 6), (1, 1)]]
    np.testing.assert_allclose(got, want, rtol=2e-7)


def test_golden_full_run_fp64_vs_oracle(golden_sim):
    """north_star tolerance: <= 1e-10 relative L2 per filtered field after the full forward+backward run,
    in-memory FP64 (no Float32 rounding of the pass outputs), production kernel vs oracle."""
    times, fields = golden_series(golden_sim)
    g = golden_grid()
    cfg = lfb200.OfflineFilterConfig(data=lfb200.InputData(times, fields), grid=g, var_names_to_filter=("b",),
                                     velocity_names=("u", "w"), Δt=1200.0, T_out=3600.0, N=2, freq_c=1e-4 / 4,
                                     output_filename="", map_to_mean=True, output_original_data=False)
    cfg.map_to_mean = True
    out = lfb200.run_offline_Lagrangian_filter(cfg, output_dtype=np.float64, save=False)
    fp = lo.bw2_params(2, 1e-4 / 4)
    ref = lo.run_offline(times, {k: [a.astype(np.float64) for a in fields[k]] for k in ("u", "w")},
                         [[a.astype(np.float64) for a in fields["b"]]], ["b"], N=g.N, H=g.H, topology=g.topology,
                         delta=g.spacing, filter_params=fp, dt=1200.0, T_out=3600.0, velocities="uw",
                         float32_outputs=False)
    for name in ("b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered", "w_Lagrangian_filtered"):
        e = rel_l2(out.stacked(name), np.stack(ref[name]))
        assert e <= TOL_RUN, (name, e)


@pytest.mark.parametrize("mode", ["y-replicated", "x-to-y"])
@pytest.mark.parametrize("halo", [3, 4])
def test_extruded_golden_case_on_production_kernel(golden_sim, golden_offline, mode, halo):
    """The reference's golden vectors on the PRODUCTION kernel: the 10 x 1 x 10 golden case extruded into 3-D
    (tests/golden_extrude.py: replicated along a Periodic y axis, and the permuted twin with golden x mapped onto y)
    runs on lfb_stage_ztma through run_offline_Lagrangian_filter.  Every replicated column of all five pinned
    fields must equal reference_offline_output to the Float32 precision the file carries (<= 1e-7), and the
    in-memory FP64 run must match the oracle's 2-D golden run to the north-star tolerance (<= 1e-10).
    halo = 4 is the lee-wave example's halo (examples/offline_filter_lee_wave.jl:49)."""
    import golden_extrude as ge
    grid_kw, times, fields, vel_names, rename = ge.extrude(golden_sim, mode, n_rep=8, halo=halo)
    g = lfb200.RectilinearGrid(**grid_kw)

    def run(dtype):
        cfg = lfb200.OfflineFilterConfig(data=lfb200.InputData(times, fields), grid=g, var_names_to_filter=("b",),
                                         velocity_names=vel_names, Δt=1200.0, T_out=3600.0, N=2, freq_c=1e-4 / 4,
                                         output_filename="", n_slots=25, output_original_data=False, kernel="ztma")
        out = lfb200.run_offline_Lagrangian_filter(cfg, output_dtype=dtype, save=False, regrid=False)
        assert out.stats["kernel"] == "ztma"
        return out

    out32 = run(np.float32)
    assert out32.iterations == list(golden_offline["iterations"])
    names = ["b_Lagrangian_filtered", "xi_" + vel_names[0], "xi_w", vel_names[0] + "_Lagrangian_filtered",
             "w_Lagrangian_filtered"]
    for name in names:
        gname = rename.get(name, name)
        for n in range(25):
            ref = ge.golden_rehaloed(golden_offline[gname][n], gname, halo)
            for col in ge.columns_of(out32[name][n], mode, halo):
                assert rel_l2(col, ref) <= 1e-7, (name, n, rel_l2(col, ref))
    # FP64 in memory against the oracle's run of the ORIGINAL 2-D golden case
    out64 = run(np.float64)
    g2 = golden_grid()
    t2, f2 = golden_series(golden_sim)
    ref = lo.run_offline(t2, {k: [a.astype(np.float64) for a in f2[k]] for k in ("u", "w")},
                         [[a.astype(np.float64) for a in f2["b"]]], ["b"], N=g2.N, H=g2.H, topology=g2.topology,
                         delta=g2.spacing, filter_params=lo.bw2_params(2, 1e-4 / 4), dt=1200.0, T_out=3600.0,
                         velocities="uw", float32_outputs=False)
    for name in names:
        gname = rename.get(name, name)
        want = np.stack([ge.golden_rehaloed(a, gname, halo) for a in ref[gname]])
        ncol = len(ge.columns_of(out64[name][0], mode, halo))
        for c in range(ncol):
            got = np.stack([ge.columns_of(out64[name][n], mode, halo)[c] for n in range(25)])
            assert rel_l2(got, want) <= TOL_RUN, (name, c, rel_l2(got, want))


CASES_3D = [
    # (size, topology, vars, vels, N)
    ((24, 12, 10), ("Periodic", "Periodic", "Bounded"), ("T", "b"), ("u", "v", "w"), 2),
    ((16, 10, 9), ("Periodic", "Bounded", "Bounded"), ("T",), ("u", "v", "w"), 4),
    ((12, 14, 8), ("Bounded", "Periodic", "Periodic"), ("b",), ("u", "v", "w"), 2),
    ((20, 9), ("Periodic", "Bounded", "Flat"), ("T", "b"), ("u", "v"), 2),
    ((33, 7, 12), ("Periodic", "Periodic", "Bounded"), ("T", "b"), ("u", "w"), 1),
]


@pytest.mark.parametrize("size,topology,var_names,vel_names,N", CASES_3D)
@pytest.mark.parametrize("strict", [True, False])
def test_3d_and_mixed_topologies_vs_oracle(size, topology, var_names, vel_names, N, strict):
    """Shapes the reference's tests never pin (3-D, y advection, Bounded x/y, Periodic z, N=1, N=4,
    ragged sizes): CUDA vs oracle, bit-exact in strict mode."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, strict=strict, kernel="generic")
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2, 0.3):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP, exact=strict)
    # outputs incl. halos (create_output_fields + halo fill pattern of SURVEY.md App. A.6)
    for q in range(len(var_names)):
        ref, got = o.output(0, q), m.output(_lib.OUT_FILTERED_VAR, q)
        assert np.array_equal(ref, got) if strict else rel_l2(got, ref) < TOL_STEP
    for mi in range(len(vel_names)):
        for kind in (_lib.OUT_XI, _lib.OUT_MEAN_VEL):
            ref, got = o.output(kind, mi), m.output(kind, mi)
            assert np.array_equal(ref, got) if strict else rel_l2(got, ref) < TOL_STEP
    m.close(); o.close()


def test_backward_pass_reuses_resident_snapshots():
    """Backward pass = re-tagged slots + on-the-fly velocity negation (no re-upload) must equal the oracle's
    reversed / negated series (reference utils:154-173, 1232-1251)."""
    g = lfb200.RectilinearGrid(size=(16, 8, 6), extent=(16.0, 8.0, 6.0), topology=("Periodic", "Periodic", "Bounded"))
    times = [0.0, 1.0, 2.0, 3.0]
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=3.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, strict=True, kernel="generic")
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for _ in range(12):
        o.step(0.25); m.step(0.25)
    # switch direction
    o.load_series(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in vel_names},
                  [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in var_names], 0.0, 3.0, backward=True)
    o.negate_maps(); o.reset_clock()
    for s, t in enumerate(times):
        m.set_slot_time(s, 3.0 - t)
    m.set_direction(-1); m.negate_maps(); m.reset_clock()
    for _ in range(12):
        o.step(0.25); m.step(0.25)
    compare_state(o, m, names, 0, exact=True)
    m.close(); o.close()


def test_eulerian_and_relaxation_paths_vs_oracle():
    g = lfb200.RectilinearGrid(size=(12, 10), x=(0, 12), z=(-10, 0), topology=("Bounded", "Flat", "Bounded"))
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=("b",), velocity_names=("u", "w"), T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=3.0)
    # advection = nothing
    o, m = make_pair(g, ("b",), ("u", "w"), fp, times, fields, strict=True, kernel="generic", advect=False)
    names = oracle_names(("b",), ("u", "w"), fp)
    o.init_from_data(); m.init_from_data()
    for _ in range(5):
        o.step(0.2); m.step(0.2)
    compare_state(o, m, names, 0, exact=True)
    m.close(); o.close()
    # boundary relaxation with a mask
    x = g.nodes(0)[:, None, None]
    mask = np.asfortranarray(np.exp(-((x - 0.0) / 2.0) ** 2) + 0 * g.nodes(2)[None, None, :])
    for strict in (True, False):
        o, m = make_pair(g, ("b",), ("u", "w"), fp, times, fields, strict=strict, kernel="generic",
                         relax_timescale=0.7, mask=mask)
        o.init_from_data(); m.init_from_data()
        for _ in range(5):
            o.step(0.2); m.step(0.2)
        compare_state(o, m, names, TOL_STEP, exact=strict)
        m.close(); o.close()


def test_streaming_slots_match_all_resident(golden_sim):
    """n_slots = 2 (snapshots streamed through the slots, forward and reversed) gives the same result as
    keeping all 25 snapshots resident."""
    times, fields = golden_series(golden_sim)
    outs = []
    for n_slots in (25, 2):
        cfg = lfb200.OfflineFilterConfig(data=lfb200.InputData(times, fields), grid=golden_grid(),
                                         var_names_to_filter=("b",), velocity_names=("u", "w"), Δt=1200.0,
                                         T_out=3600.0, N=2, freq_c=1e-4 / 4, n_slots=n_slots, output_filename="",
                                         map_to_mean=False, output_original_data=False)
        outs.append(lfb200.run_offline_Lagrangian_filter(cfg, output_dtype=np.float64, save=False))
    for name in outs[0].keys():
        assert np.array_equal(outs[0].stacked(name), outs[1].stacked(name)), name


def test_error_paths():
    g = golden_grid()
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=1e-4)
    from lfb200.model import LagrangianFilter
    m = LagrangianFilter(g, var_names_to_filter=("b",), velocity_names=("u", "w"), filter_params=fp, n_slots=2)
    with pytest.raises(_lib.LfbError) as e:
        m.step(1.0)                       # nothing resident
    assert e.value.code == -4
    with pytest.raises(_lib.LfbError):
        m.step(-1.0)
    with pytest.raises(ValueError):
        m.upload_snapshot(0, "u", np.zeros((3, 3, 3)), 0.0)
    with pytest.raises(_lib.LfbError):
        m.upload_snapshot(7, "u", np.zeros(g.face_shape(0)), 0.0)
    m.close()


# ---------------------------------------------------------------------------------------------- z-march kernel
ZM_CASES = [
    # (size, vars, N, backward)
    ((32, 8, 8), ("T", "b"), 2, False),
    ((64, 16, 12), ("T", "b"), 2, True),
    ((32, 24, 9), ("b",), 1, False),
    ((96, 8, 16), ("T",), 4, False),
]


@pytest.mark.parametrize("kernel", ["ztma"])
@pytest.mark.parametrize("size,var_names,N,backward", ZM_CASES)
def test_zmarch_kernel_vs_oracle(size, var_names, N, backward, kernel):
    """The production z-marching stage kernel (faces computed once, register z window, helper warps)
    against the CPU oracle: every tracer after several aligned and unaligned RK3 steps."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=("Periodic", "Periodic", "Bounded"))
    vel_names = ("u", "v", "w")
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel=kernel, backward=backward)
    assert m.kernel_name == kernel
    names = oracle_names(var_names, vel_names, fp)
    if backward:
        # initialise from the forward data, then run the backward pass on re-tagged slots
        o.load_series(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in vel_names},
                      [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in var_names], 0.0, 1.0, backward=False)
        for s, t in enumerate(times):
            m.set_slot_time(s, t)
        m.set_direction(1)
        o.init_from_data(); m.init_from_data()
        o.load_series(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in vel_names},
                      [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in var_names], 0.0, 1.0, backward=True)
        for s, t in enumerate(times):
            m.set_slot_time(s, 1.0 - t)
        m.set_direction(-1)
        o.negate_maps(); m.negate_maps()
    else:
        o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2, 0.3):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    m.close(); o.close()


@pytest.mark.parametrize("size", [(120, 5, 40), (40, 7, 13), (72, 20, 9)])
def test_ztma_partial_tiles_in_x_and_y(size):
    """The TMA z-march kernel on grids whose x / y extents are not multiples of the 32 x 8 tile (the reference's 3-D
    example is 120 x 5 x 40): partial tiles compute like full ones and drop their surplus columns / rows; against
    the CPU oracle on the same inputs."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=("Periodic", "Periodic", "Bounded"))
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 4.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel="auto")
    assert m.kernel_name == "ztma"
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for _ in range(5):
        o.step(0.125); m.step(0.125)
    compare_state(o, m, names, TOL_STEP)
    m.close(); o.close()


@pytest.mark.parametrize("size,topology,N", [
    ((40, 12, 10), ("Periodic", "Bounded", "Bounded"), 2),        # channel: walls in y
    ((48, 44, 10), ("Periodic", "Bounded", "Bounded"), 2),        # ... tall enough for the split launch: wall strips on the
                                                                  # general instantiation, the rows between on the all-periodic one
    ((33, 9, 12), ("Bounded", "Bounded", "Bounded"), 2),          # box
    ((104, 40, 9), ("Bounded", "Bounded", "Bounded"), 2),         # ... wide and tall enough for the split launches in x and y
    ((133, 10, 12), ("Bounded", "Periodic", "Bounded"), 1),       # ... in x only, partial last x tile
    ((36, 20, 9), ("Bounded", "Periodic", "Bounded"), 1),         # walls in x, single exponential
    ((70, 17, 14), ("Periodic", "Bounded", "Bounded"), 4),        # several x tiles, partial tiles, two coefficient sets
    ((40, 12, 16), ("Periodic", "Periodic", "Periodic"), 2),      # Periodic z: WENO5 on every z face, z halo images
    ((33, 10, 9), ("Bounded", "Periodic", "Periodic"), 2),
    ((36, 9, 13), ("Periodic", "Bounded", "Periodic"), 1),
])
def test_ztma_bounded_x_y_vs_oracle(size, topology, N):
    """The TMA z-march kernel with Bounded x / y axes (buffer schemes WENO3 / Centered2 / wall next to the walls,
    SURVEY App. A.5) against the CPU oracle: every tracer after several RK3 steps."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel="ztma")
    assert m.kernel_name == "ztma"
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    m.close(); o.close()


@pytest.mark.parametrize("size,topology,N", [
    ((40, 12, 10), ("Periodic", "Bounded", "Bounded"), 2),
    ((33, 10, 9), ("Bounded", "Bounded", "Bounded"), 1),
    ((64, 16, 12), ("Periodic", "Periodic", "Bounded"), 2),
])
def test_ztma_boundary_relaxation_vs_oracle(size, topology, N):
    """boundary_relaxation (reference utils:584-699: -(1/tau_r) (g - g_eq) mask) on the TMA z-march kernel: filtered
    variables and maps, cosine and sine tracers, single exponential, against the CPU oracle."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    y = g.nodes(1)[None, :, None]
    z = g.nodes(2)[None, None, :]
    mask = np.asfortranarray(np.exp(-((y - g.origin[1]) / 2.0) ** 2) + np.exp(-((y - g.origin[1] - g.L[1]) / 2.0) ** 2)
                             + 0.1 * np.cos(z) ** 2 + 0 * g.nodes(0)[:, None, None])
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel="ztma", relax_timescale=0.7, mask=mask)
    assert m.kernel_name == "ztma"
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    m.close(); o.close()


@pytest.mark.parametrize("vel_names", [("u", "w"), ("v", "w"), ("u", "v")])
def test_ztma_with_absent_velocity_component(vel_names):
    """A 3-D grid filtered with two of the three velocity components (the third is not advected with and has no
    map): the TMA kernel substitutes a zero field; against the CPU oracle."""
    g = lfb200.RectilinearGrid(size=(40, 12, 10), extent=(40.0, 12.0, 10.0), topology=("Periodic", "Periodic", "Bounded"))
    var_names = ("T", "b")
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel="auto")
    assert m.kernel_name == "ztma"
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    m.close(); o.close()


@pytest.mark.parametrize("size,topology,var_names,vel_names,N,backward,relax", [
    ((10, 10), ("Periodic", "Flat", "Bounded"), ("b",), ("u", "w"), 2, False, False),           # the golden case's shape: one partial row
    ((400, 80), ("Periodic", "Flat", "Bounded"), ("T", "b"), ("u", "w", "v"), 2, False, False),   # BASELINE config 1 (v forces xi_v only)
    ((100, 100), ("Periodic", "Flat", "Bounded"), ("T", "b"), ("u", "w"), 4, True, False),         # lee-wave shape, backward pass
    ((70, 24), ("Periodic", "Flat", "Bounded"), ("b",), ("u", "w"), 1, True, False),               # partial last row, single exponential
    ((45, 16), ("Periodic", "Flat", "Bounded"), ("b",), ("w", "v"), 2, False, False),              # odd Nx (no side-job prefetch), u absent
    ((96, 20), ("Bounded", "Flat", "Bounded"), ("T",), ("u", "w"), 2, False, False),               # walls in x
    ((64, 16), ("Periodic", "Flat", "Periodic"), ("b",), ("u", "w", "v"), 2, True, False),         # Periodic z, v on the Flat axis, backward
    ((80, 24), ("Bounded", "Flat", "Bounded"), ("b",), ("u", "w", "v"), 2, False, True),           # boundary relaxation
    ((2080, 12), ("Periodic", "Flat", "Bounded"), ("b",), ("u", "w"), 2, False, False),             # long enough for kernel = "auto"
])
def test_ztma_folded_2d_grids_vs_oracle(size, topology, var_names, vel_names, N, backward, relax):
    """Grids with a Flat y axis (every 2-D example of the reference) on the TMA z-march kernel: the x line is folded
    into rows of 32 cells read through overlapping tensor-map rows (profiles/microbench/tma_overlap.cu), y faces are
    skipped, a v component only forces its xi_v map.  Against the CPU oracle on the same inputs."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    times = [0.0, 0.5, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    kw = {}
    if size[0] >= 2048:      # what "auto" picks
        m0 = lfb200.LagrangianFilter(g, var_names_to_filter=var_names, velocity_names=vel_names,
                                     filter_params=lfb200.set_offline_BW2_filter_params(N=N, freq_c=1.0))
        assert m0.kernel_name == "ztma"
        m0.close()
    if relax:
        x, z = g.nodes(0)[:, None, None], g.nodes(2)[None, None, :]
        kw = dict(relax_timescale=0.7, mask=np.asfortranarray(np.exp(-((x - g.origin[0]) / 6.0) ** 2) + 0.1 * np.cos(z) ** 2))
    # kernel = "auto" keeps lines shorter than 2048 cells on the generic kernel (faster at those sizes)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields, kernel="ztma", backward=backward, **kw)
    assert m.kernel_name == "ztma"
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    for dt in (0.1, 0.1, 0.05, 0.25, 0.2, 0.3):
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    # tendencies of the final state (generic kernel on the same handle) agree too
    o.update_state()
    assert rel_l2(m.tendency(names[0])[3:-3, :, 3:-3], o.tendency(0)[3:-3, :, 3:-3]) < 1e-10
    m.close(); o.close()


def test_zmarch_matches_generic_on_larger_grid():
    g = lfb200.RectilinearGrid(size=(128, 64, 40), extent=(128.0, 64.0, 40.0), topology=("Periodic", "Periodic", "Bounded"))
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    times = [0.0, 1.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 4.0)
    from lfb200.model import LagrangianFilter
    res = {}
    for kernel in ("generic", "ztma"):
        m = LagrangianFilter(g, var_names_to_filter=var_names, velocity_names=vel_names, filter_params=fp, n_slots=2,
                             kernel=kernel)
        for s, t in enumerate(times):
            for name in vel_names + var_names:
                m.upload_snapshot(s, name, fields[name][s], t)
        m.init_from_data()
        for _ in range(6):
            m.step(0.125)
        res[kernel] = [m.tracer(n) for n in m.tracer_names]
        m.close()
    for other in ("ztma",):
        for a, b in zip(res["generic"], res[other]):
            assert rel_l2(b, a) < TOL_STEP


def test_multi_gpu_slabs_match_single_gpu():
    """y-slab decomposition with NCCL halo exchange (tests/mgpu_check.py under torchrun) when the box has >= 2 GPUs."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29517", "tests/mgpu_check.py"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


def test_multi_gpu_pipeline_matches_single_gpu():
    """The whole offline pipeline through the public API on 2 ranks (tests/mgpu_pipeline_check.py under torchrun):
    slab passes, device-side combination, gather by output time, remap sharded by output time."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29518", "tests/mgpu_pipeline_check.py"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MISMATCH" not in r.stdout


@pytest.mark.parametrize("size,topology", [((10, 10), ("Periodic", "Flat", "Bounded")),
                                           ((24, 12, 10), ("Periodic", "Periodic", "Bounded"))])
def test_device_and_host_output_stores_agree(size, topology):
    """run_offline_Lagrangian_filter with the pass outputs kept on the device (combined there, handed to the remap as
    device pointers) and with the bounded page-locked host ring: the same bits, `_at_mean` fields included."""
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    vel_names = ("u", "w") if topology[1] == "Flat" else ("u", "v", "w")
    times = [0.0, 0.5, 1.0, 1.5, 2.0]
    fields = make_series(g, times, var_names=("b",), velocity_names=vel_names, T_period=4.0, amplitude=0.5)
    data = lfb200.InputData(times, {k: [np.asarray(a, dtype=np.float32) for a in v] for k, v in fields.items()})
    outs = {}
    for store in ("device", "host"):
        cfg = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=vel_names, N=2,
                                         freq_c=2 * np.pi / 2.0, dt=0.1, T_out=0.5, n_slots=5, npad=3)
        outs[store] = lfb200.run_offline_Lagrangian_filter(cfg, save=False, output_store=store)
        assert outs[store].stats["output_store"] == store
    a, b = outs["device"], outs["host"]
    assert a.times == b.times and a.iterations == b.iterations and set(a.keys()) == set(b.keys())
    assert any(k.endswith("_at_mean") for k in a.keys())
    for name in a.keys():
        for x, y in zip(a[name], b[name]):
            assert x.dtype == y.dtype and np.array_equal(x, y, equal_nan=True), name


def test_t_start_must_sit_on_a_snapshot(golden_sim):
    g = golden_grid()
    times, fields = golden_series(golden_sim)
    data = lfb200.InputData(times, fields)
    cfg = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                     freq_c=1e-4, T_start=times[0] + 1.0, T_end=times[-1])
    with pytest.raises(ValueError, match="coincide with snapshot times"):
        lfb200.run_offline_Lagrangian_filter(cfg, save=False)


# ---------------------------------------------------------------------------------------------- BASELINE configs
EXAMPLE_SHAPES = [
    # BASELINE.json configs[0]: examples/offline_filter_geostrophic_adjustment.jl (400x1x80, P/F/B, u,w,v, N=2)
    ("geostrophic_adjustment_2d", (400, 80), ("Periodic", "Flat", "Bounded"), ("T", "b"), ("u", "w", "v"), 2),
    # configs[2]: examples/offline_filter_geostrophic_adjustment_3D.jl (120x5x40, P/P/B)
    ("geostrophic_adjustment_3d", (120, 5, 40), ("Periodic", "Periodic", "Bounded"), ("T", "b"), ("u", "v", "w"), 2),
    # configs[3]: examples/offline_filter_lee_wave.jl shape without the immersed bottom (100x1x100, N=4)
    ("lee_wave_flat_bottom", (100, 100), ("Periodic", "Flat", "Bounded"), ("T", "b"), ("u", "w"), 4),
]


@pytest.mark.parametrize("label,size,topology,var_names,vel_names,N", EXAMPLE_SHAPES)
def test_example_config_shapes_vs_oracle(label, size, topology, var_names, vel_names, N):
    """The grid shapes / tracer sets of the reference's example scripts (BASELINE.json configs), forward and
    backward pass with aligned steps, production kernels vs oracle on in-memory FP64 state."""
    ext = tuple(float(s) for s in size)
    g = lfb200.RectilinearGrid(size=size, extent=ext, topology=topology)
    times = [0.0, 1.0, 2.0]
    fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
    fp = lfb200.set_offline_BW2_filter_params(N=N, freq_c=2 * np.pi / 2.0)
    o, m = make_pair(g, var_names, vel_names, fp, times, fields)
    names = oracle_names(var_names, vel_names, fp)
    o.init_from_data(); m.init_from_data()
    steps = lfb200.aligned_time_steps(2.0, 0.3, [1.0, 0.2])
    for dt in steps:
        o.step(dt); m.step(dt)
    compare_state(o, m, names, TOL_STEP)
    o.load_series(times, {k: [np.asarray(a, dtype=np.float64) for a in fields[k]] for k in vel_names},
                  [[np.asarray(a, dtype=np.float64) for a in fields[v]] for v in var_names], 0.0, 2.0, backward=True)
    o.negate_maps(); o.reset_clock()
    for s, t in enumerate(times):
        m.set_slot_time(s, 2.0 - t)
    m.set_direction(-1); m.negate_maps(); m.reset_clock()
    for dt in steps:
        o.step(dt); m.step(dt)
    worst = compare_state(o, m, names, 1e-11)
    m.close(); o.close()


def test_full_size_properties():
    """BASELINE.json's metric configuration (1024x1024x128, 10 tracers) is far too large for the oracle, so parity
    at full size goes through size-independent properties of the path:
      * translation equivariance: shifting every input by (sx, sy) cells along the periodic axes shifts every
        tracer by the same amount, bit for bit (the z-march kernel's tiling, halo images and helper faces must
        not depend on where a cell sits in a tile);
      * a uniform variable stays uniform under advection by a divergent flow (-div(u c) + c div(u) = 0), so its
        filter tracers remain uniform in space to rounding;
      * the production kernel agrees with the generic kernel on the same inputs (<= 1e-12)."""
    import torch
    from lfb200.model import LagrangianFilter
    from lfb200.synthetic import make_snapshot
    size = (1024, 1024, 128)
    free, _ = torch.cuda.mem_get_info()
    if free < 120e9:
        pytest.skip("needs ~110 GB of free device memory")
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=("Periodic", "Periodic", "Bounded"))
    var_names, vel_names = ("T", "b"), ("u", "v", "w")
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 4.0)
    dev = torch.device("cuda", 0)
    snaps = []
    for t in (0.0, 1.0):
        s = make_snapshot(g, t, 4.0, var_names=var_names, velocity_names=vel_names, xp=torch, device=dev, dtype=np.float32)
        s["T"] = torch.full_like(s["T"], 2.5)                       # uniform variable
        snaps.append({k: v.cpu().numpy().T for k, v in s.items()})  # Fortran (ex, ey, ez) views
        del s
    torch.cuda.empty_cache()
    H = 3

    def shifted(a, sx, sy):
        """Periodic shift of a parent array by (sx, sy) cells, halos rebuilt."""
        nx, ny = size[0], size[1]
        core = a[H:H + nx, H:H + ny, :]
        core = np.roll(core, (sx, sy), axis=(0, 1))
        out = np.empty_like(a)
        out[H:H + nx, H:H + ny, :] = core
        out[:H, H:H + ny, :] = core[nx - H:, :, :]; out[H + nx:, H:H + ny, :] = core[:H, :, :]
        out[:, :H, :] = out[:, ny:ny + H, :]; out[:, H + ny:, :] = out[:, H:2 * H, :]
        return np.asfortranarray(out)

    def run(kernel, shift=None, n_steps=2, want=("T_C1", "b_S1", "xi_v_C1")):
        m = LagrangianFilter(g, var_names_to_filter=var_names, velocity_names=vel_names, filter_params=fp, n_slots=2,
                             kernel=kernel)
        for sl, t in enumerate((0.0, 1.0)):
            for name in vel_names + var_names:
                a = snaps[sl][name]
                m.upload_snapshot(sl, name, shifted(a, *shift) if shift else a, t)
        m.init_from_data()
        for _ in range(n_steps):
            m.step(0.125)
        out = {n: m.tracer(n) for n in want}
        m.close()
        return out

    base = run("auto")
    # uniform variable: tracers uniform over the interior
    tc = base["T_C1"][H:-H, H:-H, H:-H]
    assert np.ptp(tc) <= 1e-13 * abs(tc.mean()), np.ptp(tc)
    # production vs generic kernel
    gen = run("generic", want=("b_S1", "xi_v_C1"))
    for n in gen:
        assert rel_l2(base[n], gen[n]) <= TOL_STEP, n
    del gen
    # translation equivariance (37 and 5 are not multiples of the 32 x 6 tile)
    sh = run("auto", shift=(37, 5), want=("b_S1", "xi_v_C1"))
    for n in sh:
        a = base[n][H:-H, H:-H, :]
        b = sh[n][H:-H, H:-H, :]
        assert np.array_equal(np.roll(a, (37, 5), axis=(0, 1)), b), n


def test_run_until_matches_host_alignment(golden_sim):
    """lfb_run_until (Simulation time-step alignment on the C side) takes the same steps
    as the host's aligned_time_steps: same clock, iteration count and state."""
    g = golden_grid()
    times, fields = golden_series(golden_sim)
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=GOLDEN_RUN["freq_c"])
    from lfb200.model import LagrangianFilter
    states = []
    for mode in ("host", "c"):
        m = LagrangianFilter(g, var_names_to_filter=("b",), velocity_names=("u", "w"), filter_params=fp, n_slots=25,
                             strict=True)
        for s, t in enumerate(times):
            for name in ("u", "w", "b"):
                m.upload_snapshot(s, name, fields[name][s], t)
        m.init_from_data()
        if mode == "host":
            for dt in lfb200.aligned_time_steps(86400.0, 1200.0, [3600.0, 8640.0]):
                m.step(dt)
        else:
            m.run_until(86400.0, 1200.0, [3600.0, 8640.0])
        assert m.time == 86400.0 and m.iteration == 80
        states.append([m.tracer(n) for n in m.tracer_names])
        m.close()
    for a, b in zip(*states):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------- remap to mean position
def test_device_remap_matches_golden(golden_offline):
    """regrid_to_mean_position on the GPU (local Delaunay walk) against the reference's pinned `*_at_mean`
    fields (scipy / Qhull inside the reference), all 25 output times."""
    from lfb200.remap import remap_to_mean_device
    names = ["b_Lagrangian_filtered", "u_Lagrangian_filtered", "w_Lagrangian_filtered"]
    for n in range(25):
        fields = {k: golden_offline[k][n] for k in names}
        xi = {"u": golden_offline["xi_u"][n], "w": golden_offline["xi_w"][n]}
        out = remap_to_mean_device(fields, xi, GOLDEN_GRID["N"], GOLDEN_GRID["H"], GOLDEN_GRID["topology"],
                                   GOLDEN_GRID["origin"], GOLDEN_GRID["delta"], npad=5)
        for k in names:
            ref, got = golden_offline[k + "_at_mean"][n], out[k + "_at_mean"]
            assert np.array_equal(np.isnan(got), np.isnan(ref)), (n, k)
            assert np.nanmax(np.abs(got - ref)) <= 1e-12 * np.nanmax(np.abs(ref)), (n, k)


@pytest.mark.parametrize("size,topology,amp", [
    ((48, 36), ("Periodic", "Flat", "Bounded"), 0.6),
    ((40, 30), ("Periodic", "Flat", "Bounded"), 1.7),      # displacements beyond one cell: larger patches
    ((32, 24), ("Periodic", "Periodic", "Flat"), 0.8),
    ((30, 20), ("Bounded", "Flat", "Bounded"), 0.5),
    ((200, 16), ("Periodic", "Flat", "Bounded"), 0.7),     # strongly anisotropic normalised cells
])
def test_device_remap_matches_scipy_oracle(size, topology, amp):
    import remap_oracle as ro
    from lfb200.remap import remap_to_mean_device
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) * (3.0 if n == 0 else 1.0) for n, s in enumerate(size)),
                               topology=topology)
    axes = g.nonflat_axes
    rng = np.random.default_rng(11)
    shape = g.center_shape()
    coords = [g.nodes(ax) for ax in range(3)]
    mesh = np.meshgrid(*[coords[ax] if ax in axes else np.zeros(1) for ax in range(3)], indexing="ij")

    def smooth(phase):
        out = 0
        for ax in axes:
            out = out + np.sin(2 * np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax] * (1 + ax) + phase + ax)
        return out / len(axes)

    xi = {}
    for ax in axes:
        f = smooth(0.3 * ax) * amp * g.spacing[ax]
        if topology[ax] == "Bounded":          # no normal displacement at the walls
            f = f * np.sin(np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax])
        xi["uvw"[ax]] = np.asfortranarray(f + 0.05 * g.spacing[ax] * rng.standard_normal(shape))
    fields = {"T": np.asfortranarray(smooth(1.0) + 0.1 * rng.standard_normal(shape)),
              "b": np.asfortranarray(np.tanh(3 * smooth(2.0)))}
    want = ro.remap_to_mean(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=5)
    got = remap_to_mean_device(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=5)
    for k in want:
        assert np.array_equal(np.isnan(got[k]), np.isnan(want[k])), k
        assert np.nanmax(np.abs(got[k] - want[k])) <= 1e-10 * np.nanmax(np.abs(want[k])), k


@pytest.mark.parametrize("size,topology,amp", [
    ((12, 10, 8), ("Periodic", "Periodic", "Bounded"), 0.5),
    ((10, 8, 9), ("Periodic", "Periodic", "Bounded"), 1.3),       # displacements beyond one cell
    ((9, 8, 7), ("Periodic", "Periodic", "Periodic"), 0.6),
    ((10, 9, 8), ("Periodic", "Bounded", "Bounded"), 0.5),
    ((24, 6, 6), ("Periodic", "Periodic", "Bounded"), 0.6),       # anisotropic normalised cells
])
def test_device_remap_3d_matches_scipy_oracle(size, topology, amp):
    """Three non-Flat axes: local Delaunay tetrahedron walk + barycentric interpolation + the edge-plane
    interpolation on Bounded axes, against the scipy / Qhull oracle (the reference's own calls, post:630-745).
    (Topologies with a Bounded axis in front of a Periodic one are left out: the reference pads the wrong
    coordinate column there, post:527-611.)"""
    import remap_oracle as ro
    from lfb200.remap import remap_to_mean_device
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) * (2.0 if n == 0 else 1.0) for n, s in enumerate(size)),
                               topology=topology)
    rng = np.random.default_rng(5)
    shape = g.center_shape()
    coords = [g.nodes(ax) for ax in range(3)]
    mesh = np.meshgrid(*coords, indexing="ij")

    def smooth(phase):
        out = 0
        for ax in range(3):
            out = out + np.sin(2 * np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax] * (1 + ax % 2) + phase + ax)
        return out / 3

    xi = {}
    for ax in range(3):
        f = smooth(0.3 * ax) * amp * g.spacing[ax]
        if topology[ax] == "Bounded":          # no normal displacement at the walls
            f = f * np.sin(np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax])
        xi["uvw"[ax]] = np.asfortranarray(f + 0.05 * g.spacing[ax] * rng.standard_normal(shape))
    fields = {"T": np.asfortranarray(smooth(1.0) + 0.1 * rng.standard_normal(shape)),
              "b": np.asfortranarray(np.tanh(3 * smooth(2.0)))}
    want = ro.remap_to_mean(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=5)
    got = remap_to_mean_device(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=5)
    for k in want:
        assert np.array_equal(np.isnan(got[k]), np.isnan(want[k])), k
        assert np.nanmax(np.abs(got[k] - want[k])) <= 1e-10 * np.nanmax(np.abs(want[k])), k


@pytest.mark.parametrize("size,topology", [((30, 24), ("Bounded", "Periodic", "Flat")),
                                           ((10, 9, 8), ("Bounded", "Periodic", "Bounded")),
                                           ((9, 8, 10), ("Bounded", "Bounded", "Periodic"))])
def test_device_remap_bounded_axis_before_periodic_axis(size, topology):
    """Grids on which a Bounded axis precedes a Periodic one.  The reference pads the wrong coordinate column there
    (post:523-611: `column_number` only advances inside the periodic branches); liblfb200 pads each Periodic axis's own
    coordinate, as the reference's comments and documentation describe (stated in include/lfb200.h).  Checked against
    the scipy oracle in that documented mode, and shown to differ from the literal mode."""
    import remap_oracle as ro
    from lfb200.remap import remap_to_mean_device
    g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=topology)
    axes = g.nonflat_axes
    rng = np.random.default_rng(17)
    shape = g.center_shape()
    coords = [g.nodes(ax) for ax in range(3)]
    mesh = np.meshgrid(*[coords[ax] if ax in axes else np.zeros(1) for ax in range(3)], indexing="ij")
    smooth = lambda ph: sum(np.sin(2 * np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax] + ph + ax) for ax in axes) / len(axes)
    xi = {}
    for ax in axes:
        f = smooth(0.3 * ax) * 0.6 * g.spacing[ax]
        if topology[ax] == "Bounded":
            f = f * np.sin(np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax])
        xi["uvw"[ax]] = np.asfortranarray(f + 0.05 * g.spacing[ax] * rng.standard_normal(shape))
    fields = {"T": np.asfortranarray(smooth(1.0) + 0.1 * rng.standard_normal(shape))}
    want = ro.remap_to_mean(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=3, pad_columns="documented")
    literal = ro.remap_to_mean(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=3, pad_columns="reference")
    got = remap_to_mean_device(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=3)
    k = "T_at_mean"
    # Not compared: the lines where the edge planes of two Bounded axes meet.  There a node sits on the hull of the 2-D
    # cloud of an edge plane, where Qhull's and the local walk's handling of the nearly degenerate boundary slivers can
    # differ (1 node of 720 on these grids); the reference itself has no such grid in its examples or tests.
    on_edge = [np.zeros(g.center_shape(), dtype=bool) for _ in range(3)]
    for ax in axes:
        if topology[ax] == "Bounded":
            sl = [slice(None)] * 3
            for e in (g.H[ax], g.H[ax] + g.N[ax] - 1):
                sl[ax] = e
                on_edge[ax][tuple(sl)] = True
    skip = (on_edge[0].astype(int) + on_edge[1] + on_edge[2]) >= 2
    a, b = np.where(skip, 0.0, got[k]), np.where(skip, 0.0, want[k])
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) <= 1e-10 * np.nanmax(np.abs(b))
    differs = not np.array_equal(np.isnan(literal[k]), np.isnan(want[k])) or np.nanmax(np.abs(np.nan_to_num(literal[k] - want[k]))) > 1e-8
    assert differs, "the literal padding rule was expected to differ on this topology"


def test_device_remap_3d_reproduces_linear_function_at_size():
    """Size-independent property of the remap: piecewise-linear interpolation on ANY triangulation reproduces a
    linear function of the displaced coordinate.  With f = alpha (z + xi_z) + beta carried by the data points the
    remapped field must equal alpha z + beta at every node that is not on the first / last plane of the Bounded
    axis (those planes are overwritten by the reference's 2-D edge interpolation, which ignores xi_z)
    (96 x 64 x 40, displacements up to 0.8 cells)."""
    from lfb200.remap import remap_to_mean_device
    size, topology = (96, 64, 40), ("Periodic", "Periodic", "Bounded")
    g = lfb200.RectilinearGrid(size=size, extent=(192.0, 64.0, 20.0), topology=topology)
    rng = np.random.default_rng(9)
    shape = g.center_shape()
    mesh = np.meshgrid(*[g.nodes(ax) for ax in range(3)], indexing="ij")
    xi = {}
    for ax in range(3):
        f = 0.7 * g.spacing[ax] * np.sin(2 * np.pi * mesh[(ax + 1) % 3] / g.L[(ax + 1) % 3] * 3 + ax) * np.cos(2 * np.pi * mesh[ax] / g.L[ax] * 2)
        f = f + 0.05 * g.spacing[ax] * rng.standard_normal(shape)
        if topology[ax] == "Bounded":
            f = f * np.sin(np.pi * (mesh[ax] - g.origin[ax]) / g.L[ax])
        xi["uvw"[ax]] = np.asfortranarray(f)
    alpha, beta = 0.37, -1.25
    fields = {"lin": np.asfortranarray(alpha * (mesh[2] + xi["w"]) + beta)}
    got = remap_to_mean_device(fields, xi, g.N, g.H, g.topology, g.origin, g.spacing, npad=5)["lin_at_mean"]
    H = g.H
    inner = tuple(slice(H[ax], H[ax] + size[ax]) for ax in range(3))
    want = (alpha * mesh[2] + beta)[inner]
    assert not np.isnan(got[inner]).any()
    assert np.abs(got[inner] - want)[:, :, 1:-1].max() <= 1e-12 * np.abs(want).max()
    halo = got.copy(); halo[inner] = 0.0
    assert not halo.any()
