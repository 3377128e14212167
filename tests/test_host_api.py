"""Host-side logic (no GPU): the mirrors of the reference's config / naming / forcing helpers, modelled on
reference test/test_loading.jl, test/test_config_validation.jl and test/test_initialisation.jl, plus the
C-ABI library's load/export check."""
import ctypes
import logging
import os
import re

import numpy as np
import pytest

import lfb200
from conftest import ROOT


def test_exports_exist():
    # reference test/test_loading.jl:1-12
    for name in ("OfflineFilterConfig", "OnlineFilterConfig", "run_offline_Lagrangian_filter", "create_filtered_vars",
                 "create_forcing", "set_offline_BW2_filter_params", "set_online_BW_filter_params",
                 "sum_forward_backward_contributions", "regrid_to_mean_position", "compute_Eulerian_filter",
                 "compute_time_shift", "get_weight_function"):
        assert hasattr(lfb200, name), name


def test_filter_params_known_answers():
    p = lfb200.set_offline_BW2_filter_params(N=2, freq_c=1e-4 / 2)      # OfflineLagrangianFilter.jl:212
    assert (p.a1, p.b1, p.c1, p.d1, p.N_coeffs) == (1.767766952966369e-5, 1.767766952966369e-5,
                                                    3.535533905932738e-5, 3.535533905932738e-5, 1)
    p = lfb200.set_online_BW_filter_params(N=2, freq_c=5e-5)           # OnlineLagrangianFilter.jl:111
    np.testing.assert_allclose([p.a1, p.b1, p.c1, p.d1],
                               [1.421067568548072e-20, -7.071067811865475e-5, 3.535533905932738e-5,
                                -3.535533905932738e-5], rtol=1e-14, atol=1e-19)
    p = lfb200.set_offline_BW2_filter_params(N=1, freq_c=3.0)
    assert dict(p) == {"a1": 1.5, "c1": 3.0, "N_coeffs": 0.5}
    with pytest.raises(ValueError):
        lfb200.set_offline_BW2_filter_params(N=3, freq_c=1.0)


def small_grid():
    return lfb200.RectilinearGrid(size=(4, 4), x=(-1, 1), z=(-1, 0), topology=("Periodic", "Flat", "Bounded"))


def test_create_filtered_vars_order():
    # utils:423-470 -- all _C first (vars x i, then xi_<vel> x i), then all _S
    cfg = lfb200.OnlineFilterConfig(grid=small_grid(), var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                    freq_c=1e-4)
    assert lfb200.create_filtered_vars(cfg) == ("b_C1", "xi_u_C1", "xi_w_C1", "b_S1", "xi_u_S1", "xi_w_S1")
    cfg = lfb200.OnlineFilterConfig(grid=small_grid(), var_names_to_filter=("T", "b"), velocity_names=("u", "w"), N=4,
                                    freq_c=1e-4, label="_x")
    assert lfb200.create_filtered_vars(cfg)[:4] == ("T_x_C1", "T_x_C2", "b_x_C1", "b_x_C2")
    cfg = lfb200.OnlineFilterConfig(grid=small_grid(), var_names_to_filter=("b",), velocity_names=("u", "w"), N=1,
                                    freq_c=1e-4)
    assert lfb200.create_filtered_vars(cfg) == ("b_C1", "xi_u_C1", "xi_w_C1")
    cfg = lfb200.OnlineFilterConfig(grid=small_grid(), var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                    freq_c=1e-4, map_to_mean=False, compute_mean_velocities=False)
    assert lfb200.create_filtered_vars(cfg) == ("b_C1", "b_S1")


def test_create_forcing_relaxation_lengths():
    # reference test/test_initialisation.jl:119-146
    mask = lambda x, z, p: 1.0
    kw = dict(grid=small_grid(), var_names_to_filter=("b",), velocity_names=("u", "w"), N=1, freq_c=1e-4)
    c0 = lfb200.OnlineFilterConfig(**kw)
    c1 = lfb200.OnlineFilterConfig(**kw, boundary_relaxation=True, relax_timescale=3600.0, mask_func=mask,
                                   mask_params={"a": 1})
    fv = lfb200.create_filtered_vars(c0)
    f0, f1 = lfb200.create_forcing(fv, c0), lfb200.create_forcing(fv, c1)
    assert len(f0["b_C1"]) == 2 and len(f0["xi_u_C1"]) == 2
    assert len(f1["b_C1"]) == 3 and len(f1["xi_u_C1"]) == 3
    assert f1["b_C1"][:2] == f0["b_C1"]


def test_config_relaxation_validation(caplog):
    # reference test/test_config_validation.jl:1-87
    good = lambda x, z, p: 1.0
    bad = lambda x, p: 1.0
    for Config, extra in ((lfb200.OnlineFilterConfig, dict(grid=small_grid())),
                          (lfb200.OfflineFilterConfig,
                           dict(grid=small_grid(),
                                data=lfb200.InputData([0.0, 1.0], {k: [np.zeros(1)] * 2 for k in ("b", "u", "w")})))):
        kw = dict(var_names_to_filter=("b",), velocity_names=("u", "w"), N=1, freq_c=1e-4, **extra)
        with pytest.raises(ValueError):
            Config(**kw, boundary_relaxation=True, mask_func=good, mask_params={"a": 1})
        with pytest.raises(ValueError):
            Config(**kw, boundary_relaxation=True, relax_timescale=3600.0, mask_params={"a": 1})
        with pytest.raises(ValueError):
            Config(**kw, boundary_relaxation=True, relax_timescale=3600.0, mask_func=bad, mask_params={"a": 1})
        caplog.clear()
        with caplog.at_level(logging.WARNING, logger="lfb200"):
            c = Config(**kw, boundary_relaxation=True, relax_timescale=3600.0, mask_func=good, mask_params=None)
        assert any(r.levelno == logging.WARNING for r in caplog.records)
        assert c.boundary_relaxation and c.relax_timescale == 3600.0
        c2 = Config(**kw, boundary_relaxation=True, relax_timescale=3600.0, mask_func=good, mask_params={"a": 1})
        assert c2.mask_func is good and c2.mask_params == {"a": 1}


def test_offline_config_time_defaults_and_errors(golden_sim):
    nt = len(golden_sim["t"])
    data = lfb200.InputData(golden_sim["t"], {k: [golden_sim[k][n] for n in range(nt)] for k in ("u", "w", "b")})
    g = lfb200.RectilinearGrid(size=(10, 10), x=(-5000, 5000), z=(-100, 0), topology=("Periodic", "Flat", "Bounded"))
    c = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                   freq_c=1e-4 / 2)
    assert (c.T_start, c.T_end, c.T, c.T_out, c.Δt) == (0.0, 86400.0, 86400.0, 3600.0, 360.0)
    assert c.filter_params.c1 == 3.535533905932738e-5
    with pytest.raises(ValueError):
        lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("u",), velocity_names=("u", "w"), N=2, freq_c=1.0)
    with pytest.raises(ValueError):
        lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("q",), velocity_names=("u", "w"), N=2, freq_c=1.0)
    with pytest.raises(ValueError):
        lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"))
    with pytest.raises(ValueError):
        lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                   freq_c=1.0, T_start=0.0, T=10.0, T_end=30.0)
    with pytest.raises(ValueError):
        lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                   freq_c=1.0, T_end=1e9)
    with pytest.raises(FileNotFoundError):
        lfb200.OfflineFilterConfig(original_data_filename="/nonexistent.jld2", grid=g, var_names_to_filter=("b",),
                                   velocity_names=("u", "w"), N=2, freq_c=1.0)
    # advection = nothing switches map_to_mean off (OfflineLagrangianFilter.jl:449-451)
    c = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"), N=2,
                                   freq_c=1e-4, advection=None)
    assert c.map_to_mean is False
    # user-supplied coefficients: N_coeffs inferred (OfflineLagrangianFilter.jl:387-403)
    c = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"),
                                   filter_params=dict(a1=0.5, c1=1.0))
    assert c.filter_params.N_coeffs == 0.5


def test_aligned_time_steps():
    steps = lfb200.aligned_time_steps(86400.0, 1200.0, [3600.0, 8640.0])
    assert len(steps) == 80 and abs(sum(steps) - 86400.0) < 1e-6
    assert steps[7] == 240.0        # 8400 -> 8640: the progress callback's T/10 clips the step (App. A.2)


def test_grid():
    g = lfb200.RectilinearGrid(size=(10, 10), x=(-5000, 5000), z=(-100, 0), topology=("Periodic", "Flat", "Bounded"))
    assert g.N == (10, 1, 10) and g.H == (3, 0, 3) and g.spacing == (1000.0, 1.0, 10.0)
    assert g.center_shape() == (16, 1, 16) and g.face_shape(2) == (16, 1, 17) and g.face_shape(0) == (16, 1, 16)
    np.testing.assert_allclose(g.nodes(0)[3:5], [-4500.0, -3500.0])
    s = lfb200.RectilinearGrid(size=(8, 8, 4), extent=(8, 8, 4)).slab(1, 2)
    assert s.Ny == 4 and s.origin[1] == 4.0
    with pytest.raises(ValueError):
        lfb200.RectilinearGrid(size=(10,), topology=("Periodic", "Flat", "Bounded"), x=(0, 1))


def test_weight_function_and_frequency_response():
    fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2.0)
    t = np.linspace(-40, 40, 80001)
    G = lfb200.get_weight_function(t, 0.0, fp)
    assert abs(np.trapezoid(G, t) - 1.0) < 1e-6          # unit gain at zero frequency
    H = lfb200.get_offline_frequency_response(np.array([0.0, 2.0, 20.0]), fp)
    np.testing.assert_allclose(H, [1.0, 0.5, 1.0 / (1 + 10.0 ** 4)], rtol=1e-12)


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    """include/lfb200.h is the drop-in boundary: the built library must export every function it declares
    (no compute calls here: there is no GPU in the CPU test tier)."""
    from lfb200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "lfb200.h")).read()
    declared = sorted(set(re.findall(r"\b(lfb_[a-z0-9_]+)\s*\(", header)) - {"lfb_handle", "lfb_desc"})
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in lfb200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    L = _lib.load()
    assert L.lfb_version() == 110
    assert ctypes.sizeof(_lib.Desc) == 3 * 12 + 4 + 24 + 12 + 4 + 4 * 64 + 4 * 6 + 4 + 4 + 8 + 32
    # without a GPU creation must fail loudly, never fall back
    import torch
    if not torch.cuda.is_available():
        d = _lib.Desc()
        for ax in range(3):
            d.size[ax], d.halo[ax], d.spacing[ax] = 8, 3, 1.0
        d.n_vars, d.velocity_mask, d.n_coeffs, d.n_slots, d.advection = 1, 7, 1, 2, 1
        h = ctypes.c_void_p()
        rc = L.lfb_create(ctypes.byref(d), ctypes.byref(h))
        assert rc == -2 and b"no CPU fallback" in L.lfb_last_error()


def test_integration_shim_binds_every_declared_symbol():
    """INTEGRATION.md shows the reference-side binding (Julia ccall stubs): every function of include/lfb200.h has one."""
    header = open(os.path.join(ROOT, "include", "lfb200.h")).read()
    declared = sorted(set(re.findall(r"\b(lfb_[a-z0-9_]+)\s*\(", header)) - {"lfb_handle", "lfb_desc"})
    shim = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    bound = set(re.findall(r"ccall\(\(:(lfb_[a-z0-9_]+), ", shim))
    missing = [name for name in declared if name not in bound]
    assert not missing, f"no ccall stub in INTEGRATION.md for {missing}"
    assert not sorted(bound - set(declared)), "INTEGRATION.md binds symbols the header does not declare"


def test_remap_padding_deviation_is_flagged():
    """The one deliberate deviation of the device remap is announced exactly where it applies."""
    from lfb200.post import pads_documented_column
    assert not pads_documented_column(("Periodic", "Flat", "Bounded"))        # every example of the reference
    assert not pads_documented_column(("Periodic", "Periodic", "Bounded"))
    assert not pads_documented_column(("Bounded", "Bounded", "Bounded"))
    assert pads_documented_column(("Bounded", "Periodic", "Bounded"))
    assert pads_documented_column(("Bounded", "Flat", "Periodic"))
    assert not pads_documented_column(("Flat", "Periodic", "Bounded"))


def test_immersed_grid_host_evaluation():
    """ImmersedBoundaryGrid on the host: GridFittedBottom marks the cells whose centre lies at or below the bottom,
    PartialCellBottom the cells whose top face does (with the lowest wet cell capped at the minimum fraction of its
    height); periodic halos carry the images; a y-slab's arrays are cut out of the global ones."""
    g0 = lfb200.RectilinearGrid(size=(16, 12, 8), extent=(16.0, 12.0, 8.0), topology=("Periodic", "Periodic", "Bounded"))
    bottom = lambda x, y: -8 + 3.3 * np.exp(-((x - 6) ** 2 + (y - 1) ** 2) / 8)
    gf = lfb200.ImmersedBoundaryGrid(g0, lfb200.GridFittedBottom(bottom))
    gp = lfb200.ImmersedBoundaryGrid(g0, lfb200.PartialCellBottom(bottom, minimum_fractional_cell_height=0.2))
    assert gf.N == g0.N and gf.spacing == g0.spacing and gf.cell_heights() is None
    F, P, Hh = gf.immersed_flags(), gp.immersed_flags(), gp.cell_heights()
    assert F.shape == g0.center_shape() and set(np.unique(F)) == {0.0, 1.0}
    zc = g0.nodes(2, with_halos=False)
    x, y = g0.nodes(0, with_halos=False), g0.nodes(1, with_halos=False)
    zb = bottom(x[:, None], y[None, :])
    assert np.array_equal(F[3:-3, 3:-3, 3:-3], (zc[None, None, :] <= zb[:, :, None]).astype(float))
    assert P[3:-3, 3:-3, 3:-3].sum() <= F[3:-3, 3:-3, 3:-3].sum() + 16 * 12       # top-face test: at most one cell per column fewer / more
    assert np.array_equal(F[:3], F[16:19]) and np.array_equal(F[:, -3:], F[:, 3:6])     # periodic images
    inner_h = Hh[3:-3, 3:-3, 3:-3]
    assert inner_h.min() >= 0.2 * 1.0 - 1e-12 and inner_h.max() == 1.0
    for r in range(3):
        s = gp.slab(r, 3)
        assert s.center_shape() == (22, 10, 14)
        assert np.array_equal(s.immersed_flags(), P[:, 4 * r:4 * r + 10, :])
        assert np.array_equal(s.cell_heights(), Hh[:, 4 * r:4 * r + 10, :])
        assert np.array_equal(s.below_bottom_mask(), gp.below_bottom_mask()[:, 4 * r:4 * r + 10, :])


class _FakeModel:
    """Stands in for LagrangianFilter in the host-logic tests of the snapshot window (no device)."""

    def __init__(self, n_slots, names=("u", "w", "b")):
        self.n_slots = n_slots
        self.velocities_present, self.var_names = [n for n in names if n in "uvw"], [n for n in names if n not in "uvw"]
        self.slots = {}                 # slot -> (pass time, {name: file index})
        self.uploads, self.direction, self.waits = [], 0, 0

    def set_direction(self, d):
        self.direction = d

    def set_slot_time(self, slot, t):
        self.slots[slot] = (t, self.slots[slot][1])

    def invalidate_slot(self, slot):
        self.slots.pop(slot, None)

    def wait_transfers(self):
        self.waits += 1

    def upload_snapshot(self, slot, name, arr, t, wait=False):
        held = self.slots.get(slot, (t, {}))[1] if self.slots.get(slot, (None,))[0] == t else {}
        held[name] = int(arr[0])
        self.slots[slot] = (t, held)
        self.uploads.append((slot, name, int(arr[0]), t))


class _FakeSource:
    def __init__(self, times):
        self.times = list(times)

    def snapshot(self, name, n):
        return np.array([n], dtype=np.float32)


def test_snapshot_window_resident_and_streamed():
    """SnapshotFeeder (the reference's FieldTimeSeries InMemory(4) window, utils:211-212; backward pass = the same
    snapshots re-tagged T_end - t, utils:97-180): resident series are uploaded once and only re-tagged for the
    backward pass; streamed series keep the snapshots bracketing every step resident and never use more than n_slots."""
    from lfb200.offline import SnapshotFeeder
    times = [10.0, 11.0, 12.0, 13.0, 14.0]
    idx = [0, 1, 2, 3, 4]
    # all resident
    m = _FakeModel(n_slots=5)
    f = SnapshotFeeder(m, _FakeSource(times), idx, times, 10.0, 14.0)
    f.begin(+1)
    assert len(m.uploads) == 5 * 3 and m.direction == 1
    assert [m.slots[s][0] for s in range(5)] == [0.0, 1.0, 2.0, 3.0, 4.0]
    f.ensure(0.0, 4.0)
    f.begin(-1)
    assert len(m.uploads) == 15, "the backward pass must not upload again"
    assert [m.slots[s][0] for s in range(5)] == [4.0, 3.0, 2.0, 1.0, 0.0] and m.direction == -1
    assert all(m.slots[s][1] == {"u": s, "w": s, "b": s} for s in range(5))
    # streamed through two slots, forward then reversed
    for direction in (+1, -1):
        m = _FakeModel(n_slots=2)
        f = SnapshotFeeder(m, _FakeSource(times), idx, times, 10.0, 14.0)
        f.begin(direction)
        t = 0.0
        while t < 4.0 - 1e-12:
            f.ensure(t, t + 0.25)
            assert len(m.slots) <= 2
            tags = sorted(tag for tag, _ in m.slots.values())
            assert tags[0] <= t + 1e-12 and tags[-1] >= t + 0.25 - 1e-12, (direction, t, tags)
            for tag, held in m.slots.values():
                want = int(round(tag)) if direction > 0 else 4 - int(round(tag))     # file index behind a pass time
                assert held == {"u": want, "w": want, "b": want}
            t += 0.25
        assert len(m.uploads) == 5 * 3, "every snapshot travels once per pass"
    # a step that spans more snapshots than there are slots is refused with a clear message
    m = _FakeModel(n_slots=2)
    f = SnapshotFeeder(m, _FakeSource(times), idx, times, 10.0, 14.0)
    f.begin(+1)
    with pytest.raises(RuntimeError, match="slots"):
        f.ensure(0.5, 2.5)


class _FakeKeptModel:
    """Device-resident pass outputs as ids of small numpy arrays; counts allocations and releases."""

    def __init__(self):
        self.store, self.next_id, self.released, self.log = {}, 1, [], []

    def new(self, arr):
        kid = self.next_id
        self.next_id += 1
        self.store[kid] = np.array(arr, dtype=np.float64)
        return kid

    def kept_combine(self, f, lo, hi, w, sign):
        b = self.store[lo] if hi is None else (1.0 - w) * self.store[lo] + w * self.store[hi]
        self.log.append((f, lo, hi, w, sign))
        return self.new(self.store[f] + sign * b)

    def kept_gather(self, kid, root):
        return kid

    def kept_release(self, kid):
        assert kid in self.store, f"released twice or never allocated: {kid}"
        self.released.append(kid)
        del self.store[kid]

    def kept_fetch(self, kid, alloc=None, wait=True):
        return self.store[kid].copy()

    def wait_transfers(self):
        pass


def test_forward_backward_combiner_plan_and_block_bookkeeping():
    """sum_forward_backward_contributions! (post:45-170) on device-resident outputs: out(t_k) = fwd(t_k) +- bwd(T - t_k)
    (minus for the mean velocities), every forward time combined as soon as its backward partner is fed, and every
    device block released exactly once."""
    from types import SimpleNamespace
    from lfb200.post import ForwardBackwardCombiner
    cfg = SimpleNamespace(label="", advection="WENO", velocity_names=("u",), compute_mean_velocities=True, T=4.0,
                          var_names_to_filter=("b",), map_to_mean=True)
    names = ["b_Lagrangian_filtered", "xi_u", "u_Lagrangian_filtered"]
    m = _FakeKeptModel()
    times = [0.0, 1.0, 2.0, 3.0, 4.0]
    val = lambda pas, t, q: np.array([100.0 * pas + 10.0 * t + q, 1.0])
    fwd = {t: {n: m.new(val(1, t, q)) for q, n in enumerate(names)} for t in times}
    c = ForwardBackwardCombiner(cfg, fwd, {t: int(8 * t) for t in times}, m)
    combined_after = []
    for tb in times:
        c.feed(tb, {n: m.new(val(2, tb, q)) for q, n in enumerate(names)})
        combined_after.append(sorted(c.done))
    # forward time t_k waits for backward pass time T - t_k: the last forward time is combined first
    assert combined_after == [[4], [3, 4], [2, 3, 4], [1, 2, 3, 4], [0, 1, 2, 3, 4]]
    out = c.finish()
    assert out.times == times and out.iterations == [0, 8, 16, 24, 32]
    for k, t in enumerate(times):
        for q, n in enumerate(names):
            sign = -1.0 if n == "u_Lagrangian_filtered" else 1.0
            np.testing.assert_array_equal(out.fields[n][k], val(1, t, q) + sign * val(2, 4.0 - t, q))
    assert not m.store, f"device blocks never released: {sorted(m.store)}"
    assert len(m.released) == len(set(m.released)) == 3 * (5 + 5 + 5)       # forward, backward and combined blocks
    assert all(hi is None and w == 0.0 for _, _, hi, w, _ in m.log)         # same aligned steps: no interpolation needed
