"""Pins the CPU oracle (oracle/) to every golden vector the reference holds for the hot path
(SURVEY.md section 8c): coefficient known-answers from the reference docstrings, the five pinned
offline fields of test/data/reference_offline_output.jld2 (incl. halos), the remapped `*_at_mean`
fields and the Eulerian comparison fields."""
import numpy as np
import pytest

import lf_oracle as lo
import remap_oracle as ro
from conftest import GOLDEN_GRID, GOLDEN_RUN

FIELDS = ["b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered", "w_Lagrangian_filtered"]


def test_coefficients_known_answers():
    # reference OfflineLagrangianFilter.jl:212 (N=2, freq_c=5e-5)
    p = lo.bw2_params(2, 1e-4 / 2)
    assert p["a1"] == p["b1"] == 1.767766952966369e-5
    assert p["c1"] == p["d1"] == 3.535533905932738e-5
    assert p["N_coeffs"] == 1
    # docs/src/literated/offline_filter_geostrophic_adjustment.md:270 (N=2, freq_c=2.5e-5)
    p = lo.bw2_params(2, 1e-4 / 4)
    assert p["a1"] == 8.838834764831844e-6 and p["c1"] == 1.767766952966369e-5
    # docs/src/literated/offline_filter_lee_wave.md:382 (N=4, freq_c=5e-5)
    p = lo.bw2_params(4, 5e-5)
    want = (4.783542904563622e-6, 1.1548494156391084e-5, 1.913417161825449e-5, 4.619397662556434e-5,
            1.1548494156391084e-5, 4.783542904563623e-6, 4.619397662556434e-5, 1.9134171618254493e-5)
    got = tuple(p[k] for k in ("a1", "b1", "c1", "d1", "a2", "b2", "c2", "d2"))
    np.testing.assert_allclose(got, want, rtol=2e-16)
    # single exponential (utils:262-266)
    p = lo.bw2_params(1, 2.0)
    assert p == {"a1": 1.0, "c1": 2.0, "N_coeffs": 0.5}
    with pytest.raises(ValueError):
        lo.bw2_params(3, 1.0)
    # normalisation sum (a c + b d)/(c^2+d^2) = 1/2 (OfflineLagrangianFilter.jl:408-422)
    for N in (2, 4, 6, 8):
        p = lo.bw2_params(N, 3e-5)
        s = sum((p[f"a{i}"] * p[f"c{i}"] + p[f"b{i}"] * p[f"d{i}"]) / (p[f"c{i}"] ** 2 + p[f"d{i}"] ** 2)
                for i in range(1, N // 2 + 1))
        assert abs(s - 0.5) < 1e-15


def test_step_sequence_matches_golden_iterations(golden_offline):
    # 80 steps for 72 nominal; outputs at the golden file's iteration keys
    seq = lo.step_sequence(86400.0, 1200.0, 3600.0, 8640.0)
    assert len(seq) == 80
    t, keys = 0.0, [0]
    for n, s in enumerate(seq, 1):
        t += s
        if abs(t / 3600.0 - round(t / 3600.0)) < 1e-12:
            keys.append(n)
    assert keys == list(golden_offline["iterations"])


@pytest.fixture(scope="module")
def oracle_run(golden_sim):
    nt = len(golden_sim["t"])
    vel = {k: [golden_sim[k][n].astype(np.float64) for n in range(nt)] for k in ("u", "w")}
    vars_ = [[golden_sim["b"][n].astype(np.float64) for n in range(nt)]]
    fp = lo.bw2_params(GOLDEN_RUN["N_filter"], GOLDEN_RUN["freq_c"])
    return lo.run_offline(golden_sim["t"], vel, vars_, ["b"], N=GOLDEN_GRID["N"], H=GOLDEN_GRID["H"],
                          topology=GOLDEN_GRID["topology"], delta=GOLDEN_GRID["delta"], filter_params=fp,
                          dt=GOLDEN_RUN["dt"], T_out=GOLDEN_RUN["T_out"], velocities="uw")


@pytest.mark.parametrize("name", FIELDS)
def test_offline_fields_match_golden(oracle_run, golden_offline, name):
    mine = np.stack(oracle_run[name])
    ref = golden_offline[name]
    assert mine.shape == ref.shape
    # exact after the Float32 rounding of the per-pass outputs; halos included
    assert np.linalg.norm(mine - ref) <= 1e-12 * np.linalg.norm(ref)


def test_known_answer_values(oracle_run, golden_offline):
    # t = 43200 s (13th output), interior cells (1,6),(2,6),(3,6),(1,1), H = 3
    n = 12
    cells = [(1, 6), (2, 6), (3, 6), (1, 1)]
    want = {"b_Lagrangian_filtered": [-2.838078944478184e-05, -7.992282553459518e-05, -9.990392209147103e-05,
                                      -5.282615893520415e-05],
            "xi_u": [-7.943302154541016, 4.83015513420105, 0.0, 142.7440528869629],
            "xi_w": [0.3708693515509367, 2.6299325227737427, 4.016641974449158, 0.3918922543525696]}
    for name, vals in want.items():
        got = [oracle_run[name][n][i + 2, 0, k + 2] for (i, k) in cells]
        np.testing.assert_allclose(got, vals, rtol=1e-13, atol=0)


def test_remap_matches_golden(golden_offline):
    names = ["b_Lagrangian_filtered", "u_Lagrangian_filtered", "w_Lagrangian_filtered"]
    for n in range(0, 25, 3):
        fields = {k: golden_offline[k][n] for k in names}
        xi = {"u": golden_offline["xi_u"][n], "w": golden_offline["xi_w"][n]}
        out = ro.remap_to_mean(fields, xi, GOLDEN_GRID["N"], GOLDEN_GRID["H"], GOLDEN_GRID["topology"],
                               GOLDEN_GRID["origin"], GOLDEN_GRID["delta"], npad=5)
        for k in names:
            ref = golden_offline[k + "_at_mean"][n]
            got = out[k + "_at_mean"]
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got - ref)) <= 1e-13 * np.nanmax(np.abs(ref))


def test_eulerian_filter_matches_golden(golden_offline):
    fp = lo.bw2_params(GOLDEN_RUN["N_filter"], GOLDEN_RUN["freq_c"])
    for nm in ("b", "u", "w"):
        ef = ro.eulerian_filter([golden_offline[nm][n] for n in range(25)], golden_offline["t"],
                                [fp["a1"]], [fp["b1"]], [fp["c1"]], [fp["d1"]])
        ref = golden_offline[nm + "_Eulerian_filtered"]
        assert np.linalg.norm(np.stack(ef) - ref) <= 1e-13 * np.linalg.norm(ref)


# ---------------------------------------------------------------------------------------------------------------
# The golden case extruded into 3-D (tests/golden_extrude.py): pins the oracle's y-direction / 3-D code, which no
# reference test exercises, to the reference's own vectors.
@pytest.mark.parametrize("mode", ["y-replicated", "x-to-y"])
@pytest.mark.parametrize("halo", [3, 4])
def test_extruded_golden_case_matches_golden(golden_sim, golden_offline, mode, halo):
    import golden_extrude as ge
    grid, times, fields, vel_names, rename = ge.extrude(golden_sim, mode, n_rep=6, halo=halo)
    N = grid["size"]
    delta = (1000.0, 1000.0, 10.0)
    fp = lo.bw2_params(GOLDEN_RUN["N_filter"], GOLDEN_RUN["freq_c"])
    vel = {k: [a.astype(np.float64) for a in fields[k]] for k in vel_names}
    vars_ = [[a.astype(np.float64) for a in fields["b"]]]
    res = lo.run_offline(times, vel, vars_, ["b"], N=N, H=(halo,) * 3, topology=grid["topology"], delta=delta,
                         filter_params=fp, dt=GOLDEN_RUN["dt"], T_out=GOLDEN_RUN["T_out"], velocities="".join(vel_names))
    checked = 0
    for name, series in res.items():
        if name == "t":
            continue
        gname = rename.get(name, name)
        for n in (0, 7, 12, 24):
            ref = ge.golden_rehaloed(golden_offline[gname][n], gname, halo)
            for col in ge.columns_of(series[n], mode, halo):
                assert col.shape == ref.shape
                assert np.linalg.norm(col - ref) <= 1e-12 * np.linalg.norm(ref), (name, n)
                checked += 1
    assert checked == 5 * 4 * 6
