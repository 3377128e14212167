"""Host I/O: the JLD2 (HDF5-subset) writer and reader (reference on-disk layout
post_processing_utils.jl:395-404, 751-755; run_offline_lagrangian_filter.jl:77-80)."""
import os

import numpy as np
import pytest

import lfb200
from lfb200.jld2 import JLD2File, JLD2Writer, lookup3

# Known-answer vectors for the checksum, cut out of the reference's test/data/reference_sim.jld2 (written by JLD2.jl
# under Julia 1.10.9): the version-2 superblock (44 bytes + checksum) and the object header of `timeseries/t/0`.
SUPERBLOCK = bytes.fromhex("894844460d0a1a0a020808000002000000000000ffffffffffffffffd0ef01000000000030000000000000007de1489c")
OBJECT_HEADER = bytes.fromhex("4f48445202004a05020000030901040000020000000314000131203f000800000000004000340b0034ff030000"
                              "080c00000400080000000000000000000010000000000000000000000000000000000000103fe9dd")


def test_lookup3_matches_checksums_written_by_jld2_jl():
    assert lookup3(SUPERBLOCK[:-4]) == int.from_bytes(SUPERBLOCK[-4:], "little")
    assert lookup3(OBJECT_HEADER[:-4]) == int.from_bytes(OBJECT_HEADER[-4:], "little")
    assert lookup3(b"") == 0xDEADBEEF


def test_writer_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    a32 = np.asfortranarray(rng.standard_normal((16, 1, 17)).astype(np.float32))
    a64 = np.asfortranarray(rng.standard_normal((16, 5, 16)))
    path = str(tmp_path / "out.jld2")
    with JLD2Writer(path) as w:
        for it in (0, 3, 6, 10):
            w.write(f"timeseries/t/{it}", 1200.0 * it)
            w.write(f"timeseries/w/{it}", a32 * (it + 1))
            w.write(f"timeseries/b_Lagrangian_filtered_at_mean/{it}", a64 - it)
        w.write("direction", 1)
        w.write("deep/er/group/" + "n" * 300, np.arange(6, dtype=np.int64).reshape(2, 3))
    raw = open(path, "rb").read()
    assert raw.startswith(b"HDF5-based Julia Data Format, version 0.2.0\x00") and raw[512:520] == b"\x89HDF\r\n\x1a\n"
    assert lookup3(raw[512:556]) == int.from_bytes(raw[556:560], "little")          # superblock checksum
    f = JLD2File(path)
    assert f.keys("timeseries") == ["t", "w", "b_Lagrangian_filtered_at_mean"]
    assert f.iterations("w") == [0, 3, 6, 10]
    np.testing.assert_array_equal(f.times(), [0.0, 3600.0, 7200.0, 12000.0])
    got = f["timeseries/w/6"]
    assert got.dtype == np.float32 and got.flags.f_contiguous and np.array_equal(got, a32 * 7)
    assert np.array_equal(f["timeseries/b_Lagrangian_filtered_at_mean/10"], a64 - 10)
    assert f["direction"] == 1
    assert np.array_equal(f["deep/er/group/" + "n" * 300], np.arange(6).reshape(2, 3))
    # every object header carries a valid checksum (JLD2.jl verifies them)
    pos = raw.find(b"OHDR")
    n = 0
    while pos >= 0:
        size = int.from_bytes(raw[pos + 6:pos + 10], "little")
        end = pos + 10 + size
        assert lookup3(raw[pos:end]) == int.from_bytes(raw[end:end + 4], "little")
        n += 1
        pos = raw.find(b"OHDR", end)
    assert n >= 4 * 3 + 6


def test_writer_rejects_unsupported_values(tmp_path):
    w = JLD2Writer(str(tmp_path / "x.jld2"))
    with pytest.raises(TypeError):
        w.write("a", np.array(["text"]))
    w.write("g/x", 1.0)
    with pytest.raises(ValueError):
        w.write("g/x/y", 2.0)


def test_filter_output_file_has_the_reference_layout(tmp_path, golden_offline):
    """FilterOutput.save writes `timeseries/<name>/<iteration>` + `timeseries/t/<iteration>` exactly as the reference's
    combined file names them; loading it back gives the same arrays.  Checked with the golden file's own names,
    iteration keys, shapes and element types."""
    its = [int(i) for i in golden_offline["iterations"]]
    out = lfb200.FilterOutput(list(golden_offline["t"]), its)
    names = [k for k in golden_offline.files if k not in ("t", "iterations")]
    for name in names:
        out.fields[name] = [np.asfortranarray(a) for a in golden_offline[name]]
    path = out.save(str(tmp_path / "filtered_output.jld2"))
    f = JLD2File(path)
    assert sorted(f.keys("timeseries")) == sorted(names + ["t"])
    assert f.iterations("t") == its == f.iterations("b_Lagrangian_filtered_at_mean")
    back = lfb200.FilterOutput.load(path)
    assert back.times == out.times and back.iterations == its
    for name in names:
        for a, b in zip(back[name], out[name]):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    # the .npz form keeps the same keys
    z = np.load(out.save(str(tmp_path / "filtered_output.npz")))
    assert f"timeseries/xi_u/{its[3]}" in z.files and z[f"timeseries/t/{its[3]}"] == out.times[3]


def test_config_reads_a_series_written_by_the_writer(tmp_path, golden_sim):
    """The input side of the file boundary: a simulation output written with the JLD2 layout (timeseries/<name>/<iter>,
    timeseries/t/<iter>) is what OfflineFilterConfig(original_data_filename=...) opens (reference load_data,
    utils:201-215); times, iteration keys and snapshots come back unchanged."""
    path = str(tmp_path / "sim.jld2")
    nt = len(golden_sim["t"])
    its = [3 * n for n in range(nt)]
    with JLD2Writer(path) as w:
        for n, it in enumerate(its):
            w.write(f"timeseries/t/{it}", float(golden_sim["t"][n]))
            for name in ("u", "w", "b"):
                w.write(f"timeseries/{name}/{it}", np.asfortranarray(golden_sim[name][n]))
    g = lfb200.RectilinearGrid(size=(10, 10), x=(-5000, 5000), z=(-100, 0), topology=("Periodic", "Flat", "Bounded"))
    cfg = lfb200.OfflineFilterConfig(original_data_filename=path, grid=g, var_names_to_filter=("b",), velocity_names=("u", "w"),
                                     N=2, freq_c=1e-4 / 4)
    src = cfg.source()
    np.testing.assert_array_equal(src.times, golden_sim["t"])
    assert cfg.T_start == golden_sim["t"][0] and cfg.T_end == golden_sim["t"][-1] and cfg.T_out == golden_sim["t"][1] - golden_sim["t"][0]
    a = src.snapshot("w", 7)
    assert a.dtype == golden_sim["w"].dtype and np.array_equal(a, golden_sim["w"][7])
    with pytest.raises(ValueError, match="not found in original data"):
        lfb200.OfflineFilterConfig(original_data_filename=path, grid=g, var_names_to_filter=("T",), velocity_names=("u", "w"),
                                   N=2, freq_c=1e-4 / 4)
