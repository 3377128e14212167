"""Regenerates the golden fixtures in this directory from the reference's own test data.

Run in the build container only (needs /root/reference; the GPU box has no reference tree):

    python tests/golden/make_golden.py

Sources (data files, not source code):
  /root/reference/test/data/reference_sim.jld2             input of test/test_offline_integration.jl:1-32
  /root/reference/test/data/reference_offline_output.jld2  its pinned output (written by the reference under
                                                           Julia 1.10.9; generator test/reference_sim_offline.jl:20-83)
  /root/reference/test/data/reference_online_output.jld2   pinned output of test/test_online_integration.jl:113-128

Each .npz holds, per variable, one array stacked over the 25 output times in Julia (x,y,z) order
with halos (axis 0 = time), plus "t" and "iterations".  dtypes are kept as stored in the .jld2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import lfb200  # noqa: E402  (loader shim for the package directory)
from lfb200.jld2 import JLD2File  # noqa: E402

SRC = "/root/reference/test/data"


def convert(name):
    f = JLD2File(os.path.join(SRC, name + ".jld2"))
    out = {"iterations": np.array(f.iterations("t"), dtype=np.int64), "t": f.times()}
    for var in f.keys("timeseries"):
        if var == "t":
            continue
        its = f.iterations(var)
        first = f[f"timeseries/{var}/{its[0]}"]
        if np.isscalar(first):
            out[var] = np.array([f[f"timeseries/{var}/{it}"] for it in its])
        else:
            out[var] = np.stack([np.asarray(f[f"timeseries/{var}/{it}"]) for it in its], axis=0)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    for n in ("reference_sim", "reference_offline_output", "reference_online_output"):
        convert(n)
