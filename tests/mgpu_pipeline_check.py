"""Multi-GPU check of the WHOLE offline pipeline through the public API, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_pipeline_check.py

Every rank calls run_offline_Lagrangian_filter(config) on its y-slab: forward pass, backward pass, device-side
combination, gather by output time, remap to the mean position sharded by output time, merge on rank 0.  Rank 0 then
runs the same configuration undecomposed on its own GPU and compares every output field at every output time:
bit-exact with the strict kernel, <= 1e-12 (relative L2) with the TMA z-march kernel; the `_at_mean` fields through
the same bound.  Prints one line per case from rank 0 and exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import lfb200
from lfb200.synthetic import make_series


def run_cases(rank, world, local, cases=None):
    """Returns (ok, lines); usable from bench.py (small grid, before the timed region)."""
    ok, lines = True, []
    cases = cases or (("Periodic", "generic", True, (24, 8 * world, 10), True),
                      ("Bounded", "generic", True, (24, 8 * world, 10), True),
                      ("Periodic", "auto", False, (40, 16 * world, 16), True),
                      ("Bounded", "auto", False, (40, 16 * world, 16), False))     # no remap: outputs stay in slabs
    for topo_y, kernel, strict, size, regrid in cases:
        g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=("Periodic", topo_y, "Bounded"))
        var_names, vel_names = ("T", "b"), ("u", "v", "w")
        times = [0.0, 0.5, 1.0, 1.5, 2.0]
        fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=4.0, amplitude=0.6)
        data = lfb200.InputData(times, {k: [np.asarray(a, dtype=np.float32) for a in v] for k, v in fields.items()})
        cfg = lfb200.OfflineFilterConfig(data=data, grid=g, var_names_to_filter=var_names, velocity_names=vel_names,
                                         N=2, freq_c=2 * np.pi / 2.0, dt=0.1, T_out=0.5, device=local, strict=strict,
                                         kernel=kernel, n_slots=5, npad=3)
        out = lfb200.run_offline_Lagrangian_filter(cfg, save=False, regrid=regrid)     # collective: all ranks, y-slabs
        flag = torch.zeros(1, device="cuda")
        worst, kname, worst_name = 0.0, out.stats["kernel"], ""
        if rank == 0:
            ref = lfb200.run_offline_Lagrangian_filter(cfg, save=False, regrid=regrid, rank=0, world=1)
            good = out.times == ref.times and out.iterations == ref.iterations and set(out.keys()) == set(ref.keys())
            for name in ref.keys():
                for a, b in zip(out[name], ref[name]):
                    if not good:
                        break
                    if strict:
                        same = np.array_equal(a, b, equal_nan=True)
                        if not same and float(np.nanmax(np.abs(a - b))) >= worst:
                            worst, worst_name = float(np.nanmax(np.abs(a - b))), name
                        good = good and same
                    else:
                        m = np.isfinite(b)
                        if not np.array_equal(m, np.isfinite(a)):
                            good = False
                            break
                        e = float(np.linalg.norm(a[m] - b[m]) / max(np.linalg.norm(b[m]), 1e-300))
                        worst = max(worst, e)
                        good = good and e < 1e-12
            if not good:
                flag += 1
        dist.all_reduce(flag)
        ok = ok and flag.item() == 0
        lines.append(f"mgpu_pipeline world={world} y={topo_y} kernel={kname} remap={int(regrid)} fields={len(out.keys())} "
                     f"times={len(out.times)}: worst={worst:.3e} {worst_name} {'OK' if flag.item() == 0 else 'MISMATCH'}")
    return ok, lines


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok, lines = run_cases(rank, world, local)
    if rank == 0:
        for line in lines:
            print(line, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
