"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank steps its y-slab of a small 3-D problem (NCCL halo exchange overlapped with the interior)
and, on the same GPU, the undecomposed problem; the slab must reproduce its rows of the global run:
bit-exact with the strict generic kernel, <= 1e-12 with the TMA z-march kernel.  Prints one line per
kernel from rank 0 and exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import lfb200
from lfb200.distributed import init_model_comm, slab_of_parent
from lfb200.model import LagrangianFilter
from lfb200.synthetic import make_series


def log(*a):
    if os.environ.get("MGPU_VERBOSE"):
        print(f"[rank {os.environ.get('RANK')}]", *a, file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    all_ok = True
    for topo_y, kernel, strict, size in (("Periodic", "generic", True, (32, 16 * world, 10)),
                                         # immersed grid (whole cells): flags and face codes of a slab follow the GLOBAL indices
                                         ("Periodic", "ztma-immersed", False, (64, 16 * world, 16)),
                                         ("Bounded", "generic-immersed", True, (32, 16 * world, 10)),
                                         ("Bounded", "generic", True, (32, 16 * world, 10)),
                                         ("Periodic", "ztma", False, (64, 32 * world, 24)),
                                         ("Periodic", "ztma", False, (32, 20 * world, 12)),
                                         ("Bounded", "ztma", False, (40, 12 * world, 12)),
                                         # slabs tall enough for the split launch (wall strips on the general
                                         # instantiation, interior rows on the all-periodic one)
                                         ("Bounded", "ztma", False, (40, 40 * world, 12))):
        g = lfb200.RectilinearGrid(size=size, extent=tuple(float(s) for s in size), topology=("Periodic", topo_y, "Bounded"))
        if kernel.endswith("-immersed"):
            kernel = kernel.split("-")[0]
            Lx, Ly, zb, Lz = g.Lx, g.Ly, g.origin[2], g.Lz
            g = lfb200.ImmersedBoundaryGrid(g, lfb200.GridFittedBottom(
                lambda x, y: zb + 0.4 * Lz * np.exp(-((x - 0.4 * Lx) ** 2 + (y - 0.5 * Ly) ** 2) / (2 * (0.2 * Lx) ** 2))))
        var_names, vel_names = ("T", "b"), ("u", "v", "w")
        times = [0.0, 0.5, 1.0]
        fields = make_series(g, times, var_names=var_names, velocity_names=vel_names, T_period=2.0)
        fp = lfb200.set_offline_BW2_filter_params(N=2, freq_c=2 * np.pi / 2.0)
        gl = g.slab(rank, world)
        ms = LagrangianFilter(gl, var_names_to_filter=var_names, velocity_names=vel_names, filter_params=fp, n_slots=3,
                              device=local, strict=strict, kernel=kernel)
        log('created slab model'); init_model_comm(ms, rank, world); log('comm ready')
        mg = LagrangianFilter(g, var_names_to_filter=var_names, velocity_names=vel_names, filter_params=fp, n_slots=3,
                              device=local, strict=strict, kernel=kernel)
        for s, t in enumerate(times):
            for name in vel_names + var_names:
                mg.upload_snapshot(s, name, fields[name][s], t)
                ms.upload_snapshot(s, name, slab_of_parent(fields[name][s], g, rank, world, name if name in vel_names else "c"), t)
        log('uploaded'); ms.init_from_data(); log('slab init done'); mg.init_from_data()
        for dt in (0.1, 0.1, 0.05, 0.25, 0.2, 0.3):
            ms.step(dt); mg.step(dt)
        ms.sync(); log('stepped')
        worst, ok = 0.0, True
        for name in ms.tracer_names:
            got = ms.tracer(name)
            want = slab_of_parent(mg.tracer(name), g, rank, world)
            if topo_y == "Bounded":
                # rows outside the global domain: the slab's outer bounded halo rows are not exchanged
                H = g.Hy
                lo = H if rank == 0 else 0
                hi = got.shape[1] - (H if rank == world - 1 else 0)
                got, want = got[:, lo:hi, :], want[:, lo:hi, :]
            if strict:
                same = np.array_equal(got, want)
                worst = max(worst, 0.0 if same else float(np.abs(got - want).max()))
                ok = ok and same
            else:
                e = float(np.linalg.norm(got - want) / np.linalg.norm(want))
                worst = max(worst, e)
                ok = ok and e < 1e-12
        flag = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            print(f"mgpu_check world={world} y={topo_y} kernel={ms.kernel_name}: worst={worst:.3e} "
                  f"{'OK' if flag.item() == 0 else 'MISMATCH'}", flush=True)
        ms.close(); mg.close()
        all_ok = all_ok and flag.item() == 0
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
