"""Host-side helpers for the multi-GPU y-slab decomposition (one process per GPU, torch.distributed for
the plumbing, NCCL send/recv inside liblfb200 for the halo exchange).

The reference has no multi-GPU driver of its own; with an Oceananigans `Distributed` architecture every
rank owns a local grid and local fields (reference compute_lagrangian_filter_buffer_tendencies.jl:26-84 only
supplies the strip ranges).  The same convention is used here: rank r owns rows [r Ny/R, (r+1) Ny/R) of
the global grid and works on local parent arrays (its rows plus Hy halo rows on each side).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .grids import Bounded, RectilinearGrid


def slab_rows(Ny: int, rank: int, nranks: int):
    if Ny % nranks:
        raise ValueError(f"Ny={Ny} is not divisible by {nranks} ranks")
    ny = Ny // nranks
    return rank * ny, (rank + 1) * ny


def slab_of_parent(parent: np.ndarray, grid: RectilinearGrid, rank: int, nranks: int, field: str = "c") -> np.ndarray:
    """The local parent array of `rank` cut out of a GLOBAL parent array (halos included).  `field` is "c"
    for centre fields or "u"/"v"/"w"; a v field on a Bounded y axis keeps its extra face row on the last rank."""
    if nranks == 1:
        return np.asfortranarray(parent)
    j0, j1 = slab_rows(grid.Ny, rank, nranks)
    H = grid.Hy
    extra = 1 if (field == "v" and grid.topology[1] == Bounded and rank == nranks - 1) else 0
    return np.asfortranarray(parent[:, j0:j1 + 2 * H + extra, :])


def gather_parent(slabs: Sequence[np.ndarray], grid: RectilinearGrid) -> np.ndarray:
    """Inverse of slab_of_parent for centre fields: global parent array from the ranks' local ones (low halo
    rows from rank 0, high halo rows from the last rank, interior rows from their owners)."""
    R = len(slabs)
    if R == 1:
        return np.asfortranarray(slabs[0])
    H = grid.Hy
    ny = grid.Ny // R
    shape = list(slabs[0].shape)
    shape[1] = grid.Ny + 2 * H
    out = np.zeros(shape, dtype=slabs[0].dtype, order="F")
    for r, s in enumerate(slabs):
        lo = 0 if r == 0 else H
        hi = s.shape[1] if r == R - 1 else H + ny
        out[:, r * ny + lo:r * ny + hi, :] = s[:, lo:hi, :]
    return out


def init_model_comm(model, rank: int, world: int):
    """Creates the NCCL communicator of the model's handle: rank 0 makes the unique id, torch.distributed
    (any backend) broadcasts it, every rank joins."""
    import torch.distributed as dist
    if world == 1:
        return
    payload: List[object] = [type(model).comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    model.comm_init(payload[0], rank, world)


def exchange_halos_host(local: np.ndarray, H: int, rank: int, world: int, periodic: bool):
    """Reference implementation of the slab halo exchange on HOST arrays over torch.distributed (used by the
    CPU tests of the decomposition logic): fills the Hy halo rows of `local` from the neighbours' edge rows."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    lo = rank - 1 if rank > 0 else (world - 1 if periodic else None)
    hi = rank + 1 if rank < world - 1 else (0 if periodic else None)
    ny = local.shape[1] - 2 * H
    send_lo = torch.from_numpy(np.ascontiguousarray(local[:, H:2 * H, :]))
    send_hi = torch.from_numpy(np.ascontiguousarray(local[:, ny:ny + H, :]))
    recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
    ops = []
    if lo is not None:
        ops.append(dist.P2POp(dist.isend, send_lo, lo, tag=0))
    if hi is not None:
        ops.append(dist.P2POp(dist.isend, send_hi, hi, tag=1))
    if hi is not None:
        ops.append(dist.P2POp(dist.irecv, recv_hi, hi, tag=0))
    if lo is not None:
        ops.append(dist.P2POp(dist.irecv, recv_lo, lo, tag=1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    if lo is not None:
        local[:, 0:H, :] = recv_lo.numpy()
    if hi is not None:
        local[:, ny + H:ny + 2 * H, :] = recv_hi.numpy()
    return local


# ---------------------------------------------------------------------------------------------- pipeline helpers
def rank_world(rank=None, world=None):
    """(rank, world) of this process: the arguments if given, else torch.distributed's when a process group is
    initialised, else a single rank."""
    if rank is not None and world is not None:
        return int(rank), int(world)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def owner_of_time(k: int, world: int) -> int:
    """Rank that gathers, remaps and holds output time k: the remap is independent per output time (and per
    variable), so it shards as replicas without a collective."""
    return k % world


class SlabSource:
    """The input series as rank `rank` sees it: every snapshot cut down to the rank's y-slab (its rows plus Hy halo
    rows on each side, which hold the neighbours' rows / the periodic images / the global halo)."""

    def __init__(self, source, grid, rank: int, world: int):
        self.source, self.grid, self.rank, self.world = source, grid, rank, world
        self.times = source.times

    def has(self, name):
        return self.source.has(name)

    def snapshot(self, name, n):
        arr = self.source.snapshot(name, n)
        if self.world == 1:
            return arr
        return slab_of_parent(np.asarray(arr), self.grid, self.rank, self.world, name if name in ("u", "v", "w") else "c")


def merge_outputs_on_root(out, rank: int, world: int, grid=None):
    """Every rank holds the output times it owns (owner_of_time) as global arrays, or its y-slab of every output time
    (`out.slab`); rank 0 receives all of them (torch.distributed gather_object, any backend) and returns the complete
    FilterOutput, the others return theirs unchanged."""
    import torch.distributed as dist
    payload = {"times": out.times, "iterations": out.iterations, "fields": out.fields, "slab": out.slab}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return out
    merged = type(out)([], [])
    if gathered[0]["slab"] is not None:
        merged.times, merged.iterations = list(gathered[0]["times"]), list(gathered[0]["iterations"])
        for name in gathered[0]["fields"]:
            merged.fields[name] = [gather_parent([part["fields"][name][n] for part in gathered], grid)
                                   for n in range(len(merged.times))]
        return merged
    rows = []
    for part in gathered:
        for n, (t, it) in enumerate(zip(part["times"], part["iterations"])):
            rows.append((t, it, {name: series[n] for name, series in part["fields"].items()}))
    rows.sort(key=lambda r: r[0])
    merged.times = [r[0] for r in rows]
    merged.iterations = [r[1] for r in rows]
    for name in rows[0][2] if rows else ():
        merged.fields[name] = [r[2][name] for r in rows]
    return merged
