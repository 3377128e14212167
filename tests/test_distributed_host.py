"""Host logic of the multi-GPU path on CPU: world_size-2 `gloo` processes exercise the slab
decomposition helpers, the unique-id broadcast plumbing and the neighbour pattern of the halo exchange
(the same send/recv order liblfb200 uses with NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lfb200
from lfb200.distributed import gather_parent, slab_of_parent, slab_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lfb200.distributed import exchange_halos_host, init_model_comm
        topo = ("Periodic", "Periodic" if periodic else "Bounded", "Bounded")
        g = lfb200.RectilinearGrid(size=(8, 12, 4), extent=(8.0, 12.0, 4.0), topology=topo)
        rng = np.random.default_rng(7)
        glob = np.asfortranarray(rng.standard_normal(g.center_shape()))
        H = g.Hy
        if periodic:                       # global periodic halos
            glob[:, :H, :] = glob[:, g.Ny:g.Ny + H, :]
            glob[:, g.Ny + H:, :] = glob[:, H:2 * H, :]
        local = slab_of_parent(glob, g, rank, world).copy()
        want = local.copy()
        # wipe the halo rows that neighbours own, then exchange
        ny = g.Ny // world
        if periodic or rank > 0:
            local[:, :H, :] = np.nan
        if periodic or rank < world - 1:
            local[:, ny + H:, :] = np.nan
        exchange_halos_host(local, H, rank, world, periodic)
        ok_exchange = np.array_equal(local, want)

        # unique-id plumbing with a stand-in model (no GPU here)
        class FakeModel:
            got = None

            @staticmethod
            def comm_unique_id():
                return bytes(range(128))

            def comm_init(self, uid, r, w):
                FakeModel.got = (uid, r, w)

        m = FakeModel()
        init_model_comm(m, rank, world)
        ok_uid = FakeModel.got == (bytes(range(128)), rank, world)
        # gather: rank 0 assembles the global array back
        parts = [None] * world
        dist.all_gather_object(parts, want)
        ok_gather = np.array_equal(gather_parent(parts, g), glob) if rank == 0 else True
        # the pipeline's helpers: rank / world discovery, per-rank view of the input series, output times sharded
        # over the ranks and merged on rank 0
        from lfb200.distributed import SlabSource, merge_outputs_on_root, owner_of_time, rank_world
        ok_rw = rank_world() == (rank, world) and rank_world(0, 1) == (0, 1)
        src = SlabSource(lfb200.InputData([0.0, 1.0], {"b": [glob, 2 * glob]}), g, rank, world)
        ok_src = np.array_equal(src.snapshot("b", 1), slab_of_parent(2 * glob, g, rank, world)) and src.has("b")
        times = [0.0, 1.0, 2.0, 3.0, 4.0]
        mine = [k for k in range(5) if owner_of_time(k, world) == rank]
        part = lfb200.FilterOutput([times[k] for k in mine], [10 * k for k in mine])
        part.fields["f"] = [np.full((2, 2), float(k)) for k in mine]
        merged = merge_outputs_on_root(part, rank, world)
        if rank == 0:
            ok_merge = (merged.times == times and merged.iterations == [0, 10, 20, 30, 40] and
                        [a[0, 0] for a in merged.fields["f"]] == [0.0, 1.0, 2.0, 3.0, 4.0])
        else:
            ok_merge = merged is part
        q.put((rank, ok_exchange, ok_uid, ok_gather and ok_rw and ok_src and ok_merge, g.slab(rank, world).origin[1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [True, False])
def test_two_rank_slab_logic(periodic):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    for rank, ok_exchange, ok_uid, ok_gather, origin_y in res:
        assert ok_exchange and ok_uid and ok_gather
        assert origin_y == rank * 6.0


def test_slab_helpers():
    g = lfb200.RectilinearGrid(size=(4, 8, 4), extent=(4.0, 8.0, 4.0), topology=("Periodic", "Bounded", "Bounded"))
    assert slab_rows(8, 1, 2) == (4, 8)
    with pytest.raises(ValueError):
        slab_rows(9, 0, 2)
    v = np.zeros(g.face_shape(1))
    assert slab_of_parent(v, g, 0, 2, "v").shape[1] == 4 + 6
    assert slab_of_parent(v, g, 1, 2, "v").shape[1] == 4 + 6 + 1     # the wall face row stays with the last rank
    c = np.arange(np.prod(g.center_shape()), dtype=float).reshape(g.center_shape(), order="F")
    parts = [slab_of_parent(c, g, r, 4) for r in range(4)]
    assert np.array_equal(gather_parent(parts, g), c)
