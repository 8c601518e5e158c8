"""CPU tier: the N>1 host logic (sharding + tree replication) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lagrange_b200.distributed import broadcast_packed, interleaved_layers, shard_range, slab_range


def test_ranges_tile_exactly():
    for n in (0, 1, 7, 512, 1000003):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    for nz in (512, 100, 37):
        for world in (2, 4, 8):
            cuts = [slab_range(nz, r, world, align=8) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nz
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert all(a % 8 == 0 for a, _ in cuts)


def test_interleaved_layers_tile_exactly_and_balance():
    for nz in (512, 100, 37, 8, 1):
        for world in (1, 2, 3, 4, 8):
            covered = np.zeros(nz, dtype=int)
            sizes = []
            for r in range(world):
                rs = interleaved_layers(nz, r, world, depth=8)
                sizes.append(sum(b - a for a, b in rs))
                for a, b in rs:
                    assert 0 <= a < b <= nz and (a % 8 == 0)
                    covered[a:b] += 1
            assert np.all(covered == 1)
            assert max(sizes) - min(sizes) <= 8
    assert interleaved_layers(512, 0, 1) == [(0, 512)]
    assert interleaved_layers(64, 1, 2) == [(8, 16), (24, 32), (40, 48), (56, 64)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nz, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 owns the "packed tree"; everybody must end up with the same bytes
        blob = None
        if rank == 0:
            blob = np.frombuffer(np.random.Generator(np.random.PCG64(5)).bytes(100003), dtype=np.uint8).copy()
        got = broadcast_packed(blob, src=0, device=torch.device("cpu"))
        expect = np.frombuffer(np.random.Generator(np.random.PCG64(5)).bytes(100003), dtype=np.uint8)
        assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), expect)
        # every rank "classifies" its slab; no collective on the data path; results are gathered only for the check
        z0, z1 = slab_range(nz, rank, world, align=4)
        mine = torch.arange(z0, z1, dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([z1 - z0]))
        assert sum(int(s) for s in sizes) == nz
        np.save(os.path.join(out_dir, f"slab{rank}.npy"), mine.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world, nz = 2, 37
    mp.spawn(_worker, args=(world, _free_port(), nz, str(tmp_path)), nprocs=world, join=True)
    slabs = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)])
    assert np.array_equal(slabs, np.arange(nz))


def test_diagonal_shard_layout_tiles_the_lattice_exactly():
    """wn_grid_shard_layout (host code, no GPU): over all ranks every (tile layer, y part) is owned exactly once, the parts a rank
    owns cycle through all of them, and whole-layer striding is the Q = 1 special case."""
    from lagrange_b200.winding import FastWindingNumber as W

    for dims, world, want_q in (((512, 512, 512), 8, 4), ((512, 512, 512), 4, 4), ((512, 512, 512), 2, 2), ((96, 64, 100), 8, 4),
                                ((40, 24, 37), 8, 1), ((64, 48, 64), 6, 2), ((64, 64, 64), 3, 1), ((64, 64, 64), 1, 1)):
        cover = np.zeros((dims[2], dims[1]), dtype=np.int32)
        total = 0
        for rank in range(world):
            lay = W.shard_layout(dims, rank, world)
            assert lay["parts_y"] == want_q and lay["layer_step"] * lay["parts_y"] == world
            pts = 0
            for z0, z1, y0, y1 in lay["units"]:
                cover[z0:z1, y0:y1] += 1
                pts += (z1 - z0) * (y1 - y0) * dims[0]
            assert pts == lay["n_points"]
            total += pts
            if lay["parts_y"] > 1 and lay["n_units"] >= lay["parts_y"]:
                assert len({u[2] for u in lay["units"]}) == lay["parts_y"]  # every part of the lattice shows up on every rank
        assert np.all(cover == 1) and total == dims[0] * dims[1] * dims[2]
